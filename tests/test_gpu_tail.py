"""GPU parity of the loss / optimiser tail of an iteration (train.py:286-300; csrc/loss.cu, csrc/optim.cu): the fused
BCE-with-logits, the Gram tail of common_loss, and clip_grad_norm_ + Adam in two launches -- each against torch's own
expression of the reference's calls (nn.BCEWithLogitsLoss, utils.common_loss, nn.utils.clip_grad_norm_ + optim.Adam)."""
import copy

import numpy as np
import pytest
import torch as th
import torch.nn.functional as F

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from dreamgnn_b200 import _lib
    _lib.load()
    return th.device('cuda:0')


# ---- BCE -------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 7, 2430, 467641, 3_000_001])
@pytest.mark.parametrize('smoothing', [0.0, 0.1])
def test_bce_with_logits_forward_backward(dev, n, smoothing):
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(n)
    x = (th.randn(n, generator=gen, device=dev) * 4).requires_grad_(True)
    if n > 4:
        with th.no_grad():
            x[0], x[1], x[2], x[3] = 0.0, 60.0, -60.0, 1e-8       # saturated / tiny logits
    t = (th.rand(n, generator=gen, device=dev) < 0.3).float()
    loss = ops.FusedBCEWithLogitsLoss(smoothing)(x, t)
    x64 = x.detach().double().requires_grad_(True)
    ref = F.binary_cross_entropy_with_logits(x64, t.double() * (1 - smoothing) + 0.5 * smoothing)
    assert abs(float(loss) - float(ref)) <= 2e-7 * max(1.0, abs(float(ref)))
    (loss * 3.0).backward()                                          # an upstream gradient other than one
    (ref * 3.0).backward()
    assert H.rel_err(x.grad, x64.grad) <= 1e-6
    assert th.isfinite(x.grad).all()
    # same value as the fp32 module the reference calls, to fp32 rounding of a mean over n terms
    ref32 = th.nn.BCEWithLogitsLoss()(x.detach(), t * (1 - smoothing) + 0.5 * smoothing)
    assert abs(float(loss) - float(ref32)) <= 1e-5 * max(1.0, abs(float(ref32)))


def test_bce_shapes_and_errors(dev):
    from dreamgnn_b200 import ops
    x = th.randn(6, 5, device=dev)
    t = (th.rand(6, 5, device=dev) < 0.5).float()
    assert abs(float(ops.bce_with_logits(x, t)) - float(F.binary_cross_entropy_with_logits(x, t))) < 1e-6
    assert abs(float(ops.bce_with_logits(x.t(), t.t())) - float(F.binary_cross_entropy_with_logits(x, t))) < 1e-6
    with pytest.raises(ValueError):
        ops.bce_with_logits(x, t[:3])
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.bce_with_logits(x.cpu(), t.cpu())


# ---- common loss (Gram tail) -------------------------------------------------------------------------
@pytest.mark.parametrize('n,d', [(45, 16), (763, 128), (5000, 128)])
def test_gram_common_loss_tail(dev, n, d):
    """utils.py:87-95 in float64 on the host expression vs the Gram form whose tail is dg_gram_common_loss_f64."""
    from dreamgnn_b200 import ops
    from dreamgnn_b200.utils import common_loss
    gen = th.Generator(dev).manual_seed(n + d)
    a = th.randn(n, d, generator=gen, device=dev).requires_grad_(True)
    b = (0.5 * a.detach() + th.randn(n, d, generator=gen, device=dev)).requires_grad_(True)
    loss = ops.gram_common_loss(a, b)
    a64, b64 = a.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    ref = common_loss(a64, b64)
    assert abs(float(loss) - float(ref)) <= 1e-6 * abs(float(ref))
    (loss * 0.7).backward()
    (ref * 0.7).backward()
    assert H.rel_err(a.grad, a64.grad) <= 1e-5 and H.rel_err(b.grad, b64.grad) <= 1e-5


# ---- clip + Adam -----------------------------------------------------------------------------------
def _param_lists(dev, shapes, seed):
    gen = th.Generator(dev).manual_seed(seed)
    ps = [th.randn(*s, generator=gen, device=dev) for s in shapes]
    return [p.clone().requires_grad_(True) for p in ps], [p.clone().requires_grad_(True) for p in ps], gen


SHAPES_MODEL = [(2, 2), (2, 768, 341), (128, 341), (128,), (763, 768), (768,), (768, 128), (128, 256), (16, 128), (16,),
                (1, 16), (64, 128), (1, 64), (1,)]


@pytest.mark.parametrize('shapes,max_norm,wd', [
    (SHAPES_MODEL, 1.0, 1e-5),                       # the reference's call: clip 1.0, weight decay 1e-5 (train.py:217, 299)
    (SHAPES_MODEL, 1e9, 0.0),                        # never clips
    (SHAPES_MODEL, 0.0, 1e-5),                       # no clipping requested
    ([(4096,), (4097,), (1,), (3, 5, 7)] + [(33,)] * 60, 0.5, 1e-2),     # > 48 tensors: several launches per pass
])
def test_fused_adam_matches_clip_grad_norm_and_adam(dev, shapes, max_norm, wd):
    from dreamgnn_b200.optim import FusedAdam
    mine, theirs, gen = _param_lists(dev, shapes, seed=len(shapes))
    opt = FusedAdam(mine, lr=0.002, weight_decay=wd)
    ref = th.optim.Adam(theirs, lr=0.002, weight_decay=wd)
    for it in range(6):
        scale = 10.0 ** (it % 3 - 1)                  # gradient norms on both sides of the clip threshold
        for a, b in zip(mine, theirs):
            g = th.randn(a.shape, generator=gen, device=dev) * scale
            a.grad, b.grad = g.clone(), g.clone()
        norm = opt.clip_and_step(max_norm)
        if max_norm > 0:
            ref_norm = th.nn.utils.clip_grad_norm_(theirs, max_norm)
        else:
            ref_norm = th.linalg.vector_norm(th.stack([th.linalg.vector_norm(p.grad) for p in theirs]))
        ref.step()
        assert abs(float(norm) - float(ref_norm)) <= 2e-6 * float(ref_norm)
        for a, b in zip(mine, theirs):
            assert H.rel_err(a.grad, b.grad) <= 1e-6                  # the clipped gradient stays in .grad
            assert H.rel_err(a.detach(), b.detach()) <= 1e-6
            assert H.rel_err(opt.state[a]['exp_avg'], ref.state[b]['exp_avg']) <= 1e-6
            assert H.rel_err(opt.state[a]['exp_avg_sq'], ref.state[b]['exp_avg_sq']) <= 1e-6
    assert float(opt.state[mine[0]]['step']) == 6.0


def test_fused_adam_skips_missing_gradients_and_reads_a_device_learning_rate(dev):
    from dreamgnn_b200.optim import FusedAdam
    mine, theirs, gen = _param_lists(dev, [(50, 7), (9,), (300,)], seed=3)
    lr = th.tensor(0.01, device=dev)
    opt = FusedAdam(mine, lr=lr)
    ref = th.optim.Adam(theirs, lr=0.01)
    for it in range(4):
        if it == 2:
            lr.mul_(0.5)                                              # what ReduceLROnPlateau does to a tensor lr
            ref.param_groups[0]['lr'] = 0.005
        for i, (a, b) in enumerate(zip(mine, theirs)):
            if i == 1:
                a.grad = b.grad = None                                # a parameter the loss does not reach
                continue
            g = th.randn(a.shape, generator=gen, device=dev)
            a.grad, b.grad = g.clone(), g.clone()
        opt.step()
        ref.step()
    for a, b in zip(mine, theirs):
        assert H.rel_err(a.detach(), b.detach()) <= 1e-6
    assert mine[1] not in opt.state or 'exp_avg' not in opt.state[mine[1]]


def test_fused_adam_state_dict_interchanges_with_torch_adam(dev):
    from dreamgnn_b200.optim import FusedAdam
    mine, theirs, gen = _param_lists(dev, [(20, 30), (30,)], seed=8)
    opt = FusedAdam(mine, lr=0.002, weight_decay=1e-5)
    grads = [[th.randn(p.shape, generator=gen, device=dev) for p in mine] for _ in range(5)]
    for g in grads[:3]:
        for p, x in zip(mine, g):
            p.grad = x.clone()
        opt.clip_and_step(1.0)
    # FusedAdam -> torch.optim.Adam
    ref = th.optim.Adam(theirs, lr=0.002, weight_decay=1e-5)
    with th.no_grad():
        for a, b in zip(mine, theirs):
            b.copy_(a)
    ref.load_state_dict(copy.deepcopy(opt.state_dict()))
    # ... and back into a fresh FusedAdam
    again = [p.detach().clone().requires_grad_(True) for p in mine]
    opt2 = FusedAdam(again, lr=0.002, weight_decay=1e-5)
    opt2.load_state_dict(copy.deepcopy(opt.state_dict()))
    for g in grads[3:]:
        for a, b, c, x in zip(mine, theirs, again, g):
            a.grad, b.grad, c.grad = x.clone(), x.clone(), x.clone()
        opt.clip_and_step(1.0)
        opt2.clip_and_step(1.0)
        th.nn.utils.clip_grad_norm_(theirs, 1.0)
        ref.step()
    for a, b, c in zip(mine, theirs, again):
        assert H.rel_err(a.detach(), b.detach()) <= 1e-6
        assert th.equal(a.detach(), c.detach())
    assert float(opt2.state[again[0]]['step']) == 5.0


def test_fused_adam_has_no_cpu_path():
    from dreamgnn_b200.optim import FusedAdam
    p = th.zeros(4, requires_grad=True)
    p.grad = th.ones(4)
    with pytest.raises(RuntimeError, match='CUDA'):
        FusedAdam([p]).step()


def test_train_iteration_with_fused_tail_tracks_the_torch_tail(dev):
    """The same seeded model trained for a few iterations (dropout off, augmentation off) with the fused loss / clip /
    Adam and with torch's own calls: identical loss curves to fp32 rounding."""
    import argparse as ap
    from dreamgnn_b200 import ops, synthetic
    from dreamgnn_b200.model import Net
    from dreamgnn_b200.optim import FusedAdam
    from dreamgnn_b200.train import train_iteration
    from dreamgnn_b200.utils import common_loss
    spec = dict(kind='dense', n_drug=90, n_dis=70, n_pos=400, f_drug=48, f_dis=48, k=4)
    w = synthetic.make_workload(spec, dev, seed=5)
    state = synthetic.train_state(w, dev)
    margs = synthetic.model_args(w, gcn_agg_units=96, gcn_out_units=16, nhid1=40, nhid2=16)
    margs.dropout = 0.0
    margs.attention_dropout = 0.0
    curves = []
    for fused in (False, True):
        th.manual_seed(11)
        model = Net(margs).to(dev)
        if fused:
            opt, fn = FusedAdam(model.parameters(), lr=0.002, weight_decay=1e-5), ops.FusedBCEWithLogitsLoss()
        else:
            opt, fn = th.optim.Adam(model.parameters(), lr=0.002, weight_decay=1e-5), th.nn.BCEWithLogitsLoss()
        curves.append([float(train_iteration(model, opt, state, fn, [], {}, 0.001, 1.0, common_loss)) for _ in range(8)])
    a, b = np.array(curves[0]), np.array(curves[1])
    assert np.all(np.abs(a - b) <= 2e-5 * np.abs(a)), (a, b)
    assert b[-1] < b[0]


# ---- basis decomposition of the relation weights ---------------------------------------------------
@pytest.mark.parametrize('R,B,rows,D,mult', [(2, 2, 768, 341, 4), (2, 2, 128, 128, 4), (3, 2, 5, 7, 8), (4, 4, 33, 12, 32),
                                             (5, 3, 9, 6, 4)])
def test_basis_combine_matches_matmul(dev, R, B, rows, D, mult):
    """layers.py:120-121: W = matmul(att, basis.view(B, -1)).view(R, in, msg), here at the padded message width, against
    the same expression in float64 (value and both gradients); the padded columns are exactly zero."""
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(R * 100 + D)
    att = th.randn(R, B, generator=gen, device=dev).requires_grad_(True)
    basis = th.randn(B, rows, D, generator=gen, device=dev).requires_grad_(True)
    w = ops.basis_combine(att, basis, mult)
    dp = D + (-D) % mult
    assert w.shape == (R, rows, dp) and w.is_contiguous()
    assert float(w[:, :, D:].abs().sum()) == 0.0
    att64, basis64 = att.detach().double().requires_grad_(True), basis.detach().double().requires_grad_(True)
    ref = th.matmul(att64, basis64.view(B, -1)).view(R, rows, D)
    assert H.rel_err(w[:, :, :D], ref) <= 1e-6
    g = th.randn(R, rows, dp, generator=gen, device=dev)
    w.backward(g)
    ref.backward(g[:, :, :D].double())
    assert H.rel_err(att.grad, att64.grad) <= 2e-6
    assert H.rel_err(basis.grad, basis64.grad) <= 1e-6


def test_weighted_loss_sum(dev):
    from dreamgnn_b200 import ops
    for device in (dev, th.device('cpu')):
        xs = [th.tensor(v, device=device, requires_grad=True) for v in (0.7, 2.5, -1.25)]
        total = ops.WeightedLossSum.apply(xs[0], xs[1], xs[2], 0.001)
        assert abs(float(total) - (0.7 + 0.001 * (2.5 - 1.25))) < 1e-7
        (total * 2.0).backward()
        assert [round(float(x.grad), 7) for x in xs] == [2.0, 0.002, 0.002]


def test_cuda_graph_iteration_with_fused_tail(dev):
    """The captured iteration with the fused loss / clip / Adam (what train() --cuda_graph and bench.py run): replays draw
    fresh seeds from the per-iteration pool, the step counter and moments advance on the device, and the loss falls to the
    level of the eager loop with torch's own tail."""
    import argparse as ap
    from dreamgnn_b200 import ops, synthetic
    from dreamgnn_b200.graphed import GraphedIteration
    from dreamgnn_b200.model import Net
    from dreamgnn_b200.optim import FusedAdam
    from dreamgnn_b200.train import aug_params_from_args, train_iteration
    from dreamgnn_b200.utils import common_loss
    spec = dict(kind='dense', n_drug=90, n_dis=70, n_pos=400, f_drug=48, f_dis=48, k=4)
    w = synthetic.make_workload(spec, dev, seed=5)
    state = synthetic.train_state(w, dev)
    margs = synthetic.model_args(w, gcn_agg_units=96, gcn_out_units=16, nhid1=40, nhid2=16)
    side = th.cuda.Stream()
    with th.cuda.stream(side):
        th.manual_seed(9)
        model = Net(margs).to(dev)
        opt = FusedAdam(model.parameters(), lr=0.002, weight_decay=1e-5)
        step = GraphedIteration(model, opt, state, rel_loss_fn=ops.FusedBCEWithLogitsLoss())
        first = float(opt.state[next(iter(model.parameters()))]['step'])
        graphed = [float(step().detach()) for _ in range(40)]
        assert float(opt.state[next(iter(model.parameters()))]['step']) == first + 40
        th.manual_seed(9)
        ref_model = Net(margs).to(dev)
        ref_opt = th.optim.Adam(ref_model.parameters(), lr=0.002, weight_decay=1e-5)
        p = aug_params_from_args(ap.Namespace())
        eager = [float(train_iteration(ref_model, ref_opt, state, th.nn.BCEWithLogitsLoss(), ['edge_dropout', 'feature_noise'], p,
                                       0.001, 1.0, common_loss).detach()) for _ in range(40 + step.eager_iterations)]
    th.cuda.current_stream().wait_stream(side)
    assert all(np.isfinite(graphed)) and np.mean(graphed[-10:]) < np.mean(graphed[:10])
    assert len({round(x, 6) for x in graphed[:5]}) == 5                # replays are not identical: fresh random draws
    assert abs(np.mean(graphed[-10:]) - np.mean(eager[-10:])) < 0.05


# ---- small fp32 GEMM ---------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (7, 5, 3), (64, 64, 16), (65, 63, 17), (763, 128, 344), (128, 344, 763), (681, 128, 128),
                                   (128, 128, 5000), (3, 200, 700)])
@pytest.mark.parametrize('ta', [False, True])
@pytest.mark.parametrize('tb', [False, True])
def test_small_gemm_all_layouts(dev, M, N, K, ta, tb):
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(M * 7 + N * 3 + K)
    a = th.randn((K, M) if ta else (M, K), generator=gen, device=dev)
    b = th.randn((K, N) if tb else (N, K), generator=gen, device=dev)
    bias = th.randn(N, generator=gen, device=dev)
    want = (a.double().t() if ta else a.double()) @ (b.double() if tb else b.double().t())
    c = ops.small_gemm(a, b, ta, tb)
    assert c.shape == (M, N) and H.rel_err(c, want) <= 2e-6
    assert th.equal(c, ops.small_gemm(a, b, ta, tb))                  # split-K parts are added in split order: bit-identical
    assert H.rel_err(ops.small_gemm(a, b, ta, tb, bias=bias), want + bias.double()) <= 2e-6
    # row-strided views are read as stored
    wide = th.randn(a.shape[0], a.shape[1] + 5, generator=gen, device=dev)
    wide[:, 2:2 + a.shape[1]] = a
    assert th.equal(ops.small_gemm(wide[:, 2:2 + a.shape[1]], b, ta, tb), c)


@pytest.mark.parametrize('R,M,N,K', [(2, 763, 128, 128), (2, 90, 16, 16), (3, 257, 65, 300), (4, 5, 3, 2)])
def test_small_gemm_batched_and_reduced(dev, R, M, N, K):
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(R + M + N + K)
    x = th.randn(M, K, generator=gen, device=dev)
    w = th.randn(R, K, N, generator=gen, device=dev)
    y = ops.small_gemm(x, w, trans_b=True)                             # shared A, batched B stored [K, N]
    assert y.shape == (R, M, N) and H.rel_err(y, th.matmul(x.double().unsqueeze(0), w.double())) <= 2e-6
    dy = th.randn(R, M, N, generator=gen, device=dev)
    dx = ops.small_gemm(dy, w, reduce_batch=True)                      # sum_r dy[r] @ w[r]^T
    assert dx.shape == (M, K) and H.rel_err(dx, th.einsum('rmn,rkn->mk', dy.double(), w.double())) <= 2e-6
    assert th.equal(dx, ops.small_gemm(dy, w, reduce_batch=True))
    dw = ops.small_gemm(x, dy, trans_a=True, trans_b=True)             # x^T @ dy[r]
    assert dw.shape == (R, K, N) and H.rel_err(dw, th.einsum('mk,rmn->rkn', x.double(), dy.double())) <= 2e-6


@pytest.mark.parametrize('M,K,N,R', [(763, 344, 128, 2), (90, 35, 16, 2), (681, 128, 128, 2)])
def test_small_linear_and_project_autograd(dev, M, K, N, R):
    """ops.linear / ops.project below the tensor-core threshold (the real-dataset layers) against torch in float64:
    value and every gradient."""
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(M + K + N)
    x = th.randn(M, K, generator=gen, device=dev).requires_grad_(True)
    w = (th.randn(N, K, generator=gen, device=dev) / K ** 0.5).requires_grad_(True)
    b = th.randn(N, generator=gen, device=dev).requires_grad_(True)
    wr = (th.randn(R, K, N, generator=gen, device=dev) / K ** 0.5).requires_grad_(True)
    assert M * K * N < ops.GEMM_MIN_MACS
    y, h = ops.linear(x, w, b), ops.project(x, wr)
    x64, w64, b64, wr64 = (t.detach().double().requires_grad_(True) for t in (x, w, b, wr))
    y64, h64 = th.nn.functional.linear(x64, w64, b64), th.matmul(x64.unsqueeze(0), wr64)
    assert H.rel_err(y, y64) <= 2e-6 and H.rel_err(h, h64) <= 2e-6
    gy, gh = th.randn(M, N, generator=gen, device=dev), th.randn(R, M, N, generator=gen, device=dev)
    ((y * gy).sum() + (h * gh).sum()).backward()
    ((y64 * gy.double()).sum() + (h64 * gh.double()).sum()).backward()
    for got, want in ((x.grad, x64.grad), (w.grad, w64.grad), (b.grad, b64.grad), (wr.grad, wr64.grad)):
        assert H.rel_err(got, want) <= 3e-6
    assert H.rel_err(ops.linear(x.detach(), w.detach()), th.nn.functional.linear(x64, w64)) <= 2e-6      # no bias


def test_multi_copy(dev):
    """dg_multi_copy: many tensors in one launch -- every size class (empty, a few bytes, unaligned views, several 16 KB
    chunks), more items than one launch carries, mixed dtypes."""
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(3)
    srcs = [th.randn(n, generator=gen, device=dev) for n in (0, 1, 3, 4, 5, 4095, 4096, 4097, 100003)]
    srcs += [th.randint(0, 1000, (n,), generator=gen, device=dev, dtype=th.int32) for n in (7, 33, 20001)]
    srcs += [th.randint(0, 255, (n,), generator=gen, device=dev, dtype=th.uint8) for n in (1, 15, 17, 16385)]
    big = th.randn(5000, generator=gen, device=dev)
    srcs += [big[1:1 + n] for n in (6, 129, 4001)]                  # 4-byte-aligned but not 16-byte-aligned sources
    srcs += [th.randn(9, generator=gen, device=dev) for _ in range(120)]            # > 96 items: two launches
    dsts = [th.full_like(s, 7) for s in srcs]
    ops.multi_copy(dsts, srcs)
    assert all(th.equal(d, s) for d, s in zip(dsts, srcs))
    with pytest.raises(ValueError):
        ops.multi_copy([th.zeros(4, device=dev)], [th.zeros(5, device=dev)])
