"""GPU parity at BASELINE.json config 4 (synthetic 100k drugs x 50k diseases, ~20 M associations, 1024 / 768-dim
features, k = 15):

  * `test_slice_*`: a proportional 2 % slice (2 000 x 1 000 nodes, ~400 k pairs, same widths / k / mean degree, unequal
    in-dims -> per-etype weights branch, FGCN on the feature matrix) through the whole production path -- every GEMM on
    the tcgen05 kernel (256-row tiles and split-K fire at these shapes), the L2-prefetch SpMM instance forced on, the
    tcgen05 decoder -- against the CPU oracle: forward, loss, every parameter gradient.
  * `test_full_*`: the FULL 20 M-pair shape kernel by kernel -- every SpMM class of the step (GCMC relation blocks at
    d = 344 and d = 128, forward and transposed; valued FGCN at d = 768 with bias + ReLU), the decoder forward and
    backward over all 19.96 M pairs (dz1 is 2.55 G elements: int64 offsets past 2^31), the batched 100 000-row projection
    GEMM with its input- and weight-gradient products -- against float64 evaluations of the same expressions on the
    device (torch.sparse / torch.matmul in float64: test comparands, not product code).
"""
import numpy as np
import pytest
import torch as th

from oracle import restate as R
from tests import helpers as H
from tests import shapes as S

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5


@pytest.fixture(scope='module')
def dev():
    from dreamgnn_b200 import _lib
    _lib.load()
    return th.device('cuda:0')


# ----------------------------------------------------------------------------------------------------
# 2 % slice: whole-step parity against the CPU oracle
# ----------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def slice_case(dev):
    from dreamgnn_b200 import ops, synthetic
    spec = synthetic.scaled('syn20m', 0.02)
    w = synthetic.make_workload(spec, dev, seed=1234)
    saved = (ops.GEMM_MIN_MACS, ops.SPMM_PREFETCH_MIN_BYTES, ops.SPMM_PREFETCH_MIN_BYTES_WIDE)
    ops.GEMM_MIN_MACS, ops.SPMM_PREFETCH_MIN_BYTES, ops.SPMM_PREFETCH_MIN_BYTES_WIDE = 0, 0, 0     # production instances on
    yield spec, w
    ops.GEMM_MIN_MACS, ops.SPMM_PREFETCH_MIN_BYTES, ops.SPMM_PREFETCH_MIN_BYTES_WIDE = saved


def _coo(t):
    idx = t._indices().cpu().numpy()
    return idx[0], idx[1], t._values().cpu().numpy(), t.shape[0]


def test_slice_feature_knn_graphs_bit_exact(slice_case):
    """k = 15 feature-similarity graphs at 2 000 / 1 000 nodes equal the oracle's (neighbour sets and fp32 values)."""
    spec, w = slice_case
    for key, feat in (('drug_feature_graph', w['drug_feat']), ('disease_feature_graph', w['dis_feat'])):
        row, col, val = R.similarity_knn_graph(R.feature_cosine_similarity(feat.double().cpu().numpy()), spec['k'])
        gr, gc, gv, _ = _coo(w[key])
        np.testing.assert_array_equal(gr, row)
        np.testing.assert_array_equal(gc, col)
        np.testing.assert_array_equal(gv, val)


def test_slice_step_parity(slice_case, dev):
    from dreamgnn_b200 import _lib, synthetic
    from dreamgnn_b200.model import Net
    from dreamgnn_b200.utils import common_loss_gram
    spec, w = slice_case
    state = synthetic.train_state(w, dev)
    th.manual_seed(99)
    net = Net(synthetic.model_args(w, dropout=0.0, attention_dropout=0.0)).to(dev).train()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    _lib.reset_launch_count()
    with S.capture_relu_masks() as masks:
        out = net(state.enc_graph, state.dec_graph, state.drug_graph, state.drug_sim_feat, state.drug_feat, state.dis_graph,
                  state.dis_sim_feat, state.dis_feat, state.drug_feature_graph, state.disease_feature_graph)
    loss = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), state.labels) + 0.001 * (
        common_loss_gram(out[1], out[2]) + common_loss_gram(out[3], out[4]))
    loss.backward()
    th.cuda.synchronize()
    assert _lib.launch_count() > 50

    pairs = (w['pairs'][0].cpu().numpy().astype(np.int64), w['pairs'][1].cpu().numpy().astype(np.int64))
    labels = w['labels'].cpu()
    enc = R.enc_graph_from_pairs(pairs, labels.numpy(), spec['n_drug'], spec['n_dis'])
    knn = [_coo(w[k]) for k in ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph')]
    res = {}
    for dt in (th.float32, th.float64):                     # fp32: the reference's own rounding level (its own masks);
        P = S.oracle_params(sd, dt, requires_grad=True)     # float64: the exact value through the forward's masks
        g = dict(enc, ci={k: th.as_tensor(v).to(dt) for k, v in enc['ci'].items()},
                 cj={k: th.as_tensor(v).to(dt) for k, v in enc['cj'].items()})
        kn = [(r, c, th.as_tensor(v).to(dt), n) for r, c, v, n in knn]
        df, sf = w['drug_feat'].cpu().to(dt), w['dis_feat'].cpu().to(dt)
        ref = R.net_forward(P, g, pairs, kn[0], df, df, kn[1], sf, sf, kn[2], kn[3], layers=3, training=True,
                            relu_masks=masks if dt == th.float64 else None)
        rloss = R.training_loss(ref, labels.to(dt))
        rloss.backward()
        res[dt] = ([o.detach() for o in ref], float(rloss.detach()), {k: v.grad for k, v in P.items() if v.grad is not None})
    ref64, loss64, g64 = res[th.float64]
    _, _, g32 = res[th.float32]
    loss_gpu = float(loss.detach())
    assert abs(loss_gpu - loss64) <= 2e-6, (loss_gpu, loss64)
    for nm, a, b in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), out, ref64):
        e = H.rel_err(a.detach().cpu(), b)
        assert e <= FP32_TOL, (nm, e)
    with_grad = {k: p for k, p in net.named_parameters() if p.grad is not None}
    for k, p in net.named_parameters():
        if p.grad is None:
            assert k not in g64 or float(g64[k].abs().max()) == 0.0, k
    # what a CPU fp32 evaluation of the reference's ops achieves; measured on tensors untouched by the FGCN masks, and
    # through a float64 run with the oracle's OWN masks for the FGCN ones
    P0 = S.oracle_params(sd, th.float64, requires_grad=True)
    g0 = dict(enc, ci={k: th.as_tensor(v).double() for k, v in enc['ci'].items()}, cj={k: th.as_tensor(v).double() for k, v in enc['cj'].items()})
    kn0 = [(r, c, th.as_tensor(v).double(), n) for r, c, v, n in knn]
    df0, sf0 = w['drug_feat'].cpu().double(), w['dis_feat'].cpu().double()
    R.training_loss(R.net_forward(P0, g0, pairs, kn0[0], df0, df0, kn0[1], sf0, sf0, kn0[2], kn0[3], layers=3, training=True),
                    labels.double()).backward()
    ref_vs_exact = {k: H.rel_err(g32[k], P0[k].grad) for k in with_grad}
    budget = S.class_budgets(ref_vs_exact, {k: p.numel() for k, p in with_grad.items()})
    rows, failed = [], []
    for k, p in with_grad.items():
        e = H.rel_err(p.grad.cpu(), g64[k])
        rows.append('%-34s vs float64 %.2e (budget %.1e)  fp32 CPU vs float64 %.2e' % (k, e, budget[k], ref_vs_exact[k]))
        if e > budget[k]:
            failed.append(rows[-1])
    print('slice: loss %.6f (float64 oracle %.6f)\n%s' % (float(loss), loss64, '\n'.join(rows)))
    assert not failed, failed


# ----------------------------------------------------------------------------------------------------
# full 20 M-pair shape, kernel by kernel, against float64 on the device
# ----------------------------------------------------------------------------------------------------
N_D, N_S, N_PAIRS = 100_000, 50_000, 20_000_000


@pytest.fixture(scope='module')
def full(dev):
    """Pairs / labels of the full shape (the bench's generator, dreamgnn_b200/synthetic.py) and its encoder graph."""
    from dreamgnn_b200 import graph_build as GB
    gen = th.Generator(dev).manual_seed(1234)
    cells = th.unique(th.randint(0, N_D * N_S, (N_PAIRS,), generator=gen, device=dev))
    labels = (th.rand(cells.numel(), generator=gen, device=dev) < 0.01).float()
    order = th.argsort(labels, descending=True, stable=True)
    cells, labels = cells[order], labels[order].contiguous()
    pairs = ((cells // N_S).to(th.int32), (cells % N_S).to(th.int32))
    enc = GB.generate_enc_graph(pairs, labels, N_D, N_S, dev)
    return pairs, labels, enc, gen


def _coo_f64(csr):
    idx = th.stack([csr.rows().long(), csr.indices.long()])
    vals = csr.vals.double() if csr.vals is not None else th.ones(csr.nnz, dtype=th.float64, device=csr.device)
    return th.sparse_coo_tensor(idx, vals, (csr.n_rows, csr.n_cols)).coalesce()


@pytest.mark.parametrize('dst_type,d', [('disease', 344), ('drug', 344), ('disease', 128), ('drug', 128)])
def test_full_gcmc_spmm(full, dev, dst_type, d):
    """One GCMC relation-block aggregation of the step, forward and (transposed CSR) backward, all ~18-20 M edges."""
    from dreamgnn_b200 import ops
    pairs, labels, enc, gen = full
    blk = enc.block(dst_type)
    csr = blk.csr
    assert csr.nnz == pairs[0].numel() and csr.n_cols == blk.num_rel * blk.n_src
    x = th.randn(csr.n_cols, d, generator=gen, device=dev).requires_grad_(True)
    ss = th.rand(csr.n_cols, generator=gen, device=dev) + 0.5
    ds_ = th.rand(csr.n_rows, generator=gen, device=dev) + 0.5
    out = ops.spmm(csr, x, src_scale=ss, dst_scale=ds_, tag='gcmc')
    gout = th.randn(csr.n_rows, d, generator=gen, device=dev)
    out.backward(gout)
    A = _coo_f64(csr)
    ref = ds_.double()[:, None] * th.sparse.mm(A, ss.double()[:, None] * x.detach().double())
    e_fwd = H.rel_err(out.detach().cpu(), ref.cpu())
    del ref
    rgrad = ss.double()[:, None] * th.sparse.mm(A.t(), ds_.double()[:, None] * gout.double())
    e_bwd = H.rel_err(x.grad.cpu(), rgrad.cpu())
    print('gcmc spmm %s d=%d: forward %.2e, backward %.2e' % (dst_type, d, e_fwd, e_bwd))
    assert e_fwd <= FP32_TOL and e_bwd <= FP32_TOL, (e_fwd, e_bwd)


def test_full_fgcn_spmm(full, dev):
    """Valued kNN aggregation at 100 000 nodes, d = 768 (gathered operand 307 MB > L2), bias + ReLU fused, and its backward."""
    from dreamgnn_b200 import graph_build as GB, layers, ops
    _, _, _, gen = full
    n, k, d = N_D, 15, 768
    step = th.randint(1, n // (k + 1), (n, k), generator=gen, device=dev)
    nbr = (th.arange(n, device=dev).unsqueeze(1) + th.cumsum(step, 1)) % n
    adj = GB.knn_graph_from_topk(th.sort(nbr, dim=1).values.to(th.int32))
    csr = layers.adjacency_csr(adj)
    x = th.randn(n, d, generator=gen, device=dev).requires_grad_(True)
    bias = th.randn(d, generator=gen, device=dev).requires_grad_(True)
    out = ops.spmm(csr, x, bias=bias, relu=True, tag='fgcn')
    gout = th.randn(n, d, generator=gen, device=dev)
    out.backward(gout)
    A = _coo_f64(csr)
    pre = th.sparse.mm(A, x.detach().double()) + bias.detach().double()
    e_fwd = H.rel_err(out.detach().cpu(), th.relu(pre).cpu())
    # A gradient is a discontinuous function of the ReLU mask: an entry whose pre-activation lies within fp32 rounding of
    # zero may land on the other side than in float64, and ONE such entry moves a gradient norm by far more than 1e-5
    # under a random upstream gradient. The backward's contract is dx = A^T (gout * (out > 0)) for the forward's own
    # output, so the comparand takes the mask from `out`; the masks themselves may differ in a handful of entries.
    mask = out.detach() > 0
    flips = int((mask != (pre > 0)).sum())
    gm = gout.double() * mask
    e_b = H.rel_err(bias.grad.cpu(), gm.sum(0).cpu())
    e_x = H.rel_err(x.grad.cpu(), th.sparse.mm(A.t(), gm).cpu())
    print('fgcn spmm d=768: forward %.2e, dbias %.2e, dx %.2e, %d of %d mask entries differ from float64' % (e_fwd, e_b, e_x, flips, mask.numel()))
    assert e_fwd <= FP32_TOL and e_b <= FP32_TOL and e_x <= FP32_TOL and flips <= 64, (e_fwd, e_b, e_x, flips)


def test_full_decoder(full, dev):
    """MLP decoder over all ~19.96 M scored pairs, forward and backward (node gradients through the two deterministic
    segment sums), against a float64 evaluation in 1 M-pair chunks. The upstream gradient is the BCE one,
    (sigmoid(logit) - label) / E, as in training."""
    from dreamgnn_b200 import ops
    pairs, labels, enc, gen = full
    pg = ops.PairGraph(pairs[0], pairs[1], N_D, N_S)
    e = pg.n_pairs
    mk = lambda *shape, s=1.0: (th.randn(*shape, generator=gen, device=dev) * s).requires_grad_(True)
    pd, ps = mk(N_D, 128), mk(N_S, 128)
    w2, b2, w3, b3 = mk(64, 128, s=0.1), mk(64, s=0.1), mk(1, 64, s=0.2), mk(1, s=0.1)
    out = ops.decoder_mlp(pd, ps, w2, b2, w3, b3, pg, p=0.0, training=True)
    assert out.shape == (e, 1)
    # hidden-2 ReLU mask of the forward (what the backward differentiates through; see test_full_fgcn_spmm on why the
    # comparand takes the mask from the forward's own state): bit j of mask_bits[e] <=> z2[e, j] > 0
    mask_bits = ops.decoder_saved_mask(out)
    go = ((th.sigmoid(out.detach().view(-1)) - labels) / e)
    out.backward(go.view(-1, 1))
    W2, B2, W3 = w2.detach().double(), b2.detach().double(), w3.detach().double().view(-1)
    PD, PS = pd.detach().double(), ps.detach().double()
    acc = dict(dpd=th.zeros_like(PD), dps=th.zeros_like(PS), dw2=th.zeros_like(W2), db2=th.zeros_like(B2),
               dw3=th.zeros_like(W3), db3=th.zeros((), dtype=th.float64, device=dev))
    num = den = 0.0
    flips = 0
    bit = th.arange(64, device=dev, dtype=th.int64)
    chunk = 1 << 20
    for c0 in range(0, e, chunk):
        s, d = pairs[0][c0:c0 + chunk].long(), pairs[1][c0:c0 + chunk].long()
        z1 = th.relu(PD[s] + PS[d])
        z2 = th.relu(z1 @ W2.t() + B2)
        ref = z2 @ W3 + b3.detach().double()
        got = out.detach().view(-1)[c0:c0 + chunk].double()
        num += float(((got - ref) ** 2).sum())
        den += float((ref ** 2).sum())
        g = go[c0:c0 + chunk].double()
        m2 = ((mask_bits[c0:c0 + chunk, None] >> bit) & 1).bool()
        flips += int((m2 != (z2 > 0)).sum())
        dz2 = g[:, None] * W3[None, :] * m2
        acc['dw2'] += dz2.t() @ z1
        acc['db2'] += dz2.sum(0)
        acc['dw3'] += g @ z2
        acc['db3'] += g.sum()
        dz1 = (dz2 @ W2) * (z1 > 0)
        acc['dpd'].index_add_(0, s, dz1)
        acc['dps'].index_add_(0, d, dz1)
    errs = {'logits': (num / den) ** 0.5}
    for nm, t in (('dpd', pd), ('dps', ps), ('dw2', w2), ('db2', b2), ('dw3', w3), ('db3', b3)):
        errs[nm] = H.rel_err(t.grad.cpu().reshape(-1), acc[nm].cpu().reshape(-1))
    print('decoder over %d pairs: %s; %d of %d hidden-2 mask entries differ from float64' % (e, {k: '%.2e' % v for k, v in errs.items()}, flips, e * 64))
    assert flips <= 2e-6 * e * 64, flips          # pre-activations within fp32 rounding of zero
    assert max(errs.values()) <= FP32_TOL, errs


def test_full_projection_gemm(full, dev):
    """GCMC layer-0 projection of the 100 000 drugs: x [100k, 1024] @ W [2, 1024, 344] on the tcgen05 kernel (256 x 128
    CTA tiles), with dx (K' = 688) and dW (split-K over the 100 000 rows), against float64 matmuls."""
    from dreamgnn_b200 import ops
    _, _, _, gen = full
    x = th.nn.functional.normalize(th.randn(N_D, 1024, generator=gen, device=dev), dim=1).requires_grad_(True)
    w = (th.randn(2, 1024, 344, generator=gen, device=dev) * 0.05).requires_grad_(True)
    y = ops.project(x, w)
    gy = th.randn(2, N_D, 344, generator=gen, device=dev)
    y.backward(gy)
    x64, w64, g64 = x.detach().double(), w.detach().double(), gy.double()
    errs = {'y': H.rel_err(y.detach().cpu(), th.matmul(x64.unsqueeze(0), w64).cpu()),
            'dx': H.rel_err(x.grad.cpu(), (g64[0] @ w64[0].t() + g64[1] @ w64[1].t()).cpu()),
            'dw': H.rel_err(w.grad.cpu(), th.matmul(x64.t().unsqueeze(0), g64).cpu())}
    print('projection GEMM 100000 x 1024 x (2 x 344):', {k: '%.2e' % v for k, v in errs.items()})
    assert max(errs.values()) <= FP32_TOL, errs
