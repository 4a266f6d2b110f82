"""GPU parity of the fused row kernels (csrc/fused.cu): activation + dropout (layers.py:134-138, 247, 281-282) and the
two-view Attention (layers.py:324-338), forward and backward, against oracle/restate.py evaluated in float64."""
import numpy as np
import pytest
import torch as th

from oracle import restate as R
from tests import helpers as H

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5


@pytest.fixture(scope='module')
def dev():
    from dreamgnn_b200 import _lib
    _lib.load()
    return th.device('cuda:0')


@pytest.mark.parametrize('act', [None, 'leaky', 'relu'])
@pytest.mark.parametrize('shape', [(1, 4), (37, 36), (1000, 344), (5000, 128)])
def test_act_dropout_without_dropout_is_exact(dev, act, shape):
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(shape[0])
    x = th.randn(*shape, generator=gen, device=dev)
    x[0, 0] = 0.0                                                   # the kink: torch's subgradient at 0
    x.requires_grad_(True)
    y = ops.act_dropout(x, act, 0.1, p=0.0, training=True) if act else ops.act_dropout(x, act, 0.1, p=0.0) * 1.0
    ref_in = x.detach().clone().requires_grad_(True)
    ref = ref_in if act is None else (th.nn.functional.leaky_relu(ref_in, 0.1) if act == 'leaky' else th.relu(ref_in))
    assert th.equal(y.detach(), ref.detach())
    g = th.randn(*shape, generator=gen, device=dev)
    y.backward(g)
    (ref * 1.0).backward(g)
    assert th.equal(x.grad, ref_in.grad)


@pytest.mark.parametrize('p', [0.1, 0.3, 0.75])
def test_act_dropout_mask_statistics_and_backward_consistency(dev, p):
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(3)
    x = (th.randn(4096, 344, generator=gen, device=dev).abs() + 0.5) * th.where(th.rand(4096, 344, generator=gen, device=dev) < 0.5, -1.0, 1.0)
    x.requires_grad_(True)
    seed = ops.fresh_seed(dev)
    y = ops.act_dropout(x, 'leaky', 0.1, p=p, training=True, seed=seed)
    kept = y.detach() != 0
    rate = float(kept.float().mean())
    assert abs(rate - (1 - p)) < 4e-3, rate                         # 1.4 M draws: 3 sigma ~ 1.2e-3; 16-bit threshold quantisation
    scale = float((y.detach()[kept] / th.nn.functional.leaky_relu(x.detach(), 0.1)[kept]).mean())
    assert abs(scale * rate - 1.0) < 4e-3                            # unbiased: E[y] = act(x)
    y2 = ops.act_dropout(x.detach(), 'leaky', 0.1, p=p, training=True, seed=seed)
    assert th.equal(y2, y.detach())                                  # same seed -> same mask
    y3 = ops.act_dropout(x.detach(), 'leaky', 0.1, p=p, training=True)
    assert not th.equal(y3, y.detach())                              # fresh seed -> another mask
    g = th.randn(4096, 344, generator=gen, device=dev)
    y.backward(g)
    want = g * th.where(x.detach() > 0, 1.0, 0.1) * kept * scale
    assert H.rel_err(x.grad.cpu(), want.cpu()) <= 1e-6
    # eval mode: no dropout
    assert th.equal(ops.act_dropout(x.detach(), 'leaky', 0.1, p=p, training=False), th.nn.functional.leaky_relu(x.detach(), 0.1))


def _att_params(gen, dev, d, hidden):
    return (th.randn(hidden, d, generator=gen, device=dev) * 0.2, th.randn(hidden, generator=gen, device=dev) * 0.1,
            th.randn(1, hidden, generator=gen, device=dev) * 0.5)


@pytest.mark.parametrize('n,d,hidden', [(1, 128, 16), (763, 128, 16), (45, 16, 16), (60, 8, 16), (1000, 256, 16), (333, 128, 5),
                                        (20000, 128, 16)])
def test_attention_forward_backward(dev, n, d, hidden):
    """Out, beta and every gradient (both views, W1, b1, w2), beta used in the loss as well, vs the float64 oracle."""
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(n + d)
    za = th.randn(n, d, generator=gen, device=dev).requires_grad_(True)
    zb = th.randn(n, d, generator=gen, device=dev).requires_grad_(True)
    w1, b1, w2 = (t.requires_grad_(True) for t in _att_params(gen, dev, d, hidden))
    out, beta = ops.attention_fuse(za, zb, w1, b1, w2)
    gout = th.randn(n, d, generator=gen, device=dev)
    gbeta = th.randn(n, 2, generator=gen, device=dev)
    ((out * gout).sum() + (beta * gbeta).sum()).backward()
    P = {'a.project.0.weight': w1.detach().cpu().double().requires_grad_(True),
         'a.project.0.bias': b1.detach().cpu().double().requires_grad_(True),
         'a.project.2.weight': w2.detach().cpu().double().requires_grad_(True)}
    ra, rb = za.detach().cpu().double().requires_grad_(True), zb.detach().cpu().double().requires_grad_(True)
    rout, rbeta = R.attention(P, 'a.', th.stack([ra, rb], 1))
    ((rout * gout.cpu().double()).sum() + (rbeta.squeeze(-1) * gbeta.cpu().double()).sum()).backward()
    assert H.rel_err(out.detach().cpu(), rout.detach()) <= FP32_TOL
    assert H.rel_err(beta.detach().cpu(), rbeta.detach().squeeze(-1)) <= FP32_TOL
    for nm, a, b in (('dza', za, ra), ('dzb', zb, rb), ('dW1', w1, P['a.project.0.weight']), ('db1', b1, P['a.project.0.bias']),
                     ('dw2', w2, P['a.project.2.weight'])):
        assert H.rel_err(a.grad.cpu(), b.grad) <= FP32_TOL, nm


def test_attention_module_matches_reference_expression_and_is_deterministic(dev):
    """layers.Attention: forward(z) on the stacked [N, 2, D] tensor (the reference's call) == fuse(za, zb) == the torch
    expression of layers.py:334-338; two backward passes are bit-identical."""
    from dreamgnn_b200.layers import Attention
    th.manual_seed(0)
    att = Attention(128, dropout_rate=0.0).to(dev)
    gen = th.Generator(dev).manual_seed(1)
    za, zb = th.randn(500, 128, generator=gen, device=dev), th.randn(500, 128, generator=gen, device=dev)
    z = th.stack((za, zb), 1)
    o1, b1 = att(z)
    o2, b2 = att.fuse(za, zb)
    o3, b3 = att._forward_torch(z)
    assert th.equal(o1, o2) and th.equal(b1, b2) and b1.shape == (500, 2, 1)
    assert H.rel_err(o1.detach().cpu(), o3.detach().cpu()) <= 1e-6 and H.rel_err(b1.detach().cpu(), b3.detach().cpu()) <= 1e-6
    grads = []
    for _ in range(2):
        att.zero_grad()
        att.fuse(za, zb)[0].pow(2).sum().backward()
        grads.append([p.grad.clone() for p in att.parameters()])
    assert all(th.equal(a, b) for a, b in zip(*grads))


@pytest.mark.parametrize('p', [0.1, 0.5])
def test_attention_dropout(dev, p):
    """beta entries are dropped independently (layers.py:336: nn.Dropout on the [N, 2, 1] weights), kept ones scaled by
    1 / (1 - p); the backward differentiates through exactly the forward's mask."""
    from dreamgnn_b200 import ops
    gen = th.Generator(dev).manual_seed(7)
    n, d = 50000, 128
    za = th.randn(n, d, generator=gen, device=dev).requires_grad_(True)
    zb = th.randn(n, d, generator=gen, device=dev).requires_grad_(True)
    w1, b1, w2 = (t.requires_grad_(True) for t in _att_params(gen, dev, d, 16))
    seed = ops.fresh_seed(dev)
    out, beta = ops.attention_fuse(za, zb, w1, b1, w2, p=p, training=True, seed=seed)
    soft = ops.attention_fuse(za.detach(), zb.detach(), w1.detach(), b1.detach(), w2.detach())[1]
    kept = beta.detach() != 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 6e-3
    ratio = beta.detach()[kept] / soft[kept]
    assert float((ratio - ratio.mean()).abs().max()) < 1e-5 and abs(float(ratio.mean()) * (1 - p) - 1.0) < 2e-3
    both = float((kept[:, 0] & kept[:, 1]).float().mean())
    assert abs(both - (1 - p) ** 2) < 8e-3                           # the two views of a node are dropped independently
    gout = th.randn(n, d, generator=gen, device=dev)
    (out * gout).sum().backward()
    # the same computation in torch with the forward's mask
    ra, rb = za.detach().clone().requires_grad_(True), zb.detach().clone().requires_grad_(True)
    rw1, rb1, rw2 = (t.detach().clone().requires_grad_(True) for t in (w1, b1, w2))
    z = th.stack((ra, rb), 1)
    w = th.tanh(z @ rw1.t() + rb1) @ rw2.t()
    rbeta = th.softmax(w, dim=1) * (kept.float() * float(ratio.mean())).unsqueeze(-1)
    ((rbeta * z).sum(1) * gout).sum().backward()
    assert H.rel_err(out.detach().cpu(), (rbeta * z).sum(1).detach().cpu()) <= FP32_TOL
    for nm, a, b in (('dza', za, ra), ('dzb', zb, rb), ('dW1', w1, rw1), ('db1', b1, rb1), ('dw2', w2, rw2)):
        assert H.rel_err(a.grad.cpu(), b.grad.cpu()) <= 2e-5, nm       # comparand is fp32 torch here
