"""GPU parity tests, kernel level: every C-ABI entry point against the CPU oracle (oracle/restate.py,
numpy, scipy) on the same seeded inputs. Bit-exact for integer / index work and for the fp32
normaliser / adjacency values; norm-wise <= 1e-5 for fp32 arithmetic, <= 2e-2 for bf16 storage."""
import numpy as np
import pytest
import torch as th

from oracle import restate as R
from tests import helpers as H

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5      # north_star: within 1e-5 relative for fp32 forward and backward
BF16_TOL = 2e-2      # north_star: within 2e-2 for the bf16 path


@pytest.fixture(scope='module')
def dev():
    from dreamgnn_b200 import _lib
    _lib.load()                      # fail loudly if the extension is missing
    return th.device('cuda:0')


def ops():
    from dreamgnn_b200 import ops as _ops
    return _ops


# ---- index primitives -----------------------------------------------------------------------------
@pytest.mark.parametrize('n', [0, 1, 5, 4096, 4097, 100003, 1 << 21])
def test_exclusive_scan(dev, n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 50, size=n).astype(np.int32)
    got = ops().exclusive_scan_i32(th.tensor(x, device=dev)).cpu().numpy()
    want = np.concatenate([[0], np.cumsum(x, dtype=np.int64)]).astype(np.int32)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('n,bits', [(1, 8), (2049, 13), (50000, 35), (300001, 47)])
def test_radix_sort_is_stable(dev, n, bits):
    rng = np.random.default_rng(bits)
    keys = rng.integers(0, 1 << min(bits, 12), size=n, dtype=np.int64) << max(bits - 12, 0)   # many duplicates
    keys |= rng.integers(0, 4, size=n, dtype=np.int64)
    vals = np.arange(n, dtype=np.int32)
    ko, vo = ops().sort_pairs_u64(th.tensor(keys, device=dev), th.tensor(vals, device=dev), bits)
    order = np.argsort(keys, kind='stable')
    np.testing.assert_array_equal(ko.cpu().numpy(), keys[order])
    np.testing.assert_array_equal(vo.cpu().numpy(), vals[order])


# ---- CSR build / transpose / normalisers ----------------------------------------------------------
@pytest.mark.parametrize('n_rows,n_cols,n_edges', [(7, 5, 0), (1, 1, 1), (45, 120, 2430), (1000, 777, 60000),
                                                  (3, 100000, 5000)])
def test_csr_build_matches_oracle(dev, n_rows, n_cols, n_edges):
    rng = np.random.default_rng(n_edges + 1)
    row = rng.integers(0, n_rows, size=n_edges)
    col = rng.integers(0, n_cols, size=n_edges)                # duplicates and empty rows on purpose
    csr = ops().CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n_rows, n_cols)
    indptr, indices, eid = R.csr_from_pairs(row, col, n_rows)
    np.testing.assert_array_equal(csr.indptr.cpu().numpy(), indptr)
    np.testing.assert_array_equal(csr.indices.cpu().numpy(), indices)
    np.testing.assert_array_equal(csr.eid.cpu().numpy(), eid)
    t = csr.transpose()
    tp, ti, te = R.csr_from_pairs(col, row, n_cols)
    np.testing.assert_array_equal(t.indptr.cpu().numpy(), tp)
    np.testing.assert_array_equal(t.indices.cpu().numpy(), ti)
    np.testing.assert_array_equal(t.eid.cpu().numpy(), te)
    np.testing.assert_array_equal(csr.degree_norm().cpu().numpy(), R.degree_norm(np.diff(indptr)))   # bit-exact
    if n_edges:
        np.testing.assert_array_equal(csr.rows().cpu().numpy(), row[eid])


def test_csr_matches_scipy_on_distinct_pairs(dev):
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    cells = rng.choice(300 * 200, size=9000, replace=False)
    row, col = cells // 200, cells % 200
    csr = ops().CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), 300, 200)
    m = sp.csr_matrix((np.ones(9000), (row, col)), shape=(300, 200))
    m.sort_indices()
    np.testing.assert_array_equal(csr.indptr.cpu().numpy(), m.indptr)
    np.testing.assert_array_equal(csr.indices.cpu().numpy(), m.indices)


def test_degree_norm_exhaustive_small_degrees(dev):
    from dreamgnn_b200 import graph_build
    deg = th.arange(0, 70000, device=dev)
    got = graph_build._norm_from_degrees(deg).cpu().numpy()
    np.testing.assert_array_equal(got, R.degree_norm(np.arange(70000)))


# ---- edge dropout as CSR compaction ----------------------------------------------------------------
@pytest.mark.parametrize('rate', [0.1, 0.5, 0.999])
def test_edge_dropout_compaction(dev, rate):
    rng = np.random.default_rng(11)
    n_rows, n_cols, e = 83, 61, 4000
    row, col = rng.integers(0, n_rows, e), rng.integers(0, n_cols, e)
    val = rng.random(e).astype(np.float32)
    o = ops()
    base = o.CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n_rows, n_cols,
                          th.tensor(val, device=dev))
    perm = th.randperm(e, generator=th.Generator().manual_seed(5))
    keep = np.sort(R.edge_dropout_keep(perm.numpy(), rate))    # compaction keeps base order: ties by base edge id
    flags = o.keep_flags(e, [(perm.to(dev), len(keep), 0)], dev)
    got = o.csr_dropout(base, flags, len(keep))
    for g, (r_, c_) in ((got, (row, col)), (got.transpose(), (col, row))):
        indptr, indices, eid = R.csr_from_pairs(r_[keep], c_[keep], g.n_rows)
        np.testing.assert_array_equal(g.indptr.cpu().numpy(), indptr)
        np.testing.assert_array_equal(g.indices.cpu().numpy(), indices)
        np.testing.assert_array_equal(g.eid.cpu().numpy(), keep[eid])          # ids still name base edges
        np.testing.assert_array_equal(g.vals.cpu().numpy(), val[keep][eid])


# ---- SpMM ------------------------------------------------------------------------------------------
def _dense_spmm(row, col, val, n_rows, x, ss, ds, bias, relu):
    x64 = x.double() * (ss.double()[:, None] if ss is not None else 1.0)
    out = th.zeros(n_rows, x.shape[1], dtype=th.float64)
    w = th.tensor(val, dtype=th.float64)[:, None] if val is not None else 1.0
    out.index_add_(0, th.tensor(row), x64[th.tensor(col)] * w)
    if ds is not None:
        out = out * ds.double()[:, None]
    if bias is not None:
        out = out + bias.double()
    return out.relu() if relu else out


@pytest.mark.parametrize('d', [4, 36, 128, 256, 344, 512, 768])
@pytest.mark.parametrize('weighted,scaled,epi', [(False, False, False), (False, True, False), (True, False, True),
                                                  (True, True, True)])
def test_spmm_forward_backward(dev, d, weighted, scaled, epi):
    rng = np.random.default_rng(d)
    n_rows, n_cols, e = 70, 90, 2500
    row, col = rng.integers(0, n_rows, e), rng.integers(0, n_cols, e)
    row[row == 3] = 4                                            # an empty row
    val = rng.random(e).astype(np.float32) if weighted else None
    g = th.Generator().manual_seed(d)
    x = th.randn(n_cols, d, generator=g)
    ss = th.rand(n_cols, generator=g) if scaled else None
    ds = th.rand(n_rows, generator=g) if scaled else None
    bias = th.randn(d, generator=g) if epi else None
    o = ops()
    csr = o.CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n_rows, n_cols,
                         th.tensor(val, device=dev) if weighted else None)
    xg = x.to(dev).requires_grad_(True)
    bg = bias.to(dev).requires_grad_(True) if epi else None
    out = o.spmm(csr, xg, ss.to(dev) if scaled else None, ds.to(dev) if scaled else None, bg, relu=epi)
    want = _dense_spmm(row, col, val, n_rows, x, ss, ds, bias, epi)
    assert H.rel_err(out.detach().cpu(), want) <= FP32_TOL
    gout = th.randn(n_rows, d, generator=g)
    out.backward(gout.to(dev))
    xr = x.clone().double().requires_grad_(True)
    br = bias.clone().double().requires_grad_(True) if epi else None
    ref = _dense_spmm(row, col, val, n_rows, xr, ss, ds, br, epi)
    ref.backward(gout.double())
    assert H.rel_err(xg.grad.cpu(), xr.grad) <= FP32_TOL
    if epi:
        assert H.rel_err(bg.grad.cpu(), br.grad) <= FP32_TOL
    # atomic-free claim: a second run is bit-identical
    out2 = o.spmm(csr, xg.detach(), ss.to(dev) if scaled else None, ds.to(dev) if scaled else None,
                  bg.detach() if epi else None, relu=epi)
    assert th.equal(out2, out.detach())


@pytest.mark.parametrize('d', [8, 128, 344, 768])
def test_spmm_bf16_storage(dev, d):
    rng = np.random.default_rng(d)
    n_rows, n_cols, e = 64, 80, 3000
    row, col = rng.integers(0, n_rows, e), rng.integers(0, n_cols, e)
    x = th.randn(n_cols, d, generator=th.Generator().manual_seed(1))
    ss = th.rand(n_cols, generator=th.Generator().manual_seed(2))
    o = ops()
    csr = o.CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n_rows, n_cols)
    out = o.spmm(csr, x.to(dev).to(th.bfloat16), ss.to(dev), None)
    assert out.dtype == th.float32
    assert H.rel_err(out.cpu(), _dense_spmm(row, col, None, n_rows, x, ss, None, None, False)) <= BF16_TOL
    exact = _dense_spmm(row, col, None, n_rows, x.to(th.bfloat16).float(), ss, None, None, False)
    assert H.rel_err(out.cpu(), exact) <= FP32_TOL               # only the storage rounding differs


def test_spmm_rejects_bad_layout(dev):
    o = ops()
    csr = o.CSR.from_coo(th.tensor([0], device=dev), th.tensor([0], device=dev), 1, 1)
    with pytest.raises(RuntimeError, match='multiples'):
        o.spmm(csr, th.randn(1, 5, device=dev))


@pytest.mark.parametrize('d', [128, 344])
@pytest.mark.parametrize('weighted,scaled,epi', [(False, True, False), (True, True, True)])
def test_spmm_chunked_aggregation_of_skewed_rows(dev, d, weighted, scaled, epi):
    """A graph with a 5 000-edge row and a 4 500-edge column (Zipf stress set): from_coo flags both orientations
    for chunked aggregation (ops.split_plan: chunk partials, then per-row sums with the epilogue); forward and backward
    equal the dense float64 restatement, equal the one-launch path up to summation order, and are deterministic."""
    rng = np.random.default_rng(d + weighted)
    n_rows, n_cols = 6000, 5000
    row = np.concatenate([np.full(n_cols, 7), rng.choice(n_rows, 4500, replace=False), rng.integers(0, n_rows, 20000)])
    col = np.concatenate([np.arange(n_cols), np.full(4500, 11), rng.integers(0, n_cols, 20000)])
    e = len(row)
    val = rng.random(e).astype(np.float32) if weighted else None
    g = th.Generator().manual_seed(d)
    x = th.randn(n_cols, d, generator=g)
    ss, ds = (th.rand(n_cols, generator=g), th.rand(n_rows, generator=g)) if scaled else (None, None)
    bias = th.randn(d, generator=g) if epi else None
    o = ops()
    mk = lambda: o.CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n_rows, n_cols,
                                th.tensor(val, device=dev) if weighted else None)
    csr = mk()
    assert csr.split_T == o.SPMM_SPLIT_T and csr.transpose().split_T == o.SPMM_SPLIT_T
    to = lambda t: None if t is None else t.to(dev)
    xg = x.to(dev).requires_grad_(True)
    bg = bias.to(dev).requires_grad_(True) if epi else None
    out = o.spmm(csr, xg, to(ss), to(ds), bg, relu=epi)
    assert csr._plan is not None and csr._plan[0].n_rows == n_rows + e // o.SPMM_SPLIT_T
    want = _dense_spmm(row, col, val, n_rows, x, ss, ds, bias, epi)
    assert H.rel_err(out.detach().cpu(), want) <= FP32_TOL
    gout = th.randn(n_rows, d, generator=g)
    out.backward(gout.to(dev))
    xr = x.clone().double().requires_grad_(True)
    br = bias.clone().double().requires_grad_(True) if epi else None
    _dense_spmm(row, col, val, n_rows, xr, ss, ds, br, epi).backward(gout.double())
    assert H.rel_err(xg.grad.cpu(), xr.grad) <= FP32_TOL
    if epi:
        assert H.rel_err(bg.grad.cpu(), br.grad) <= FP32_TOL
    assert th.equal(o.spmm(csr, xg.detach(), to(ss), to(ds), None if bg is None else bg.detach(), relu=epi), out.detach())
    # the one-launch path on the same graph: same sums in another order
    plain = mk()
    plain.split_T = 0
    plain.transpose().split_T = 0
    one = o.spmm(plain, xg.detach(), to(ss), to(ds), None if bg is None else bg.detach(), relu=epi)
    assert plain._plan is None and H.rel_err(out.detach().cpu(), one.cpu()) <= 5e-6
    # a dropout compaction keeps the decision in both orientations
    flags = th.ones(e, dtype=th.uint8, device=dev)
    dropped = o.csr_dropout(csr, flags, e)
    assert dropped.split_T == o.SPMM_SPLIT_T and dropped.transpose().split_T == o.SPMM_SPLIT_T
    assert th.equal(o.spmm(dropped, xg.detach(), to(ss), to(ds), None if bg is None else bg.detach(), relu=epi), out.detach())


# ---- decoder ---------------------------------------------------------------------------------------
def _decoder_params(gen, n_in):
    P = {'decoder.lin1.weight': th.randn(128, 2 * n_in, generator=gen) * 0.2,
         'decoder.lin1.bias': th.randn(128, generator=gen) * 0.1,
         'decoder.lin2.weight': th.randn(64, 128, generator=gen) * 0.2,
         'decoder.lin2.bias': th.randn(64, generator=gen) * 0.1,
         'decoder.lin3.weight': th.randn(1, 64, generator=gen) * 0.2,
         'decoder.lin3.bias': th.randn(1, generator=gen) * 0.1}
    return P


@pytest.mark.parametrize('impl,order', [('tc', 'label'), ('tc', 'by-drug'), ('simt', 'label')])
@pytest.mark.parametrize('n_pairs', [0, 1, 63, 128, 1000, 40000])
def test_decoder_forward_backward(dev, n_pairs, impl, order, monkeypatch):
    """impl: the tcgen05 kernels (default) or the fp32 FMA ones; order: label order or the by-drug walk."""
    monkeypatch.setenv('DG_DECODER', impl)
    monkeypatch.setenv('DG_DECODER_ORDER', order)
    from dreamgnn_b200.layers import MLPDecoder
    from dreamgnn_b200 import graph as G
    gen = th.Generator().manual_seed(n_pairs)
    n_d, n_s, n_in = 37, 29, 16
    rng = np.random.default_rng(n_pairs)
    src, dst = rng.integers(0, n_d, n_pairs), rng.integers(0, n_s, n_pairs)
    hd, hs = th.randn(n_d, n_in, generator=gen), th.randn(n_s, n_in, generator=gen)
    P = _decoder_params(gen, n_in)
    dec = MLPDecoder(n_in, dropout_rate=0.3)
    dec.load_state_dict({k[len('decoder.'):]: v for k, v in P.items()})
    dec = dec.to(dev).eval()
    g = G.heterograph({('drug', 'rate', 'disease'): (src, dst)}, {'drug': n_d, 'disease': n_s}).int().to(dev)
    hdg, hsg = hd.to(dev).requires_grad_(True), hs.to(dev).requires_grad_(True)
    out = dec(g, hdg, hsg)
    assert out.shape == (n_pairs, 1)
    Pr = {k: v.clone().double().requires_grad_(True) for k, v in P.items()}
    hdr, hsr = hd.double().requires_grad_(True), hs.double().requires_grad_(True)
    ref = R.mlp_decoder(Pr, 'decoder.', src, dst, hdr, hsr)
    assert H.rel_err(out.detach().cpu(), ref.detach()) <= FP32_TOL
    if n_pairs == 0:
        return
    gout = th.randn(n_pairs, 1, generator=gen)
    out.backward(gout.to(dev))
    ref.backward(gout.double())
    assert H.rel_err(hdg.grad.cpu(), hdr.grad) <= FP32_TOL
    assert H.rel_err(hsg.grad.cpu(), hsr.grad) <= FP32_TOL
    for k, p in dec.named_parameters():
        assert H.rel_err(p.grad.cpu(), Pr['decoder.' + k].grad) <= FP32_TOL, k
    # deterministic (atomic-free) backward
    g1 = hdg.grad.clone()
    hdg.grad = None
    dec.zero_grad()
    dec(g, hdg, hsg).backward(gout.to(dev))
    assert th.equal(hdg.grad, g1)


@pytest.mark.parametrize('n_pairs,sorted_src', [(5000, True), (5000, False), (17, True), (40001, True)])
def test_decoder_fused_source_segment_sum(dev, n_pairs, sorted_src, monkeypatch):
    """The segment sum by source node fused into the backward's epilogue (one partial row per run of equal source inside
    aligned 16-pair groups, then a sum over slots) against the plain segment-sum SpMM over dz1: same gradient to fp32
    reassociation, and the slot plan has one slot per 16 pairs plus one per extra run."""
    o = ops()
    gen = th.Generator().manual_seed(n_pairs)
    rng = np.random.default_rng(n_pairs)
    n_d, n_s = 41, 33
    src, dst = rng.integers(0, n_d, n_pairs), rng.integers(0, n_s, n_pairs)
    if sorted_src:
        src = np.sort(src)
    pairs = o.PairGraph(th.tensor(src, device=dev), th.tensor(dst, device=dev), n_d, n_s)
    slot, n_slots, seg = pairs.source_slots()
    runs = 1 + int(((np.arange(1, n_pairs) % 16 == 0) | (src[1:] != src[:-1])).sum())
    assert n_slots == runs and slot.numel() == n_pairs and int(slot[-1]) == n_slots - 1
    mk = lambda *sh: (th.randn(*sh, generator=gen) * 0.3).to(dev)
    ps, w2, b2, w3, b3 = mk(n_s, 128), mk(64, 128), mk(64), mk(1, 64), mk(1)
    pd0 = mk(n_d, 128)
    gout = th.randn(n_pairs, 1, generator=gen).to(dev)
    grads = {}
    for mode in ('fused', 'spmm'):
        monkeypatch.setenv('DG_DECODER_SEG', mode)
        pd = pd0.clone().requires_grad_(True)
        o.decoder_mlp(pd, ps, w2, b2, w3, b3, pairs, training=True).backward(gout)
        grads[mode] = pd.grad.clone()
    assert H.rel_err(grads['fused'].cpu(), grads['spmm'].cpu()) <= 2e-6


@pytest.mark.parametrize('order', ['label', 'by-drug'])
def test_decoder_dropout_is_consistent_between_forward_and_backward(dev, order, monkeypatch):
    """Training mode: the backward regenerates the forward's masks. Checked as a directional derivative of
    the (piecewise-linear, fixed-seed) function, plus the keep rate and the 1/(1-p) scaling."""
    monkeypatch.setenv('DG_DECODER_ORDER', order)
    o = ops()
    gen = th.Generator().manual_seed(7)
    n_d, n_s, e, p = 50, 40, 6000, 0.3
    rng = np.random.default_rng(7)
    pairs = o.PairGraph(th.tensor(rng.integers(0, n_d, e), device=dev), th.tensor(rng.integers(0, n_s, e), device=dev),
                        n_d, n_s)
    mk = lambda *s: (th.randn(*s, generator=gen) * 0.3).to(dev)
    pd, ps, w2, b2, w3, b3 = mk(n_d, 128), mk(n_s, 128), mk(64, 128), mk(64), mk(1, 64), mk(1)

    def f(pd_, w2_):
        return o.decoder_mlp(pd_, ps, w2_, b2, w3, b3, pairs, p=p, seed=1234, training=True)
    pd_g, w2_g = pd.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    out = f(pd_g, w2_g)
    assert th.equal(out.detach(), f(pd, w2))                      # same seed -> same masks
    assert not th.equal(out.detach(), o.decoder_mlp(pd, ps, w2, b2, w3, b3, pairs, p=p, seed=99, training=True))
    gout = th.randn(e, 1, generator=gen).to(dev)
    out.backward(gout)
    dpd, dw2 = mk(n_d, 128), mk(64, 128)
    eps = 1e-3
    fd = ((f(pd + eps * dpd, w2 + eps * dw2).double() - f(pd - eps * dpd, w2 - eps * dw2).double()) * gout).sum() / (2 * eps)
    an = (pd_g.grad.double() * dpd).sum() + (w2_g.grad.double() * dw2).sum()
    # the directional derivative can be small next to the gradient (cancellation): scale the tolerance by the
    # Cauchy-Schwarz bound of the two inner products, not by |an|
    scale = float(pd_g.grad.double().norm() * dpd.double().norm() + w2_g.grad.double().norm() * dw2.double().norm())
    assert abs(float(fd - an)) <= 5e-3 * scale
    # keep-rate of the first dropout: eval/(train) ratio on a linear probe
    z2 = th.empty(e, 64, device=dev)
    lib = o.L.load()
    outp = th.empty(e, device=dev)
    big = th.full((n_d, 128), 5.0, device=dev)
    # keep every operand alive until the launch is enqueued (temporaries would be recycled by the allocator)
    zs, w_ones, b_zero, w3_ones, b3_zero = (th.zeros_like(ps), th.ones(64, 128, device=dev), th.zeros(64, device=dev),
                                            th.ones(64, device=dev), th.zeros(1, device=dev))
    o.L.check(lib.dg_decoder_fwd_f32(o.L.ptr(pairs.src), o.L.ptr(pairs.dst), None, e, o.L.ptr(big), o.L.ptr(zs), o.L.ptr(w_ones),
                                     o.L.ptr(b_zero), o.L.ptr(w3_ones), o.L.ptr(b3_zero), p, 42, None, o.L.ptr(outp), o.L.ptr(z2),
                                     o.L.stream()), 'decoder_fwd')
    th.cuda.synchronize()
    kept2 = float((z2 > 0).float().mean())
    assert abs(kept2 - (1 - p)) < 0.01
    # every z2 entry = (number of kept z1 units) * 5/(1-p) /(1-p) when kept: mean over kept ~ 128*5/(1-p)
    mean_kept = float(z2[z2 > 0].mean())
    assert abs(mean_kept - 128 * 5.0 / (1 - p)) / (128 * 5.0 / (1 - p)) < 0.01


# ---- kNN --------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,k', [(8, 3), (60, 4), (333, 15), (1000, 33), (200, 64)])
def test_topk_rows(dev, n, k):
    rng = np.random.default_rng(n)
    sim = rng.random((n, n + 7))
    sim[:, 5] = sim[:, 2]                                         # exact ties -> (value desc, index asc)
    got = ops().topk_rows(th.tensor(sim, device=dev), k).cpu().numpy()
    order = np.lexsort((np.broadcast_to(np.arange(n + 7), sim.shape), -sim), axis=1)[:, :k]
    np.testing.assert_array_equal(got, np.sort(order, axis=1))


@pytest.mark.parametrize('n,k', [(5, 2), (60, 4), (500, 15), (3000, 40)])
def test_knn_graph_from_neighbors_bit_exact(dev, n, k):
    rng = np.random.default_rng(n + k)
    nbr = np.stack([rng.choice(n, size=k, replace=False) for _ in range(n)])
    for i in range(n // 2):                                       # self among the neighbours for half the rows
        if i not in nbr[i]:
            nbr[i, 0] = i
    nbr = np.sort(nbr, axis=1)
    csr, rows = ops().knn_graph_from_neighbors(th.tensor(nbr, device=dev))
    row, col, val = R.knn_graph_from_neighbors(nbr, n)
    np.testing.assert_array_equal(rows.cpu().numpy(), row)
    np.testing.assert_array_equal(csr.indices.cpu().numpy(), col)
    np.testing.assert_array_equal(csr.vals.cpu().numpy(), val)     # bit-exact fp32
    np.testing.assert_array_equal(csr.indptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(row, minlength=n))]))


def test_launch_counter_counts_kernels(dev):
    from dreamgnn_b200 import _lib
    _lib.reset_launch_count()
    ops().exclusive_scan_i32(th.ones(10000, dtype=th.int32, device=dev))
    assert _lib.launch_count() == 3


# ---- tcgen05 projection GEMM -----------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K,R', [(1, 1, 1, 1), (128, 128, 32, 1), (130, 70, 36, 1), (257, 129, 100, 3),
                                    (1000, 344, 1024, 2), (344, 96, 20000, 2)])
def test_gemm_nt_3xtf32(dev, M, N, K, R):
    gen = th.Generator().manual_seed(M + N + K)
    a = th.randn(M, K, generator=gen)
    b = th.randn(R, N, K, generator=gen) if R > 1 else th.randn(N, K, generator=gen)
    rs = th.rand(R * M, generator=gen)
    ref = (a.double() @ b.double().transpose(-1, -2)) * rs.double().view(R, M, 1).squeeze(0) if R > 1 else \
        (a.double() @ b.double().t()) * rs.double().view(M, 1)
    got = ops().gemm_nt(a.to(dev), b.to(dev), row_scale=rs.to(dev), precision=0)
    assert H.rel_err(got.cpu(), ref) <= FP32_TOL
    assert th.equal(got, ops().gemm_nt(a.to(dev), b.to(dev), row_scale=rs.to(dev), precision=0))    # deterministic split-K
    tf32 = ops().gemm_nt(a.to(dev), b.to(dev), row_scale=rs.to(dev), precision=1)
    assert H.rel_err(tf32.cpu(), ref) <= 2e-3


@pytest.mark.parametrize('ta,tb', [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize('M,N,K,R', [(1, 1, 1, 1), (128, 128, 32, 1), (130, 70, 36, 1), (257, 129, 100, 3), (96, 344, 5000, 2),
                                    (1024, 341, 777, 2)])
def test_gemm_transposed_operands(dev, M, N, K, R, ta, tb):
    """op(A) [M,K] / op(B) [N,K] given as [K,M] / [K,N] row-major: read MN-major by the tensor cores (and through the
    packing copy when a row stride is not a multiple of 16 bytes), same result as the explicit transposes."""
    gen = th.Generator().manual_seed(M + N + K)
    a = th.randn(M, K, generator=gen)
    b = th.randn(R, N, K, generator=gen) if R > 1 else th.randn(N, K, generator=gen)
    ref = a.double() @ b.double().transpose(-1, -2)
    a_in = a.t().contiguous() if ta else a
    b_in = b.transpose(-1, -2).contiguous() if tb else b
    got = ops().gemm(a_in.to(dev), b_in.to(dev), trans_a=ta, trans_b=tb)
    assert tuple(got.shape) == tuple(ref.shape)
    assert H.rel_err(got.cpu(), ref) <= FP32_TOL
    assert th.equal(got, ops().gemm(a_in.to(dev), b_in.to(dev), trans_a=ta, trans_b=tb))


@pytest.mark.parametrize('M,K,N,bias', [(700, 341, 128, True), (513, 256, 128, True), (300, 128, 128, False)])
def test_linear_autograd_on_tensor_cores(dev, M, K, N, bias):
    """nn.Linear forward / backward through dg_gemm_f32 (unaligned 341-wide rows go through the packing copy)."""
    o = ops()
    gen = th.Generator().manual_seed(M + K)
    x, w, b = th.randn(M, K, generator=gen), th.randn(N, K, generator=gen) * 0.1, th.randn(N, generator=gen)
    gout = th.randn(M, N, generator=gen)
    old = o.GEMM_MIN_MACS
    o.GEMM_MIN_MACS = 0
    try:
        xg, wg, bg = (t.to(dev).requires_grad_(True) for t in (x, w, b))
        y = o.linear(xg, wg, bg if bias else None)
        y.backward(gout.to(dev))
    finally:
        o.GEMM_MIN_MACS = old
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    yr = th.nn.functional.linear(xr, wr, br if bias else None)
    yr.backward(gout.double())
    assert H.rel_err(y.detach().cpu(), yr.detach()) <= FP32_TOL
    assert H.rel_err(xg.grad.cpu(), xr.grad) <= FP32_TOL and H.rel_err(wg.grad.cpu(), wr.grad) <= FP32_TOL
    if bias:
        assert H.rel_err(bg.grad.cpu(), br.grad) <= FP32_TOL


def test_project_autograd_on_tensor_cores(dev):
    o = ops()
    gen = th.Generator().manual_seed(3)
    x = th.randn(700, 96, generator=gen)
    w = th.randn(2, 96, 40, generator=gen)
    gout = th.randn(2, 700, 40, generator=gen)
    old = o.GEMM_MIN_MACS
    o.GEMM_MIN_MACS = 0
    try:
        xg, wg = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
        y = o.project(xg, wg)
        y.backward(gout.to(dev))
    finally:
        o.GEMM_MIN_MACS = old
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = th.matmul(xr.unsqueeze(0), wr)
    yr.backward(gout.double())
    assert H.rel_err(y.detach().cpu(), yr.detach()) <= FP32_TOL
    assert H.rel_err(xg.grad.cpu(), xr.grad) <= FP32_TOL and H.rel_err(wg.grad.cpu(), wr.grad) <= FP32_TOL


# ---- row-streaming helpers (rowops.cu) ------------------------------------------------------------------
@pytest.mark.parametrize('n,d', [(0, 8), (1, 4), (7, 128), (1000, 344), (5003, 768), (100001, 128), (300, 1024)])
@pytest.mark.parametrize('gated', [False, True])
def test_colsum_is_deterministic_and_exact(dev, n, d, gated):
    gen = th.Generator().manual_seed(n + d)
    x = th.randn(n, d, generator=gen)
    gate = th.randn(n, d, generator=gen) if gated else None
    xs = x.to(dev)
    gs = gate.to(dev) if gated else None
    masked, s = ops().colsum(xs, gate=gs, want_masked=True)
    want_m = x.double() * (gate > 0).double() if gated else x.double()
    assert th.equal(masked.cpu().double(), want_m)                           # the mask is exact
    want = want_m.sum(0)
    err = (s.cpu().double() - want).abs().max().item() if n else 0.0
    assert err <= 1e-6 * max(1.0, want_m.abs().sum(0).max().item() if n else 1.0)
    s2 = ops().colsum(xs, gate=gs)
    assert th.equal(s, s2)                                                    # fixed summation order: bit-identical
    # a strided view (rows 16-byte aligned) goes through the kernel as it is, an unaligned one through an aligned copy
    if d >= 8 and n:
        v = xs[:, 4:d]
        np.testing.assert_allclose(ops().colsum(v).cpu().numpy(), x[:, 4:d].double().sum(0).numpy(), rtol=0, atol=1e-3)
        u = xs[:, 1:d]
        np.testing.assert_allclose(ops().colsum(u).cpu().numpy(), x[:, 1:d].double().sum(0).numpy(), rtol=0, atol=1e-3)


@pytest.mark.parametrize('n,d', [(3, 4), (50, 16), (777, 128), (20000, 128), (3000, 64)])
def test_gram_common_loss_matches_reference_form(dev, n, d):
    """utils.py:87-95 (N x N form) on CPU in float64 vs the explicit-kernel Gram form: value and both gradients."""
    from dreamgnn_b200.utils import common_loss, common_loss_gram, common_loss_gram_torch
    gen = th.Generator().manual_seed(n * 31 + d)
    a = th.randn(n, d, generator=gen) + 0.3
    b = a * 0.5 + th.randn(n, d, generator=gen)
    ref_in = [t.double().requires_grad_(True) for t in (a, b)]
    ref = common_loss(*ref_in)
    ref.backward()
    got_in = [t.to(dev).requires_grad_(True) for t in (a, b)]
    got = common_loss_gram(*got_in)
    assert got.dtype == th.float32
    (got * 3.0).backward()
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-12
    for g, r in zip(got_in, ref_in):
        assert H.rel_err(g.grad.cpu().double() / 3.0, r.grad) <= FP32_TOL
    # and the traced torch expression it replaces
    tor_in = [t.to(dev).requires_grad_(True) for t in (a, b)]
    tor = common_loss_gram_torch(*tor_in)
    tor.backward()
    assert abs(got.item() - tor.item()) <= 1e-5 * abs(tor.item()) + 1e-12
    for g, t in zip(got_in, tor_in):
        assert H.rel_err(g.grad / 3.0, t.grad) <= FP32_TOL


def test_gram_common_loss_one_sided_gradient(dev):
    from dreamgnn_b200.utils import common_loss_gram, common_loss_gram_torch
    gen = th.Generator().manual_seed(3)
    a, b = th.randn(500, 32, generator=gen).to(dev), th.randn(500, 32, generator=gen).to(dev)
    a1, a2 = a.clone().requires_grad_(True), a.clone().requires_grad_(True)
    common_loss_gram(a1, b).backward()
    common_loss_gram_torch(a2, b).backward()
    assert H.rel_err(a1.grad, a2.grad) <= FP32_TOL


def test_spmm_relu_bias_backward_fused(dev):
    """GraphConvolution epilogue (layers.py:311-314 + F.relu): bias gradient and ReLU mask from one pass."""
    rng = np.random.default_rng(11)
    n, d, nnz = 400, 64, 3000
    row, col = rng.integers(0, n, nnz), rng.integers(0, n, nnz)
    vals = rng.random(nnz).astype(np.float32)
    csr = ops().CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n, n, th.tensor(vals, device=dev))
    x = th.randn(n, d, device=dev, requires_grad=True)
    bias = th.randn(d, device=dev, requires_grad=True)
    out = ops().spmm(csr, x, bias=bias, relu=True)
    w = th.randn(n, d, device=dev)
    (out * w).sum().backward()
    A = th.zeros(n, n, dtype=th.float64)
    A.index_put_((th.tensor(row), th.tensor(col)), th.tensor(vals).double(), accumulate=True)
    xr, br = x.detach().cpu().double().requires_grad_(True), bias.detach().cpu().double().requires_grad_(True)
    ref = th.relu(A @ xr + br)
    (ref * w.cpu().double()).sum().backward()
    assert H.rel_err(out.detach().cpu().double(), ref.detach()) <= FP32_TOL
    assert H.rel_err(x.grad.cpu().double(), xr.grad) <= FP32_TOL
    assert H.rel_err(bias.grad.cpu().double(), br.grad) <= FP32_TOL


@pytest.mark.parametrize('d', [4, 128, 344, 768])
@pytest.mark.parametrize('weighted', [False, True])
def test_spmm_l2_prefetch_instance_is_bit_identical(dev, d, weighted, monkeypatch):
    """DG_SPMM_PREFETCH only adds prefetch.global.L2 hints one group ahead of the demand loads (rows of the current
    batch, of the next batch, start of a row): same sums in the same order, on rows of every length class."""
    o = ops()
    rng = np.random.default_rng(d + 7)
    lengths = [0, 1, 3, 4, 5, 31, 32, 33, 36, 63, 64, 65, 100, 257, 1000]
    n_rows, n_cols = len(lengths), 300
    row = np.concatenate([np.full(n, i) for i, n in enumerate(lengths)])
    col = rng.integers(0, n_cols, row.size)
    val = th.tensor(rng.random(row.size).astype(np.float32), device=dev) if weighted else None
    csr = o.CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n_rows, n_cols, val)
    x = th.randn(n_cols, d, device=dev)
    ss = th.rand(n_cols, device=dev) if weighted else None
    outs = {}
    for on in (False, True):
        monkeypatch.setattr(o, 'SPMM_PREFETCH_MIN_BYTES', 0 if on else 1 << 60)
        monkeypatch.setattr(o, 'SPMM_PREFETCH_MIN_BYTES_WIDE', 0 if on else 1 << 60)
        outs[on] = o._spmm_raw(csr, x, ss, None, None, 0)
        outs[(on, 'bf16')] = o._spmm_raw(csr, x.to(th.bfloat16), ss, None, None, 0) if d % 8 == 0 else None
    assert th.equal(outs[False], outs[True])
    if d % 8 == 0:
        assert th.equal(outs[(False, 'bf16')], outs[(True, 'bf16')])
    want = _dense_spmm(row, col, None if val is None else val.cpu().numpy(), n_rows, x.cpu(), None if ss is None else ss.cpu(),
                       None, None, False)
    assert H.rel_err(outs[True].cpu(), want) <= FP32_TOL


@pytest.mark.parametrize('d', [4, 128, 344, 768])
@pytest.mark.parametrize('weighted', [False, True])
def test_spmm_rowsplit_instance(dev, d, weighted, monkeypatch):
    """DG_SPMM_ROWSPLIT (few, long rows: CTA per row, 8 warps summed in warp order): parity with the dense float64
    product, bit-reproducible, forward and backward, on rows of every length class (empty rows, < 8 batches, ragged)."""
    o = ops()
    rng = np.random.default_rng(d + 11)
    lengths = [0, 1, 31, 32, 33, 255, 256, 257, 300, 1000, 1537, 5]
    n_rows, n_cols = len(lengths), 400
    row = np.concatenate([np.full(n, i) for i, n in enumerate(lengths)])
    col = rng.integers(0, n_cols, row.size)
    val = rng.random(row.size).astype(np.float32) if weighted else None
    csr = o.CSR.from_coo(th.tensor(row, device=dev), th.tensor(col, device=dev), n_rows, n_cols,
                         th.tensor(val, device=dev) if weighted else None)
    gen = th.Generator().manual_seed(d)
    x = th.randn(n_cols, d, generator=gen)
    ss, ds, bias = th.rand(n_cols, generator=gen), th.rand(n_rows, generator=gen), th.randn(d, generator=gen)
    monkeypatch.setattr(o, 'SPMM_ROWSPLIT_MIN_AVG_LEN', 0)
    assert csr.n_rows <= o.SPMM_ROWSPLIT_MAX_ROWS
    n0 = o.L.launch_count()
    xg = x.to(dev).requires_grad_(True)
    bg = bias.to(dev).requires_grad_(True)
    out = o.spmm(csr, xg, ss.to(dev), ds.to(dev), bg, relu=True)
    want = _dense_spmm(row, col, val, n_rows, x, ss, ds, bias, True)
    assert H.rel_err(out.detach().cpu(), want) <= FP32_TOL
    gout = th.randn(n_rows, d, generator=gen)
    out.backward(gout.to(dev))
    xr, br = x.clone().double().requires_grad_(True), bias.clone().double().requires_grad_(True)
    _dense_spmm(row, col, val, n_rows, xr, ss, ds, br, True).backward(gout.double())
    assert H.rel_err(xg.grad.cpu(), xr.grad) <= FP32_TOL
    assert H.rel_err(bg.grad.cpu(), br.grad) <= FP32_TOL
    out2 = o.spmm(csr, xg.detach(), ss.to(dev), ds.to(dev), bg.detach(), relu=True)
    assert th.equal(out2, out.detach())
    assert o.L.launch_count() > n0
    # and against the warp-per-row kernel: same sums up to the rounding order
    monkeypatch.setattr(o, 'SPMM_ROWSPLIT_MAX_ROWS', 0)
    out3 = o.spmm(csr, xg.detach(), ss.to(dev), ds.to(dev), bg.detach(), relu=True)
    assert H.rel_err(out3.cpu(), out.detach().cpu()) <= 5e-6


# ---- sort-free edge-dropout sampler (select.cu) -----------------------------------------------------------
@pytest.mark.parametrize('n', [1, 2, 5, 257, 3001, 16384, 16385, 100003, 3_000_000])      # <= 16 384: the one-launch instance
def test_random_subset_flags_is_the_k_smallest_keys(dev, n):
    """dg_random_subset_flags keeps exactly the num_keep smallest of the keys (rnd << bits(n-1)) | i -- the set a full
    sort of those keys would put first (what randperm[:num_keep] does with torch's keys)."""
    o = ops()
    rng = np.random.default_rng(n)
    rnd = rng.integers(0, 2 ** 63 - 1, size=n, dtype=np.int64)
    if n > 100:
        rnd[::7] = rnd[3]                                   # many equal random parts: the index bits break the ties
    bits = 0
    while (1 << bits) < n:
        bits += 1
    key = (rnd.astype(np.uint64) << np.uint64(bits)) | np.arange(n, dtype=np.uint64)
    order = np.argsort(key, kind='stable')
    for k in sorted({0, 1, n // 3, n - 1, n} & set(range(n + 1))):
        flags = th.zeros(n, dtype=th.uint8, device=dev)
        o.random_subset_flags(n, k, flags, rnd=th.tensor(rnd, device=dev))
        want = np.zeros(n, dtype=np.uint8)
        want[order[:k]] = 1
        np.testing.assert_array_equal(flags.cpu().numpy(), want)


def test_random_subset_sampler_is_uniform_and_fresh(dev):
    o = ops()
    th.manual_seed(123)
    n, k, draws = 4000, 1000, 400
    acc = th.zeros(n, device=dev)
    prev = None
    for _ in range(draws):
        f = o.keep_flags(n, [(o.RandomSubset(n), k, 0)], dev)
        assert int(f.sum()) == k
        assert prev is None or not th.equal(prev, f)
        prev = f
        acc += f.float()
    freq = (acc / draws).cpu().numpy()
    # every edge is kept with probability k/n = 0.25: binomial std 0.0217 per edge over 400 draws
    assert abs(freq.mean() - 0.25) < 1e-9 and freq.std() < 0.03 and freq.min() > 0.12 and freq.max() < 0.38
    # position-independent: first and second half of the index range are kept equally often
    assert abs(freq[: n // 2].mean() - freq[n // 2:].mean()) < 0.005


def test_edge_dropout_with_select_sampler(dev, monkeypatch):
    """augment_graph_data with DG_EDGE_SAMPLER=select: same structure invariants as the randperm path (exact kept
    counts per relation, forward / transposed blocks consistent, kept edges are a subset of the base graph)."""
    from dreamgnn_b200 import graph_build as GB
    from dreamgnn_b200.augmentation import GraphAugmentation, num_keep_edges
    monkeypatch.setenv('DG_EDGE_SAMPLER', 'select')
    rng = np.random.default_rng(4)
    n_d, n_s, e = 300, 200, 20000
    cells = rng.choice(n_d * n_s, size=e, replace=False)
    pairs = (cells // n_s, cells % n_s)
    labels = (rng.random(e) < 0.1).astype(np.float32)
    g = GB.generate_enc_graph(pairs, labels, n_d, n_s, dev).int()
    out = GraphAugmentation.random_edge_dropout(g, 0.1)
    for c in g.canonical_etypes:
        assert out.number_of_edges(c) == num_keep_edges(g.number_of_edges(c), 0.1)
        s0, d0 = g.edges(etype=c)
        s1, d1 = out.edges(etype=c)
        base = set(zip(s0.cpu().tolist(), d0.cpu().tolist()))
        kept = list(zip(s1.cpu().tolist(), d1.cpu().tolist()))
        assert len(kept) == len(set(kept)) == out.number_of_edges(c) and set(kept) <= base
    for dt in ('drug', 'disease'):
        blk = out.block(dt)
        t = blk.csr.transpose()
        assert blk.csr.nnz == t.nnz == int(blk.csr.indptr[-1]) == int(t.indptr[-1])
        assert th.equal(th.sort(blk.csr.eid).values, th.sort(t.eid).values)
