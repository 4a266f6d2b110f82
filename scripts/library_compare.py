"""The hot kernels against the library kernels a stock PyTorch / DGL GPU path would run, on the same box and the same
inputs (SURVEY.md 8d "GPU comparison kernels"), at the syn20m shape (run on the GPU box):

    python scripts/library_compare.py [--out gpurun_out/library_compare.jsonl] [--pairs 20000000]

  * SpMM  (GCMC relation block, d = 344 / 128, both orientations): `dg_spmm_csr_f32` with `cj[src]` / `ci[dst]` fused
    vs `torch.sparse.mm` on a CSR tensor (cuSPARSE SpMM; what DGL's GPU `update_all(copy_u, sum)` calls) -- alone, and with
    the two scale passes the library path needs -- and vs gather + `index_add_` (the stand-in's / a naive port's form);
  * decoder (all scored pairs): `dg_decoder_fwd/bwd_f32` vs gather + `F.linear` in the split-`lin1` form (already an
    optimisation over layers.py:360-375, which concatenates [E, 256]) under autograd;
  * projection GEMM [100k, 1024] x [1024, 344]: `dg_gemm_f32` (3xTF32 on tcgen05) vs `torch.mm` in strict fp32 and with
    TF32 allowed;
  * the same SpMMs on the Zipf(1.0) stress set (`synthetic.zipf_cells`: drug-side row lengths 1 ... 50 000).

Every result is one JSON line, written and flushed as soon as it exists (a run cut short keeps what it measured). Times
are CUDA events around one launch after a warm-up: minimum and median of `--reps`. Numbers taken here are kernel
comparisons, not bench values."""
import argparse
import json
import os
import sys
import time

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamgnn_b200 import ops, synthetic  # noqa: E402

T0 = time.perf_counter()


def timeit(fn, reps):
    fn()
    th.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        th.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {'min_ms': round(ts[0], 4), 'median_ms': round(ts[len(ts) // 2], 4)}


def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


class Out:
    def __init__(self, path):
        self.fh = open(path, 'w') if path else None

    def __call__(self, **kw):
        kw['t_s'] = round(time.perf_counter() - T0, 1)
        line = json.dumps(kw)
        print(line, flush=True)
        if self.fh:
            self.fh.write(line + '\n')
            self.fh.flush()
            os.fsync(self.fh.fileno())


def guarded(out, what, fn):
    try:
        fn()
    except Exception as e:                                       # noqa: BLE001 -- the other comparisons still run
        out(what=what, error='%s: %s' % (type(e).__name__, str(e).splitlines()[0][:300] if str(e) else ''))
        th.cuda.synchronize()


def relation_block(dst, src, rel_id, n_dst, n_src, R=2):
    """Relation-major relation block (column = r * n_src + src), as graph.RelBlock builds it."""
    return ops.CSR.from_coo(dst, rel_id * n_src + src, n_dst, n_src * R)


def spmm_compare(out, label, csr, d, reps, gen, with_index_add):
    dev = csr.device
    x = th.randn(csr.n_cols, d, device=dev, generator=gen)
    ss, ds = th.rand(csr.n_cols, device=dev, generator=gen), th.rand(csr.n_rows, device=dev, generator=gen)
    gather_gb = (csr.nnz * (4 + d * 4) + csr.n_rows * d * 4) / 1e9
    own = timeit(lambda: ops._spmm_raw(csr, x, ss, ds), reps)
    res = ops._spmm_raw(csr, x, ss, ds)
    base = dict(what='spmm', graph=label, rows=csr.n_rows, cols=csr.n_cols, nnz=csr.nnz, d=d, gather_model_GB=round(gather_gb, 2))
    out(impl='dg_spmm_csr_f32 (cj, ci fused)%s' % (', chunked aggregation T=%d' % csr.split_T if csr.split_T else ''),
        gather_TBps=round(gather_gb / own['min_ms'], 2), **base, **own)
    if csr.split_T:
        # the skewed-graph path (ops.split_plan) against the one-launch kernel on the same graph, and other chunk lengths
        t_default = csr.split_T
        for t_chunk in (0, 128, 256, 1024, 2048):
            csr.split_T, csr._plan = t_chunk, None
            t = timeit(lambda: ops._spmm_raw(csr, x, ss, ds), reps)
            out(impl='dg_spmm_csr_f32, %s' % ('chunks of %d' % t_chunk if t_chunk else 'one launch (warp per whole row)'),
                vs_default=round(t['min_ms'] / own['min_ms'], 2), own_rel_err_vs_default=rel(ops._spmm_raw(csr, x, ss, ds), res), **base, **t)
        csr.split_T, csr._plan = t_default, None

    def lib():
        a = th.sparse_csr_tensor(csr.indptr, csr.indices, th.ones(csr.nnz, device=dev), size=(csr.n_rows, csr.n_cols))
        t_alone = timeit(lambda: th.sparse.mm(a, x), reps)
        t_full = timeit(lambda: th.sparse.mm(a, x * ss.unsqueeze(1)) * ds.unsqueeze(1), reps)
        got = th.sparse.mm(a, x * ss.unsqueeze(1)) * ds.unsqueeze(1)
        out(impl='torch.sparse.mm CSR (cuSPARSE), SpMM alone', speedup_own=round(t_alone['min_ms'] / own['min_ms'], 2), **base, **t_alone)
        out(impl='torch.sparse.mm CSR (cuSPARSE) + cj / ci scale passes', speedup_own=round(t_full['min_ms'] / own['min_ms'], 2),
            own_vs_library_rel_err=rel(res, got), **base, **t_full)
    guarded(out, 'spmm cuSPARSE ' + label, lib)

    def gather_add():
        rows, cols = csr.rows().long(), csr.indices.long()
        acc = th.empty(csr.n_rows, d, device=dev)

        def run():
            acc.zero_()
            acc.index_add_(0, rows, x[cols])
        t = timeit(run, max(reps // 2, 2))
        out(impl='x[src] gather + index_add_ (no scales)', speedup_own=round(t['min_ms'] / own['min_ms'], 2), **base, **t)
    if with_index_add:
        guarded(out, 'spmm index_add ' + label, gather_add)


def decoder_compare(out, drug, dis, n_d, n_s, reps, gen):
    dev = drug.device
    e = drug.numel()
    mk = lambda *s: th.randn(*s, device=dev, generator=gen) * 0.3
    pd, ps, w2, b2, w3, b3 = mk(n_d, 128), mk(n_s, 128), mk(64, 128), mk(64), mk(1, 64), mk(1)
    pairs = ops.PairGraph(drug, dis, n_d, n_s)
    gout = th.randn(e, 1, device=dev, generator=gen)
    base = dict(what='decoder', pairs=e)

    def own_fwd():
        with th.no_grad():
            return ops.decoder_mlp(pd, ps, w2, b2, w3, b3, pairs)

    def own_fb():
        leaves = [t.detach().requires_grad_(True) for t in (pd, ps, w2, b2, w3, b3)]
        ops.decoder_mlp(*leaves, pairs).backward(gout)
        return leaves
    sl, dl = drug.long(), dis.long()

    def lib_fwd_of(pd_, ps_, w2_, b2_, w3_, b3_):
        z1 = th.relu(pd_[sl] + ps_[dl])
        return th.nn.functional.linear(th.relu(th.nn.functional.linear(z1, w2_, b2_)), w3_, b3_)

    def lib_fwd():
        with th.no_grad():
            return lib_fwd_of(pd, ps, w2, b2, w3, b3)

    def lib_fb():
        leaves = [t.detach().requires_grad_(True) for t in (pd, ps, w2, b2, w3, b3)]
        lib_fwd_of(*leaves).backward(gout)
        return leaves
    of, ofb = timeit(own_fwd, reps), timeit(own_fb, reps)
    out(impl='dg_decoder_fwd_f32 (inference, no z2 saved)', **base, **of)
    out(impl='dg_decoder fwd + bwd + node-gradient sums (deterministic)', **base, **ofb)

    def lib():
        lf, lfb = timeit(lib_fwd, max(reps // 2, 2)), timeit(lib_fb, max(reps // 2, 2))
        out(impl='torch gather + F.linear, split lin1 (forward)', speedup_own=round(lf['min_ms'] / of['min_ms'], 2), **base, **lf)
        out(impl='torch gather + F.linear under autograd (fwd + bwd; index_put accumulate)',
            speedup_own=round(lfb['min_ms'] / ofb['min_ms'], 2), **base, **lfb)
        a, b = own_fb(), lib_fb()
        out(what='decoder parity own vs torch', pairs=e, logits=rel(own_fwd(), lib_fwd()),
            **{n: rel(x.grad, y.grad) for n, x, y in zip(('dpd', 'dps', 'dw2', 'db2', 'dw3', 'db3'), a, b)})
    guarded(out, 'decoder torch', lib)


def gemm_compare(out, reps, gen, dev):
    m, k, n = 100_000, 1024, 344
    x, w = th.randn(m, k, device=dev, generator=gen), th.randn(k, n, device=dev, generator=gen) * 0.03
    base = dict(what='gemm', M=m, K=k, N=n, GFLOP=round(2 * m * k * n / 1e9, 1))
    own = timeit(lambda: ops.gemm(x, w, trans_b=True), reps)
    ref = x.double() @ w.double()
    out(impl='dg_gemm_f32 (3xTF32, tcgen05)', TFLOPs_fp32_equiv=round(2 * m * k * n / own['min_ms'] / 1e9, 1),
        rel_err_vs_f64=rel(ops.gemm(x, w, trans_b=True), ref), **base, **own)
    was = th.backends.cuda.matmul.allow_tf32
    for tf32 in (False, True):
        th.backends.cuda.matmul.allow_tf32 = tf32
        t = timeit(lambda: th.mm(x, w), reps)
        out(impl='torch.mm (cuBLAS, %s)' % ('TF32 allowed' if tf32 else 'strict fp32'), speedup_own=round(t['min_ms'] / own['min_ms'], 2),
            rel_err_vs_f64=rel(th.mm(x, w), ref), **base, **t)
    th.backends.cuda.matmul.allow_tf32 = was


def bench_block(dev, pairs=20_000_000, n_d=100_000, n_s=50_000, reps=3):
    """The short form bench.py adds to its line as `vs_library` (kernels alone, CUDA events, same inputs, this box): the
    d=344 relation-block SpMM against cuSPARSE, the projection GEMM against cuBLAS, the decoder forward against torch."""
    rows = []
    out = lambda **kw: rows.append(kw)
    gen = th.Generator(dev).manual_seed(1234)
    cells = th.unique(th.randint(0, n_d * n_s, (pairs,), generator=gen, device=dev))
    lab = (th.rand(cells.numel(), generator=gen, device=dev) < 0.01)
    order = th.argsort(lab.float(), descending=True, stable=True)
    cells, lab = cells[order], lab[order].int()
    drug, dis = (cells // n_s).int(), (cells % n_s).int()
    del cells, order
    by_dis = relation_block(dis, drug, lab, n_s, n_d)
    guarded(out, 'spmm', lambda: spmm_compare(out, 'uniform, dst=disease', by_dis, 344, reps, gen, False))
    del by_dis
    guarded(out, 'gemm', lambda: gemm_compare(out, reps, gen, dev))

    def dec():
        mk = lambda *s_: th.randn(*s_, device=dev, generator=gen) * 0.3
        pd, ps, w2, b2, w3, b3 = mk(n_d, 128), mk(n_s, 128), mk(64, 128), mk(64), mk(1, 64), mk(1)
        pg = ops.PairGraph(drug, dis, n_d, n_s)
        sl, dl = drug.long(), dis.long()
        lin = th.nn.functional.linear
        with th.no_grad():
            own = timeit(lambda: ops.decoder_mlp(pd, ps, w2, b2, w3, b3, pg), reps)
            lib = timeit(lambda: lin(th.relu(lin(th.relu(pd[sl] + ps[dl]), w2, b2)), w3, b3), reps)
        out(what='decoder forward', pairs=int(drug.numel()), impl='dg_decoder_fwd_f32', **own)
        out(what='decoder forward', pairs=int(drug.numel()), impl='torch gather + F.linear (split lin1)',
            speedup_own=round(lib['min_ms'] / own['min_ms'], 2), **lib)
    guarded(out, 'decoder', dec)
    th.cuda.empty_cache()
    keep = ('what', 'impl', 'graph', 'd', 'nnz', 'pairs', 'M', 'K', 'N', 'min_ms', 'median_ms', 'speedup_own', 'gather_TBps',
            'TFLOPs_fp32_equiv', 'rel_err_vs_f64', 'own_vs_library_rel_err', 'error')
    return {'how': 'kernels alone after a warm-up, CUDA events, minimum / median of %d, same inputs on this box; speedup_own = '
                   'library time / own time (scripts/library_compare.py; full tables: profiles/r02c_library_compare.md)' % reps,
            'rows': [{k: r[k] for k in keep if k in r} for r in rows]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='')
    ap.add_argument('--pairs', type=int, default=20_000_000)
    ap.add_argument('--nd', type=int, default=100_000)
    ap.add_argument('--ns', type=int, default=50_000)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--skip', default='', help='comma list of: spmm, decoder, gemm, zipf')
    ap.add_argument('--device', default='cuda:0')
    args = ap.parse_args()
    skip = set(filter(None, args.skip.split(',')))
    out = Out(args.out)
    dev = th.device(args.device)
    gen = th.Generator(dev).manual_seed(1234)
    n_d, n_s = args.nd, args.ns
    out(what='setup', device=th.cuda.get_device_name(0), torch=th.__version__, pairs_drawn=args.pairs, n_drug=n_d, n_dis=n_s)

    def edges_of(cells):
        labels = (th.rand(cells.numel(), generator=gen, device=dev) < 0.01)
        order = th.argsort(labels.float(), descending=True, stable=True)          # label order: positives first, by cell inside
        cells, labels = cells[order], labels[order]
        return (cells // n_s).int(), (cells % n_s).int(), labels.int()

    # most informative comparisons first: a run cut short keeps them
    cells = th.unique(th.randint(0, n_d * n_s, (args.pairs,), generator=gen, device=dev))
    drug, dis, lab = edges_of(cells)
    del cells
    by_dis = None
    if 'spmm' not in skip:
        by_dis = relation_block(dis, drug, lab, n_s, n_d)
        guarded(out, 'spmm uniform dst=disease d=344', lambda: spmm_compare(out, 'uniform, dst=disease', by_dis, 344, args.reps, gen, True))
        th.cuda.empty_cache()
    if 'gemm' not in skip:
        guarded(out, 'gemm', lambda: gemm_compare(out, args.reps, gen, dev))
        th.cuda.empty_cache()
    if 'decoder' not in skip:
        guarded(out, 'decoder', lambda: decoder_compare(out, drug, dis, n_d, n_s, args.reps, gen))
        th.cuda.empty_cache()
    if 'zipf' not in skip:
        cells = synthetic.zipf_cells(n_d, n_s, args.pairs, gen, dev)
        zdrug, zdis, zlab = edges_of(cells)
        del cells
        deg = th.bincount(zdrug.long(), minlength=n_d)
        out(what='zipf graph', distinct_pairs=int(zdrug.numel()), drug_degree_max=int(deg.max()), drug_degree_median=int(deg.median()),
            drugs_with_every_disease=int((deg == n_s).sum()), drugs_without_pairs=int((deg == 0).sum()))
        z_by_drug = relation_block(zdrug, zdis, zlab, n_d, n_s)
        guarded(out, 'spmm zipf dst=drug d=344',
                lambda: spmm_compare(out, 'zipf, dst=drug (skewed row lengths)', z_by_drug, 344, args.reps, gen, False))
        del z_by_drug
        th.cuda.empty_cache()
    if 'spmm' not in skip:
        guarded(out, 'spmm uniform dst=disease d=128', lambda: spmm_compare(out, 'uniform, dst=disease', by_dis, 128, args.reps, gen, False))
        by_dis = None
        by_drug = relation_block(drug, dis, lab, n_d, n_s)
        guarded(out, 'spmm uniform dst=drug d=344', lambda: spmm_compare(out, 'uniform, dst=drug', by_drug, 344, args.reps, gen, False))
        del by_drug
        th.cuda.empty_cache()
    if 'zipf' not in skip:
        z_by_dis = relation_block(zdis, zdrug, zlab, n_s, n_d)
        guarded(out, 'spmm zipf dst=disease d=344',
                lambda: spmm_compare(out, 'zipf, dst=disease (skewed columns)', z_by_dis, 344, args.reps, gen, False))
    out(what='done')


if __name__ == '__main__':
    main()
