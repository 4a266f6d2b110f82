"""Micro-benchmark of dg_spmm_csr at the syn-20M shapes (run on the GPU box):
    python scripts/spmm_bench.py [--pairs 20000000]
Times forward (rows = destination nodes) and backward (transposed) launches for d = 344 and 128
under every DG_SPMM_VARIANT, and for the two column encodings of the relation block."""
import argparse
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamgnn_b200 import ops  # noqa: E402


def timeit(fn, reps=5):
    fn()
    th.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        th.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=20_000_000)
    ap.add_argument('--nd', type=int, default=100_000)
    ap.add_argument('--ns', type=int, default=50_000)
    args = ap.parse_args()
    dev = th.device('cuda:0')
    g = th.Generator(dev).manual_seed(0)
    cells = th.unique(th.randint(0, args.nd * args.ns, (args.pairs,), generator=g, device=dev))
    cells = cells[th.randperm(cells.numel(), generator=g, device=dev)[:int(cells.numel() * 0.9)]]   # after dropout
    drug, dis = (cells // args.ns).int(), (cells % args.ns).int()
    rel = (th.rand(cells.numel(), generator=g, device=dev) < 0.01).int()
    R = 2
    for enc in ('interleaved', 'relation-major'):
        for name, dst, src, n_dst, n_src in (('dst=disease', dis, drug, args.ns, args.nd), ('dst=drug', drug, dis, args.nd, args.ns)):
            col = src * R + rel if enc == 'interleaved' else rel * n_src + src
            csr = ops.CSR.from_coo(dst, col, n_dst, n_src * R)
            csr_t = csr.transpose()
            for d in (344, 128):
                x = th.randn(n_src * R, d, device=dev)
                gout = th.randn(n_dst, d, device=dev)
                ss, ds = th.rand(n_src * R, device=dev), th.rand(n_dst, device=dev)
                gb = (csr.nnz * (4 + d * 4) + n_dst * d * 4) / 1e9
                for var in range(5):
                    os.environ['DG_SPMM_VARIANT'] = str(var)
                    tf = timeit(lambda: ops._spmm_raw(csr, x, ss, ds))
                    tb = timeit(lambda: ops._spmm_raw(csr_t, gout, ds, ss))
                    print('%-14s %-12s d=%3d var=%d  fwd %.3f ms (%.0f GB/s)  bwd %.3f ms (%.0f GB/s)'
                          % (enc, name, d, var, tf, gb / tf * 1e3, tb, gb / tb * 1e3), flush=True)


if __name__ == '__main__':
    main()
