#!/usr/bin/env python
"""Measure the read-bandwidth ceilings the SpMM roofline is quoted against and write profiles/l2_peak.json:

    python scripts/l2_peak.py [--out profiles/l2_peak.json]

Every figure is the best of 5 timed launches (CUDA events on the launching stream, 2 warm-up launches) of
`dg_bench_read_rows` (dreamgnn_b200/csrc/microbench.cu): one warp per row, 128-bit L1-bypassing loads, 4 rows in flight
per warp, 148 x 8 CTAs of 8 warps -- the SpMM kernels' own access shape.
  l2_seq        consecutive 512-byte rows of a 48 MiB buffer (L2-resident after the warm-up)
  l2_gather_dN  pseudo-random rows of N fp32 from a 48 MiB buffer          (L2 -> SM gather ceiling at that row width)
  hbm_gather_dN pseudo-random rows of N fp32 from a 4 GiB buffer           (HBM gather ceiling at that row width)
  hbm_seq       consecutive rows of a 4 GiB buffer, each byte read once     (HBM streaming read)
"""
import argparse
import json
import os
import sys

import torch as th

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run(lib, L, buf, n_rows, row_floats, rows_per_warp, random, sink, ctas_per_sm=8, reps=5):
    warps = 148 * ctas_per_sm * 8
    nbytes = warps * rows_per_warp * row_floats * 4
    best = None
    for i in range(reps + 2):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.dg_bench_read_rows(buf.data_ptr(), n_rows, row_floats, rows_per_warp, random, ctas_per_sm,
                                       sink.data_ptr(), L.stream()), 'bench_read_rows')
        e1.record()
        th.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if i >= 2:
            best = ms if best is None else min(best, ms)
    return {'GBps': round(nbytes / (best / 1e3) / 1e9, 1), 'ms': round(best, 4), 'bytes': nbytes}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(REPO, 'profiles', 'l2_peak.json'))
    args = ap.parse_args()
    from dreamgnn_b200 import _lib as L
    lib = L.load()
    dev = th.device('cuda:0')
    sink = th.zeros(4, device=dev)
    out = {'device': th.cuda.get_device_name(0), 'kernel': 'dg_bench_read_rows (csrc/microbench.cu)',
           'command': 'python scripts/l2_peak.py', 'how': __doc__.split('\n\n')[2].strip()}
    small = th.randn(48 << 18, device=dev)              # 48 MiB
    big = th.empty(1 << 30, device=dev).normal_()       # 4 GiB
    warps = 148 * 8 * 8
    out['l2_seq'] = run(lib, L, small, small.numel() // 128, 128, 4096, 0, sink)
    for d in (128, 344, 768):
        out['l2_gather_d%d' % d] = run(lib, L, small, small.numel() // d, d, 4096 if d <= 344 else 1024, 1, sink)
        out['hbm_gather_d%d' % d] = run(lib, L, big, big.numel() // d, d, 1024 if d <= 344 else 512, 1, sink)
    rows = big.numel() // 128
    out['hbm_seq'] = run(lib, L, big, rows, 128, (rows // warps) // 4 * 4, 0, sink)
    peaks = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.isfile(peaks):
        out['hbm_copy_peak_GBps (MEASURED_PEAKS.json)'] = json.load(open(peaks))['hbm_gbs']
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
