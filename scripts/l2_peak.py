#!/usr/bin/env python
"""Measure the read-bandwidth ceilings the SpMM roofline is quoted against and write profiles/l2_peak.json:

    python scripts/l2_peak.py [--out profiles/l2_peak.json]

Every figure is the best of 5 timed launches (CUDA events on the launching stream, 2 warm-up launches) of
`dg_bench_read_rows` (dreamgnn_b200/csrc/microbench.cu): one warp per row, 128-bit L1-bypassing loads, U rows in flight
per warp, 148 x C CTAs of 8 warps -- the SpMM kernels' own access shape; for every row width the best (U, C) is kept
and all points are listed.
  l2_gather_dN  pseudo-random rows of N fp32 from a 48 MiB buffer (L2-resident after the warm-up): L2 -> SM gather ceiling
  hbm_gather_dN the same rows from a 4 GiB buffer: HBM gather ceiling at that row width
  l2_seq / hbm_seq   consecutive 512-byte rows of the two buffers (streaming read; hbm_seq reads every byte once)
d = 344 is the GCMC layer-0 message width (1376-byte rows: not a multiple of the 128-byte line), d = 352 the same padded
to whole lines, d = 128 the layer-1/2 messages, d = 768 the FGCN hidden width.
"""
import argparse
import json
import os
import sys

import torch as th

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run(lib, L, buf, n_rows, row_floats, random, sink, ctas_per_sm, in_flight, total_bytes, reps=5):
    warps = 148 * ctas_per_sm * 8
    rows_per_warp = max(in_flight, int(total_bytes / (warps * row_floats * 4)) // in_flight * in_flight)
    nbytes = warps * rows_per_warp * row_floats * 4
    best = None
    for i in range(reps + 2):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.dg_bench_read_rows(buf.data_ptr(), n_rows, row_floats, rows_per_warp, random, ctas_per_sm, in_flight,
                                       sink.data_ptr(), L.stream()), 'bench_read_rows')
        e1.record()
        th.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if i >= 2:
            best = ms if best is None else min(best, ms)
    return {'GBps': round(nbytes / (best / 1e3) / 1e9, 1), 'ms': round(best, 4), 'bytes': nbytes, 'rows_in_flight': in_flight,
            'ctas_per_sm': ctas_per_sm}


def sweep(lib, L, buf, d, random, sink, total_bytes, n_rows=None):
    pts = []
    for c in (4, 8):
        for u in (4, 8, 16):
            if u * ((d + 127) // 128) > 48:                     # > 48 float4 in flight per lane would spill
                continue
            pts.append(run(lib, L, buf, n_rows or buf.numel() // d, d, random, sink, c, u, total_bytes))
    best = max(pts, key=lambda p: p['GBps'])
    return dict(best, points=[(p['ctas_per_sm'], p['rows_in_flight'], p['GBps']) for p in pts])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(REPO, 'profiles', 'l2_peak.json'))
    args = ap.parse_args()
    from dreamgnn_b200 import _lib as L
    lib = L.load()
    dev = th.device('cuda:0')
    sink = th.zeros(4, device=dev)
    out = {'device': th.cuda.get_device_name(0), 'kernel': 'dg_bench_read_rows (csrc/microbench.cu)',
           'command': 'python scripts/l2_peak.py', 'how': ' '.join(__doc__.split('\n\n')[2].split()),
           'points': '(ctas_per_sm, rows_in_flight, GB/s)'}
    small = th.randn(48 << 18, device=dev)              # 48 MiB
    big = th.empty(1 << 30, device=dev).normal_()       # 4 GiB
    out['l2_seq'] = sweep(lib, L, small, 128, 0, sink, 24e9)
    for d in (128, 344, 352, 768):
        out['l2_gather_d%d' % d] = sweep(lib, L, small, d, 1, sink, 24e9)
        out['hbm_gather_d%d' % d] = sweep(lib, L, big, d, 1, sink, 8e9)
    out['hbm_seq'] = sweep(lib, L, big, 128, 0, sink, 4.29e9)
    peaks = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.isfile(peaks):
        out['hbm_copy_peak_GBps (MEASURED_PEAKS.json)'] = json.load(open(peaks))['hbm_gbs']
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: (v['GBps'] if isinstance(v, dict) else v) for k, v in out.items()}))


if __name__ == '__main__':
    main()
