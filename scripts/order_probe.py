"""Diagnostic: do the real-dataset shapes time the same when measured one after another in ONE process (as bench.py's
`extra_workloads` does) as on their own?  python scripts/order_probe.py cdataset lrssl gdataset cdataset"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = argparse.Namespace(serial_aug=False, pipeline_aug=False, parallel_routes=False, messages='f32')
ctx = bench.Ctx(args)
for w in sys.argv[1:]:
    g = bench.measure(ctx, w, 1.0, steps=50, warmup=5, cuda_graph=True, e2e=True, seed=1234)
    print(w, round(g['ms_per_step'], 4), 'e2e', round(g['e2e_ms_per_step'], 4), flush=True)
