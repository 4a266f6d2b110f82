"""Standalone decoder forward + backward at the syn20m pair count, for ncu source-level captures:
    ncu --set full --import-source on -k regex:decoder_.*tc -o /tmp/dec python scripts/decoder_profile.py
    ncu -i /tmp/dec.ncu-rep --page source --csv > gpurun_out/decoder_source.csv
Pairs are sorted by drug inside two label classes like the bench's generator; prints CUDA-event times."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamgnn_b200 import ops as o  # noqa: E402

dev = th.device('cuda:0')
gen = th.Generator(dev).manual_seed(1234)
n_d, n_s, n = 100_000, 50_000, int(os.environ.get('PAIRS', 20_000_000))
cells = th.unique(th.randint(0, n_d * n_s, (n,), generator=gen, device=dev))
labels = (th.rand(cells.numel(), generator=gen, device=dev) < 0.01).float()
cells = cells[th.argsort(labels, descending=True, stable=True)]
pairs = o.PairGraph((cells // n_s).to(th.int32), (cells % n_s).to(th.int32), n_d, n_s)
mk = lambda *s, k=1.0: (th.randn(*s, generator=gen, device=dev) * k).requires_grad_(True)
pd, ps, w2, b2, w3, b3 = mk(n_d, 128), mk(n_s, 128), mk(64, 128, k=0.1), mk(64, k=0.1), mk(1, 64, k=0.2), mk(1, k=0.1)
gout = th.randn(pairs.n_pairs, 1, generator=gen, device=dev) / pairs.n_pairs
for it in range(int(os.environ.get('REPS', 3))):
    e = [th.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = o.decoder_mlp(pd, ps, w2, b2, w3, b3, pairs, p=0.3, seed=5, training=True)
    e[1].record()
    out.backward(gout)
    e[2].record()
    th.cuda.synchronize()
    print('pairs %d: forward %.3f ms, backward (incl. node-gradient sums) %.3f ms' % (pairs.n_pairs, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
