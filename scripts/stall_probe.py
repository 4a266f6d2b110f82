"""Debug: find what the host thread is doing during slow steps (run on the GPU box)."""
import collections
import os
import sys
import threading
import time
import traceback

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamgnn_b200 import synthetic  # noqa: E402
from dreamgnn_b200.model import Net  # noqa: E402
from dreamgnn_b200.train import train_iteration, aug_params_from_args  # noqa: E402
from dreamgnn_b200.utils import common_loss_gram  # noqa: E402
import argparse  # noqa: E402

dev = th.device('cuda:0')
spec = synthetic.scaled('syn20m', 1.0)
w = synthetic.make_workload(spec, dev, seed=1234)
state = synthetic.train_state(w, dev)
model = Net(synthetic.model_args(w)).to(dev)
opt = th.optim.Adam(model.parameters(), lr=0.002, weight_decay=1e-5)
loss_fn = th.nn.BCEWithLogitsLoss()
ap = aug_params_from_args(argparse.Namespace())
step = lambda: train_iteration(model, opt, state, loss_fn, ['edge_dropout', 'feature_noise'], ap, 0.001, 1.0, common_loss_gram)
for _ in range(3):
    step()
th.cuda.synchronize()
main_id = threading.main_thread().ident
samples = []
stop = threading.Event()


def sampler():
    while not stop.is_set():
        fr = sys._current_frames().get(main_id)
        if fr is not None:
            st = traceback.extract_stack(fr)
            samples.append((time.perf_counter(), ' <- '.join('%s:%d' % (os.path.basename(f.filename), f.lineno) for f in st[-4:])))
        time.sleep(0.005)


t = threading.Thread(target=sampler, daemon=True)
t.start()
bounds = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    t0 = time.perf_counter()
    step()
    bounds.append((t0, time.perf_counter()))
th.cuda.synchronize()
stop.set()
durs = [round((b - a) * 1e3, 1) for a, b in bounds]
print('host ms per step:', durs)
med = sorted(durs)[len(durs) // 2]
for (a, b), d in zip(bounds, durs):
    if d > med * 1.4:
        c = collections.Counter(s for ts, s in samples if a <= ts <= b)
        print('SLOW step %.1f ms; top frames:' % d)
        for s, n in c.most_common(4):
            print('   %3d  %s' % (n, s))
c = collections.Counter(s for ts, s in samples)
print('overall top frames:')
for s, n in c.most_common(6):
    print('   %4d  %s' % (n, s))
