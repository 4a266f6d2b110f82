"""lrssl-shape training iteration: eager vs CUDA-graph replay (run on the GPU box)."""
import argparse
import os
import sys
import time

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamgnn_b200 import synthetic  # noqa: E402
from dreamgnn_b200.graphed import GraphedIteration  # noqa: E402
from dreamgnn_b200.model import Net  # noqa: E402
from dreamgnn_b200.train import aug_params_from_args, train_iteration  # noqa: E402
from dreamgnn_b200.utils import common_loss  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'lrssl'
dev = th.device('cuda:0')
w = synthetic.make_workload(synthetic.scaled(name, 1.0), dev, seed=0)
state = synthetic.train_state(w, dev)


def run(graphed, iters=60):
    th.manual_seed(3)
    model = Net(synthetic.model_args(w)).to(dev)
    opt = th.optim.Adam(model.parameters(), lr=0.002, weight_decay=1e-5, capturable=graphed)
    if graphed:
        it = GraphedIteration(model, opt, state)
        step = it
    else:
        loss_fn = th.nn.BCEWithLogitsLoss()
        ap = aug_params_from_args(argparse.Namespace())
        step = lambda: train_iteration(model, opt, state, loss_fn, ['edge_dropout', 'feature_noise'], ap, 0.001, 1.0, common_loss)
        for _ in range(3):
            step()
    th.cuda.synchronize()
    losses = []
    t0 = time.perf_counter()
    for i in range(iters):
        loss = step()
        if i % 20 == 19:
            losses.append(round(float(loss), 4))
    th.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3, losses


for graphed in (False, True):
    ms, losses = run(graphed)
    print('%s %s: %.3f ms/iter (%.1f it/s), loss every 20 iters %s' % (name, 'graph' if graphed else 'eager', ms, 1e3 / ms, losses), flush=True)
