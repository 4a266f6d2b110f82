"""CUDA-graph capture of one whole training iteration (augmentation + forward + loss + backward + clip +
Adam) for the launch-bound real-dataset shapes (lrssl / Gdataset / Cdataset: ~650 kernel launches of a few
microseconds each per iteration, SURVEY.md 7 hard part 11).

All shapes of an iteration are static (the number of kept edges is fixed by the dropout rate), the host-side
code has no device synchronisation, every random draw comes from torch's graph-safe CUDA generator (edge
permutations, feature noise, dropout masks, and the device-resident seed of the fused decoder kernel), and
the C-ABI kernels launch on the capturing stream -- so the iteration records into one `torch.cuda.CUDAGraph`
and replays with a single launch.

The model must never have run on the legacy default stream before capture (autograd ties every parameter's
AccumulateGrad node to the stream of its first use): make a side stream current first, as `train()` and
`bench.py --cuda-graph` do.
"""
import torch as th
import torch.nn as nn

from .train import train_iteration
from .utils import common_loss


class GraphedIteration:
    def __init__(self, model, optimizer, state, rel_loss_fn=None, aug_methods=('edge_dropout', 'feature_noise'),
                 aug_params=None, beta=0.001, grad_clip=1.0, common_loss_fn=common_loss, warmup=3):
        if not all(g.get('capturable', False) for g in optimizer.param_groups):
            raise ValueError('the optimizer must be built with capturable=True (e.g. torch.optim.Adam(..., capturable=True))')
        self.model, self.optimizer, self.state = model, optimizer, state
        loss_fn = rel_loss_fn or nn.BCEWithLogitsLoss()
        aug_params = aug_params or {'edge_dropout_rate': 0.1, 'feature_noise_scale': 0.05}
        self._step = lambda: train_iteration(model, optimizer, state, loss_fn, list(aug_methods), aug_params, beta,
                                             grad_clip, common_loss_fn)
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):                       # eager warm-up: builds every cached CSR, cuBLAS handles, ...
            for _ in range(max(warmup, 1)):              # at least one eager iteration: lazy caches must exist before capture
                self._step()
        th.cuda.current_stream().wait_stream(side)
        th.cuda.synchronize()
        self.graph = th.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with th.cuda.graph(self.graph):
            self.loss = self._step()

    def __call__(self):
        """Run one iteration; returns the (static) device tensor holding its loss."""
        self.graph.replay()
        return self.loss
