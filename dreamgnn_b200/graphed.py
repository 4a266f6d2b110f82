"""CUDA-graph capture of one whole training iteration (augmentation + forward + loss + backward + clip +
Adam) for the launch-bound real-dataset shapes (lrssl / Gdataset / Cdataset: ~650 kernel launches of a few
microseconds each per iteration, SURVEY.md 7 hard part 11).

All shapes of an iteration are static (the number of kept edges is fixed by the dropout rate), the host-side
code has no device synchronisation, every random draw comes from torch's graph-safe CUDA generator (edge
permutations, feature noise, dropout masks, and the device-resident seed of the fused decoder kernel), and
the C-ABI kernels launch on the capturing stream -- so the iteration records into one `torch.cuda.CUDAGraph`
and replays with a single launch.

**Pipelined augmentation** (`pipeline_aug=True`; the default `None` enables it for launch-bound shapes: fewer
than `PIPELINE_MAX_PAIRS` scored pairs). The augmentation of train.py:254-277 depends
only on the resident base graphs / features and on the generator -- not on the parameters -- so the draw for
iteration i+1 (randperm radix sorts, keep flags, CSR compaction in both orientations, feature noise: ~10 % of
an iteration at the 20 M-edge shape, all HBM streaming work with nothing else to overlap inside its own
iteration) is recorded on a second stream as a parallel branch of iteration i's graph:

    replay i :  main   live <- staged ........ forward / loss / backward / clip / Adam on `live`
                side            (after the copy) staged <- augment(base)            [for replay i+1]

`staged` is a persistent tree of buffers (allocated by one eager augmentation before capture); `live` is its
captured clone. Every iteration still trains on exactly one fresh augmentation drawn by the same code from
the same generator; only the order in which the generator is consumed differs from the serial loop (the
draw for i+1 is interleaved with the dropout masks of i).

Measured on B200: lrssl 4.71 -> 4.31 ms / iteration, Gdataset 3.61 -> 3.28 ms (the kernels are a few microseconds
each and leave most SMs idle, so a second branch is free); at the 20 M-edge synthetic shape 68.3 -> 69.5 ms -- every
kernel there already saturates L2 / HBM, the concurrent branch only takes bandwidth from the main one and the
staged -> live copy (1.7 GB) is pure overhead -- hence the size switch.

The model must never have run on the legacy default stream before capture (autograd ties every parameter's
AccumulateGrad node to the stream of its first use): make a side stream current first, as `train()` and
`bench.py --cuda-graph` do.
"""
import torch as th
import torch.nn as nn

from . import ops
from .graph import HeteroGraph, RelBlock, _LazyEdges
from .train import augment_state, train_iteration
from .utils import common_loss

PIPELINE_MAX_PAIRS = 4_000_000     # above this the iteration is bandwidth-bound and a concurrent branch does not pay


# ---- structural clone / in-place refresh of an augmentation result -----------------------------------
class _Arena:
    """Bump allocator over ONE device buffer. Two arenas filled by the same sequence of requests have identical layouts,
    so a whole tree of tensors living in one arena is copied into its twin with a single device-to-device copy (the
    `live <- staged` step every replay starts with: ~60 tensors = ~60 serial copy nodes at the head of the graph before)."""
    ALIGN = 256

    def __init__(self, nbytes, device):
        self.buf = th.empty(max(int(nbytes), 1), dtype=th.uint8, device=device)
        self.off = 0

    @classmethod
    def span(cls, t):
        return (t.numel() * t.element_size() + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN

    def view_like(self, t):
        """An uninitialised contiguous tensor of t's dtype and shape inside the arena."""
        n = t.numel() * t.element_size()
        if self.off + n > self.buf.numel():
            raise ValueError('arena too small')
        v = self.buf[self.off:self.off + n].view(t.dtype).view(t.shape)
        self.off += self.span(t)
        return v

    def take_like(self, t):
        """...holding a copy of t."""
        v = self.view_like(t)
        v.copy_(t)
        return v


def _csr_pair(c):
    """The CSR and (if present) its cached transpose, each once."""
    return [c] if c._t is None else [c, c._t]


def _clone_csr(c, take):
    def one(x):
        n = ops.CSR(take(x.indptr), take(x.indices), take(x.eid), None if x.vals is None else take(x.vals), x.n_rows, x.n_cols)
        n.slot_order, n.eid_is_slot, n.parent_nnz = x.slot_order, x.eid_is_slot, x.parent_nnz
        ops.inherit_layout(n, x)
        return n
    n = one(c)
    if c._t is not None:
        n._t = one(c._t)
        n._t._t = n
    return n


def _csr_tensors(c):
    out = []
    for x in _csr_pair(c):
        out += [x.indptr, x.indices, x.eid] + ([] if x.vals is None else [x.vals])
    return out


def _clone_entry(v, take):
    """Deep copy of one value of an augmentation dict -- dense tensor, sparse-COO adjacency with its CSR sidecar, or a
    dropped HeteroGraph (relation blocks prebuilt) -- with every tensor obtained from `take(source)` in the order
    `_entry_tensors` lists them. Anything else cannot be staged."""
    if isinstance(v, th.Tensor) and v.is_sparse:
        csr = getattr(v, '_dg_csr', None)
        if csr is None:
            raise ValueError('sparse adjacency without a CSR sidecar')
        t = th.sparse_coo_tensor(take(v._indices()), take(v._values()), v.shape, device=v.device, check_invariants=False)
        t._dg_csr = _clone_csr(csr, take)
        return t
    if isinstance(v, th.Tensor):
        return take(v)
    if isinstance(v, HeteroGraph):
        if not v._blocks or any(not isinstance(e, _LazyEdges) for e in v._edges.values()):
            raise ValueError('only edge-dropped graphs (prebuilt relation blocks) can be staged')
        blocks, lazy = {}, {}
        for dt in sorted(v._blocks):
            b = v._blocks[dt]
            nb = RelBlock(b.etypes, b.src_type, b.dst_type, b.n_src, b.n_dst, _clone_csr(b.csr, take), b.offsets)
            blocks[dt] = nb
            for r, c in enumerate(b.etypes):
                lazy[c] = _LazyEdges(nb, r, v._edges[c].count, v.idtype)
        nd = {nt: {k: take(v._ndata[nt][k]) for k in sorted(v._ndata[nt])} for nt in sorted(v._ndata)}
        return HeteroGraph(lazy, v._num_nodes, nd, None, v.idtype, blocks)
    raise ValueError('cannot stage %r' % type(v))


def _entry_tensors(v):
    if isinstance(v, th.Tensor) and v.is_sparse:
        return [v._indices(), v._values()] + _csr_tensors(v._dg_csr)
    if isinstance(v, th.Tensor):
        return [v]
    out = []
    for dt in sorted(v._blocks):
        out += _csr_tensors(v._blocks[dt].csr)
    for nt in sorted(v._ndata):
        out += [v._ndata[nt][k] for k in sorted(v._ndata[nt])]
    return out


class StagedAugmentation:
    """Persistent buffers holding one augmentation result, all inside one arena: `clone()` gives an independent copy with
    the same structure in a twin arena (ONE device copy), `refresh(new)` overwrites the buffers in place with a freshly
    drawn result (the shapes are checked; one multi-tensor copy launch on the augmentation branch, off the main chain)."""

    def __init__(self, aug, base):
        # entries the augmentation left untouched alias the resident inputs and need no staging
        self.keys = [k for k, v in aug.items() if v is not None and v is not base.get(k)]
        self.passthrough = {k: v for k, v in aug.items() if k not in self.keys}
        for k in self.keys:
            _clone_entry(aug[k], lambda t: t)                      # raises for anything that cannot be staged
        device = next(t.device for k in self.keys for t in _entry_tensors(aug[k])) if self.keys else 'cpu'
        self._nbytes = sum(_Arena.span(t) for k in self.keys for t in _entry_tensors(aug[k]))
        self.arena = _Arena(self._nbytes, device)
        self.tree = {k: _clone_entry(aug[k], self.arena.take_like) for k in self.keys}

    def clone(self):
        """An independent copy (what `live` is): same structure over a twin arena filled by one device-to-device copy."""
        twin = _Arena(self._nbytes, self.arena.buf.device)
        out = dict(self.passthrough)
        out.update({k: _clone_entry(v, twin.view_like) for k, v in self.tree.items()})
        import os
        if os.environ.get('DG_STAGE_ARENA', '1') != '0':
            twin.buf.copy_(self.arena.buf)
        else:                                                      # A/B: one copy per tensor, as before the arena
            for k in self.keys:
                for d, s_ in zip(_entry_tensors(out[k]), _entry_tensors(self.tree[k])):
                    d.copy_(s_)
        self._twins = getattr(self, '_twins', []) + [twin]         # keep the twin alive with its views' owner
        return out

    def refresh(self, new):
        dsts, srcs = [], []
        for k in self.keys:
            dst, src = _entry_tensors(self.tree[k]), _entry_tensors(new[k])
            if len(dst) != len(src) or any(d.shape != s.shape or d.dtype != s.dtype for d, s in zip(dst, src)):
                raise ValueError('augmentation result %r changed shape between iterations' % k)
            dsts += dst
            srcs += src
        import os
        if dsts and dsts[0].is_cuda and os.environ.get('DG_MULTI_COPY', '1') != '0':
            ops.multi_copy(dsts, [s if s.is_contiguous() else s.contiguous() for s in srcs])      # one launch for all of them
        else:
            for d, s in zip(dsts, srcs):
                d.copy_(s)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for k in self.keys for t in _entry_tensors(self.tree[k]))


class GraphedIteration:
    def __init__(self, model, optimizer, state, rel_loss_fn=None, aug_methods=('edge_dropout', 'feature_noise'),
                 aug_params=None, beta=0.001, grad_clip=1.0, common_loss_fn=common_loss, warmup=3, pipeline_aug=None,
                 parallel_routes=None):
        if not all(g.get('capturable', False) for g in optimizer.param_groups):
            raise ValueError('the optimizer must be built with capturable=True (e.g. torch.optim.Adam(..., capturable=True))')
        self.model, self.optimizer, self.state = model, optimizer, state
        loss_fn = rel_loss_fn or nn.BCEWithLogitsLoss()
        aug_methods = list(aug_methods)
        aug_params = aug_params or {'edge_dropout_rate': 0.1, 'feature_noise_scale': 0.05}
        self._step = lambda aug=None: train_iteration(model, optimizer, state, loss_fn, aug_methods, aug_params, beta,
                                                      grad_clip, common_loss_fn, aug=aug)
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):                       # eager warm-up: builds every cached CSR, cuBLAS handles, ...
            self.eager_iterations = max(warmup, 1)       # at least one: lazy caches must exist before capture
            for _ in range(self.eager_iterations):
                self._step()
        th.cuda.current_stream().wait_stream(side)
        th.cuda.synchronize()

        def base_inputs():
            return {'enc_graph': state.enc_graph, 'drug_graph': state.drug_graph, 'disease_graph': state.dis_graph,
                    'drug_feature_graph': state.drug_feature_graph,
                    'disease_feature_graph': state.disease_feature_graph, 'drug_feat': state.drug_feat,
                    'disease_feat': state.dis_feat, 'drug_sim_feat': state.drug_sim_feat,
                    'disease_sim_feat': state.dis_sim_feat}

        self.staged = None
        small = int(state.labels.numel()) < PIPELINE_MAX_PAIRS
        if parallel_routes is None:
            parallel_routes = small
        if hasattr(model, 'parallel_routes'):
            model.parallel_routes = bool(parallel_routes)       # GCMC route || FGCN route as two branches of the graph
        if pipeline_aug is None:
            pipeline_aug = small
        if pipeline_aug and aug_methods:
            try:
                # the first replay trains on this eagerly drawn augmentation
                self.staged = StagedAugmentation(augment_state(state, aug_methods, aug_params), base_inputs())
            except ValueError:
                self.staged = None                       # e.g. add_random_edges: falls back to the serial capture
            th.cuda.synchronize()
        self.graph = th.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        from . import _lib
        launches0 = _lib.launch_count()
        with th.cuda.graph(self.graph):
            if self.staged is None:
                self.loss = self._step()
            else:
                main = th.cuda.current_stream()
                self._live = self.staged.clone()                         # live <- staged (the copy every replay runs)
                self._aug_stream = ops.augmentation_stream(main)
                self._aug_stream.wait_stream(main)                       # fork: after the copy has read `staged`
                with th.cuda.stream(self._aug_stream):
                    self.staged.refresh(augment_state(state, aug_methods, aug_params))
                self.loss = self._step(self._live)
                main.wait_stream(self._aug_stream)                       # join
        ops.drop_seed_pool()                                             # its buffer now belongs to the graph
        if hasattr(model, 'parallel_routes'):
            model.parallel_routes = False                                # the branches are in the graph; eager calls stay serial
        self.launches_per_replay = _lib.launch_count() - launches0       # C-ABI kernels recorded into the graph

    def __call__(self):
        """Run one iteration; returns the (static) device tensor holding its loss."""
        self.graph.replay()
        return self.loss
