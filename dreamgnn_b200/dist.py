"""1-D row partitioning of one large graph across the GPUs of a box (BASELINE config 5).

Every rank owns a contiguous block of drugs and of diseases: their input features, their rows of every
relation-block CSR (all in-edges of the owned destination nodes), their rows of the kNN adjacencies and
a slice of the scored pairs. The only data-path exchange is one all-gather of node rows per aggregation:

    forward   h_local [rows_local, D]  --all-gather over NVLink-->  H [world * rows_local, D]
    backward  dH [world * rows_local, D]  --reduce-scatter(sum)-->  dh_local            (autograd adjoint)

so the SpMM / decoder kernels run unchanged on rank-local CSRs whose column ids index the gathered
buffer (rank-major: column of (node i of relation r) = owner(i) * R * n_loc + r * n_loc + i % n_loc).
Weights are replicated; their gradients are summed with one flat all-reduce per step; scalar losses and
the 128x128 Gram matrices of the common loss go through a differentiable all-reduce.

The joint objective is sum_p L_p: each rank back-propagates L_p = (local BCE sum) / E_global
+ beta * common / world, and the collectives' adjoints (reduce-scatter for all-gather, all-reduce for
all-reduce) make the per-rank backward passes add up to the gradient of the global loss.

Edge dropout draws one th.randperm per relation over the rank's LOCAL edges (the reference draws one
global permutation, augmentation.py:51): every rank keeps exactly int(E_local * (1 - rate)) of its edges,
distributionally the same keep rule.
"""
import torch as th
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .graph import RelBlock


# Optional collective log for bench.py: when PROFILE is a list every collective appends
# (kind, payload bytes received / reduced on this rank, start event, end event) on the compute stream.
PROFILE = None


class _Timed:
    def __init__(self, kind, nbytes):
        self.kind, self.nbytes = kind, int(nbytes)

    def __enter__(self):
        if PROFILE is not None:
            self.e0, self.e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.kind, self.nbytes, self.e0, self.e1))
        return False


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


# Forward collectives issued ahead of their consumer (`all_gather_rows(x, deferred=True)`): NCCL runs them on its own
# stream, so kernels enqueued on the compute stream in the meantime (the other node type's projection / aggregation)
# overlap the transfer; `wait_rows(t)` makes the compute stream wait right before the first read. Keyed by storage.
_PENDING = {}


def wait_rows(t):
    """Block the current STREAM (not the host) until a deferred all-gather into `t` has landed. No-op otherwise."""
    work = _PENDING.pop(t.data_ptr(), None) if isinstance(t, th.Tensor) else None
    if work is not None:
        with _Timed('all_gather_wait', 0):
            work.wait()
    return t


class _AllGatherRows(th.autograd.Function):
    @staticmethod
    def forward(ctx, x, deferred):
        world = _world()
        x = x.contiguous()
        out = th.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        if deferred and dist.get_backend() == 'nccl':
            if PROFILE is not None:
                PROFILE.append(('all_gather_bytes_deferred', out.numel() * out.element_size(), None, None))
            _PENDING[out.data_ptr()] = dist.all_gather_into_tensor(out, x, async_op=True)
            return out
        with _Timed('all_gather', out.numel() * out.element_size()):
            dist.all_gather_into_tensor(out, x)
        return out

    @staticmethod
    def backward(ctx, g):
        world = _world()
        g = g.contiguous()
        n = g.shape[0] // world
        if dist.get_backend() == 'gloo':                      # gloo (CPU tests) has no reduce-scatter: all-reduce + slice
            g = g.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            return g[_rank() * n:(_rank() + 1) * n].clone(), None
        out = th.empty((n,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        with _Timed('reduce_scatter', g.numel() * g.element_size()):
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
        return out, None


class _AllReduceSum(th.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = x.clone()
        with _Timed('all_reduce', y.numel() * y.element_size()):
            dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        with _Timed('all_reduce', g.numel() * g.element_size()):
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


def all_gather_rows(x, deferred=False):
    """[n_loc, ...] on every rank -> [world * n_loc, ...] (rank-major); backward = reduce-scatter(sum).
    `deferred`: enqueue the collective and return at once -- call `wait_rows` on the result before it is read."""
    return _AllGatherRows.apply(x, deferred) if _world() > 1 else x


def all_gather_plain(x):
    """Non-differentiable all-gather of a small per-node vector (dropout-scaled cj: data, not a parameter)."""
    if _world() == 1:
        return x
    x = x.contiguous()
    out = th.empty((_world() * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    with _Timed('all_gather', out.numel() * out.element_size()):
        dist.all_gather_into_tensor(out, x)
    return out


def all_reduce_sum(x):
    """Differentiable sum over ranks (adjoint = all-reduce of the incoming gradient)."""
    return _AllReduceSum.apply(x) if _world() > 1 else x


class Partition:
    """Contiguous equal blocks of each node type; N must be divisible by the world size."""

    def __init__(self, num_nodes, rank=None, world=None):
        self.rank = _rank() if rank is None else rank
        self.world = _world() if world is None else world
        self.n = dict(num_nodes)
        for nt, n in self.n.items():
            if n % self.world:
                raise ValueError('%s count %d is not divisible by the world size %d' % (nt, n, self.world))
        self.n_loc = {nt: n // self.world for nt, n in self.n.items()}

    def lo(self, nt):
        return self.rank * self.n_loc[nt]

    def owned(self, ids, nt):
        return th.div(ids, self.n_loc[nt], rounding_mode='floor') == self.rank

    def local(self, ids, nt):
        return ids - self.lo(nt)

    def gathered_index(self, ids, nt, num_rel=1, rel=0):
        """Row of (node `ids`, relation `rel`) in the all-gathered [world * R * n_loc, D] buffer."""
        n_loc = self.n_loc[nt]
        owner = th.div(ids, n_loc, rounding_mode='floor')
        return owner * (num_rel * n_loc) + rel * n_loc + (ids - owner * n_loc)

    def rows(self, x, nt):
        """This rank's rows of a replicated [N, ...] tensor."""
        return x[self.lo(nt):self.lo(nt) + self.n_loc[nt]]


class _NodeView:
    def __init__(self, data):
        self.data = data


class _Nodes:
    def __init__(self, store):
        self._store = store

    def __getitem__(self, nt):
        return _NodeView(self._store[nt])


class PartitionedEncGraph:
    """Rank-local part of the encoder graph: for each destination type the relation block of the OWNED
    destination rows (columns index the gathered message buffer) + local ci / cj."""

    ETYPES = {'drug': [('disease', 'rev-0', 'drug'), ('disease', 'rev-1', 'drug')],
              'disease': [('drug', '0', 'disease'), ('drug', '1', 'disease')]}

    def __init__(self, partition, blocks, ndata):
        self.partition, self._blocks, self._ndata = partition, blocks, ndata
        self.nodes = _Nodes(ndata)

    @property
    def canonical_etypes(self):
        return self.ETYPES['drug'] + self.ETYPES['disease']

    @property
    def device(self):
        return self._blocks['drug'].csr.device

    def block(self, dst_type):
        return self._blocks[dst_type]

    def number_of_edges(self, c):
        blk = self._blocks[c[2]]
        r = blk.etypes.index(c)
        return blk.rel_counts[r]

    @staticmethod
    def from_pairs(pairs, labels, partition, device):
        """pairs = (drug ids, disease ids) of ALL scored pairs (replicated), labels in {0,1}: relation r."""
        dev = th.device(device)
        drug, dis = pairs[0].to(dev).long(), pairs[1].to(dev).long()
        rel = labels.to(dev).long()
        blocks, ndata = {}, {}
        for dst_type, (dst, src, src_type) in (('drug', (drug, dis, 'disease')), ('disease', (dis, drug, 'drug'))):
            own = partition.owned(dst, dst_type)
            d, s, r = dst[own], src[own], rel[own]
            order = th.argsort(r, stable=True)                           # relation-major edge ids, like RelBlock
            d, s, r = d[order], s[order], r[order]
            counts = [int((r == k).sum()) for k in (0, 1)]
            cols = partition.gathered_index(s, src_type, 2, 0) + r * partition.n_loc[src_type]
            csr = ops.CSR.from_coo(partition.local(d, dst_type), cols, partition.n_loc[dst_type],
                                   partition.world * 2 * partition.n_loc[src_type])
            blk = RelBlock(PartitionedEncGraph.ETYPES[dst_type], src_type, dst_type,
                           partition.world * partition.n_loc[src_type], partition.n_loc[dst_type], csr, [0, counts[0]])
            blk.rel_counts = counts
            blocks[dst_type] = blk
            ci = csr.degree_norm().unsqueeze(1)
            ndata[dst_type] = {'ci': ci, 'cj': ci.clone()}               # symmetric construction: cj == ci
        return PartitionedEncGraph(partition, blocks, ndata)

    def edge_dropout(self, rate):
        from .augmentation import num_keep_edges
        blocks = {}
        for dt, base in self._blocks.items():
            plist, n_keep, counts = [], 0, []
            for off, cnt in zip(base.offsets, base.rel_counts):
                k = num_keep_edges(cnt, rate) if cnt else 0
                if cnt:
                    plist.append((th.randperm(cnt, device=base.csr.device), k, off))
                n_keep += k
                counts.append(k)
            flags = ops.keep_flags(base.csr.nnz, plist, base.csr.device)
            blk = RelBlock(base.etypes, base.src_type, base.dst_type, base.n_src, base.n_dst,
                           ops.csr_dropout(base.csr, flags, n_keep), base.offsets)
            blk.rel_counts = counts
            blocks[dt] = blk
        return PartitionedEncGraph(self.partition, blocks, self._ndata)


class PartitionedAdjacency:
    """Owned rows of a kNN adjacency; columns index the all-gathered [N, D] support."""

    def __init__(self, csr, partition, nt):
        self.csr, self.partition, self.nt = csr, partition, nt

    @staticmethod
    def from_sparse(adj, partition, nt):
        idx, val = adj._indices(), adj._values()
        own = partition.owned(idx[0], nt)
        csr = ops.CSR.from_coo(partition.local(idx[0][own], nt), idx[1][own], partition.n_loc[nt], partition.n[nt],
                               val[own].to(th.float32))
        return PartitionedAdjacency(csr, partition, nt)

    def edge_dropout(self, rate):
        from .augmentation import num_keep_edges
        n = self.csr.nnz
        k = num_keep_edges(n, rate)
        flags = ops.keep_flags(n, [(th.randperm(n, device=self.csr.device), k, 0)], self.csr.device)
        return PartitionedAdjacency(ops.csr_dropout(self.csr, flags, k), self.partition, self.nt)


class PartitionedPairs:
    """This rank's slice of the scored pairs (label order kept inside the slice); ids index the
    all-gathered node projections."""

    def __init__(self, pairs, labels, partition, device):
        dev = th.device(device)
        e = pairs[0].numel()
        per = (e + partition.world - 1) // partition.world
        lo, hi = min(partition.rank * per, e), min((partition.rank + 1) * per, e)
        self.n_global = e
        self.labels = labels[lo:hi].to(dev).float()
        self.pairs = ops.PairGraph(pairs[0][lo:hi].to(dev), pairs[1][lo:hi].to(dev), partition.n['drug'],
                                   partition.n['disease'])
        self.partition = partition


def gcmc_exchange(x, wstack, scale):
    """First half of the distributed GCMC aggregation into one destination type: make the messages of ALL source nodes
    available on this rank, in the rank-major layout the partitioned block's columns index
    (row = owner * R * n_loc + r * n_loc + local id), moving as few bytes over NVLink as the layer allows:

      * wide inputs (layer 0: 1024 / 768 features -> R x 344 messages): project the owned rows, pre-scale with
        dropout(cj), all-gather the [R * n_loc, D_msg] messages;
      * narrow inputs (layers 1-2: 128 -> R x 128): all-gather the [n_loc, 128] EMBEDDINGS (R times fewer bytes; the
        per-relation dropout(cj) scales travel as [R * n_loc] floats), project after the gather -- every rank repeats the
        small N x 128 x (R * 128) product -- and let the SpMM apply the scales.

    Returns (gathered buffer or pending gather, src_scale for the SpMM or None, finish): `finish(buf)` completes the
    exchange (waits for a deferred gather, projects if the embeddings were gathered) and returns the [world*R*n_loc, D]
    message matrix. The collective is issued deferred, so work enqueued between `gcmc_exchange` and `finish` overlaps it."""
    R, k_in, dp = wstack.shape
    n_loc = x.shape[0]
    world = _world()
    if k_in < R * dp:                                   # gather embeddings, project afterwards
        xg = all_gather_rows(x, deferred=True)          # [world * n_loc, k_in]
        sg = all_gather_plain(scale.view(R, n_loc)).reshape(-1)      # [world, R, n_loc] = the buffer's row order

        def finish(buf):
            h = ops.project(wait_rows(buf), wstack)     # [R, world * n_loc, dp]
            h = h.view(R, world, n_loc, dp).permute(1, 0, 2, 3).reshape(world * R * n_loc, dp)
            return h
        return xg, sg, finish
    h = ops.project(x, wstack)                          # [R, n_loc, dp]
    hg = all_gather_rows((h * scale.view(R, n_loc, 1)).reshape(R * n_loc, dp), deferred=True)
    return hg, None, wait_rows


def gcmc_aggregate(layer_scale, h, blk, ci):
    """One-shot form (messages already projected): pre-scale with dropout(cj), all-gather, aggregate the owned rows."""
    R, n_loc, dp = h.shape
    hg = all_gather_rows((h * layer_scale.view(R, n_loc, 1)).reshape(R * n_loc, dp))
    return ops.spmm(blk.csr, hg, src_scale=None, dst_scale=ci, tag='gcmc')


def common_loss_partitioned(emb1, emb2, n_global):
    """utils.common_loss over row-partitioned embeddings in its Gram form (no N x N, no gather of rows):
    global means and the three 128 x 128 Gram matrices are summed over ranks."""
    return common_losses_partitioned([(emb1, emb2, n_global)])[0]


def common_losses_partitioned(pairs):
    """[common_loss(emb1, emb2) for (emb1, emb2, n_global) in pairs] with TWO collectives for all of them (forward; two
    more in the backward): the column sums of every embedding travel in one all-reduce, the Gram matrices of every pair in
    a second one. At 8 ranks the 20 separate all-reduces of the two node types' losses cost as much as the 1.9 GB of
    all-gathers (latency, not bytes)."""
    d = pairs[0][0].shape[1]
    sums = all_reduce_sum(th.cat([e.sum(0) for e1, e2, _ in pairs for e in (e1, e2)]))
    zs = []
    for i, (e1, e2, n) in enumerate(pairs):
        m1, m2 = sums[(2 * i) * d:(2 * i + 1) * d] / float(n), sums[(2 * i + 1) * d:(2 * i + 2) * d] / float(n)
        zs.append((F.normalize(e1 - m1, p=2, dim=1).double(), F.normalize(e2 - m2, p=2, dim=1).double()))
    grams = all_reduce_sum(th.stack([g for z1, z2 in zs for g in (z1.t() @ z1, z2.t() @ z2, z1.t() @ z2)]))
    out = []
    for i, (_, _, n) in enumerate(pairs):
        g11, g22, g12 = grams[3 * i], grams[3 * i + 1], grams[3 * i + 2]
        out.append((((g11 ** 2).sum() + (g22 ** 2).sum() - 2.0 * (g12 ** 2).sum()) / float(n) / float(n)).float())
    return out


def all_reduce_gradients(params):
    """Sum the replicated weights' gradients over ranks with one flat all-reduce."""
    if _world() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = th.cat([g.reshape(-1) for g in grads])
    with _Timed('all_reduce_gradients', flat.numel() * flat.element_size()):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off, parts = 0, []
    for g in grads:
        n = g.numel()
        parts.append(flat[off:off + n].view(g.shape))
        off += n
    if flat.is_cuda and all(g.is_contiguous() for g in grads):
        ops.multi_copy(grads, parts)                       # one launch instead of one copy per parameter
    else:
        for g, p_ in zip(grads, parts):
            g.copy_(p_)


class PartitionedState:
    """Rank-local inputs of the training loop for one partitioned graph."""

    def __init__(self, partition, pairs, labels, knn, drug_feat, dis_feat, drug_sim_feat, dis_sim_feat, device):
        self.partition = partition
        self.enc_graph = PartitionedEncGraph.from_pairs(pairs, labels, partition, device)
        self.dec = PartitionedPairs(pairs, labels, partition, device)
        names = ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph')
        self.knn = {k: PartitionedAdjacency.from_sparse(knn[k], partition, 'drug' if k.startswith('drug') else 'disease')
                    for k in names}
        self.drug_feat = partition.rows(drug_feat, 'drug').contiguous()
        self.dis_feat = partition.rows(dis_feat, 'disease').contiguous()
        self.drug_sim_feat = partition.rows(drug_sim_feat, 'drug').contiguous()
        self.dis_sim_feat = partition.rows(dis_sim_feat, 'disease').contiguous()


def partitioned_loss(model, state, enc_graph, knn, feats, beta=0.001):
    """Forward of the drop-in Net on rank-local inputs + this rank's share L_p of the global loss."""
    drug_feat, dis_feat, drug_sim, dis_sim = feats
    pred, drug_out, drug_sim_out, dis_out, dis_sim_out = model(
        enc_graph, state.dec, knn['drug_graph'], drug_sim, drug_feat, knn['disease_graph'], dis_sim, dis_feat,
        knn['drug_feature_graph'], knn['disease_feature_graph'])
    part = state.partition
    bce = F.binary_cross_entropy_with_logits(pred.squeeze(-1), state.dec.labels, reduction='sum') / float(state.dec.n_global)
    c_drug, c_dis = common_losses_partitioned([(drug_out, drug_sim_out, part.n['drug']), (dis_out, dis_sim_out, part.n['disease'])])
    common = c_drug + c_dis
    return bce + beta * common / float(part.world), bce, common


def train_iteration_partitioned(model, optimizer, state, beta=0.001, grad_clip=1.0, edge_dropout_rate=0.1,
                                feature_noise_scale=0.05, augment=True):
    """train.py:250-300 on a row-partitioned graph. Returns the global loss (same value on every rank)."""
    model.train()
    ops.begin_seed_pool(state.drug_feat.device)       # this iteration's dropout seeds: one launch
    if augment:
        enc = state.enc_graph.edge_dropout(edge_dropout_rate)
        knn = {k: a.edge_dropout(edge_dropout_rate) for k, a in state.knn.items()}
        noise = lambda x, s: th.add(x, th.randn_like(x), alpha=s)
        feats = (noise(state.drug_feat, feature_noise_scale), noise(state.dis_feat, feature_noise_scale),
                 noise(state.drug_sim_feat, 0.05), noise(state.dis_sim_feat, 0.05))
    else:
        enc, knn = state.enc_graph, state.knn
        feats = (state.drug_feat, state.dis_feat, state.drug_sim_feat, state.dis_sim_feat)
    local, bce, common = partitioned_loss(model, state, enc, knn, feats, beta)
    optimizer.zero_grad()
    local.backward()
    params = [p for p in model.parameters()]
    all_reduce_gradients(params)
    from .train import clip_and_step
    clip_and_step(model, optimizer, grad_clip)
    with th.no_grad():
        total = all_reduce_sum(bce.detach()) + beta * common.detach()
    return total


class GraphedPartitionedIteration:
    """One row-partitioned training iteration -- augmentation, forward with its NCCL all-gathers, loss, backward with the
    reduce-scatters, gradient all-reduce, clip, Adam -- captured into one CUDA graph per rank and replayed. The eager loop
    enqueues ~900 launches and ~60 collectives per step from Python (~75 ms per step at 2 GPUs, where the device work is
    under 40 ms); every shape of the step is static and every random draw comes from the device generator, so the whole
    step records. NCCL's collectives are captured like any other kernel on the capture stream (the watchdog thread is kept
    out by `capture_error_mode='thread_local'`); all ranks capture and replay in lockstep.

    The optimizer must be capturable (`torch.optim.Adam(..., capturable=True)`) and the model must not have run on the
    legacy default stream (make a side stream current first, as bench.py does)."""

    def __init__(self, model, optimizer, state, warmup=3, parallel_branches=None, **step_kwargs):
        if not all(g.get('capturable', False) for g in optimizer.param_groups):
            raise ValueError('the optimizer must be built with capturable=True')
        self._step = lambda: train_iteration_partitioned(model, optimizer, state, **step_kwargs)
        # The two node types (and the GCMC / FGCN routes) as parallel stream branches of the graph: every collective blocks
        # only the branch that needs it, so an all-gather (forward) or reduce-scatter (backward: autograd replays each node
        # on its forward stream) of one branch runs under the other branch's projection / SpMM. The issue order of the
        # collectives is the host order of the branches -- the same on every rank.
        if parallel_branches is None:
            import os
            parallel_branches = os.environ.get('DG_ROWS_BRANCHES', '1') != '0'
        self.parallel_branches = bool(parallel_branches) and hasattr(model, 'parallel_routes')
        if self.parallel_branches:
            model.parallel_routes = True
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):                       # eager warm-up: cached transposes, NCCL communicators, cuBLAS handles
            for _ in range(max(warmup, 1)):
                self._step()
        th.cuda.current_stream().wait_stream(side)
        th.cuda.synchronize()
        if _world() > 1:
            dist.barrier()
        optimizer.zero_grad(set_to_none=True)
        self.graph = th.cuda.CUDAGraph()
        with th.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.loss = self._step()
        ops.drop_seed_pool()                               # its buffer now belongs to the graph
        if self.parallel_branches:
            model.parallel_routes = False                  # the branches are in the graph; eager calls stay serial
        th.cuda.synchronize()

    def __call__(self):
        self.graph.replay()
        return self.loss
