"""Per-iteration graph augmentation on device -- mirror of the reference's `augmentation.py`
(`GraphAugmentation` statics at augmentation.py:13-337, `augment_graph_data` at :402-530).

Random sources are kept exactly where the reference has them (`th.randperm` per canonical etype in
sorted order, then per kNN graph, then `th.randn_like` per feature tensor), so the kept-edge sets
are the reference's for the same device generator state. What changes is everything after the
draw: instead of indexing COO lists and rebuilding a DGL heterograph (+ lazy CSR/CSC inside DGL),
the kept set becomes a flag array and the cached canonical CSRs (forward and transposed) are
compacted in order by CUDA kernels -- no sort, no host synchronisation.

Differences that are deliberate (documented in DESIGN.md):
  * kept edges come back in canonical CSR order rather than permutation order (same set);
  * un-augmented entries of the returned dict alias the inputs (the reference deep-copies them);
  * failures raise instead of being printed and swallowed (augmentation.py:446-447 etc.);
  * add_random_edges draws candidates with torch's device generator, not Python's `random`.
"""
import os

import numpy as np
import torch as th

from . import ops
from .graph import HeteroGraph
from .layers import adjacency_csr


def num_keep_edges(num_edges, dropout_rate):
    """augmentation.py:48 / :113 -- evaluated in Python double arithmetic like the reference."""
    return max(1, int(num_edges * (1 - dropout_rate)))


# below this the select buys nothing measurable: its 17 small launches against a sort of a few hundred thousand pairs
# (A/B on one box: lrssl 2.46 ms / iteration with randperm vs 2.64 with the select, Gdataset 1.93 vs 1.79 -- inside the
# run-to-run spread of these latency-chain iterations), so the sampler whose kept sets match the reference's stays
SELECT_MIN_EDGES = int(os.environ.get('DG_SELECT_MIN_EDGES', str(1 << 20)))
# ... and at or below this the select is ONE single-CTA launch (select_small_kernel) where the sort is rand + argsort's passes
SELECT_SMALL_EDGES = int(os.environ.get('DG_SELECT_SMALL_EDGES', str(1 << 14)))


def _randperm(n, device):
    """The random order whose first `num_keep` entries are kept (augmentation.py:51, :117).

    Eager: `th.randperm` on the graph's device, exactly as the reference -- the kept sets are the reference's for a
    given generator state. Inside a CUDA-graph capture (where the generator's offsets already differ from an eager
    run, and torch's small-n randperm, drawn on the CPU, cannot be recorded at all) only the kept SET is needed: an
    `ops.RandomSubset` marker makes `ops.keep_flags` draw a uniformly random num_keep-subset with a sort-free radix
    select -- the same distribution as randperm[:num_keep] -- for relations of at least SELECT_MIN_EDGES edges (its
    passes beat the sort's there) and of at most SELECT_SMALL_EDGES (one single-CTA launch).
    DG_EDGE_SAMPLER=randperm / select forces either."""
    mode = os.environ.get('DG_EDGE_SAMPLER', 'auto')
    on_cuda = th.device(device).type == 'cuda'
    capturing = on_cuda and th.cuda.is_current_stream_capturing()
    if on_cuda and (mode == 'select' or (mode == 'auto' and capturing and (n >= SELECT_MIN_EDGES or n <= SELECT_SMALL_EDGES))):
        return ops.RandomSubset(n)
    if n < 30000 and capturing:
        return th.argsort(th.rand(n, device=device))
    return th.randperm(n, device=device)


def _coo_position(csr, base):
    """For every slot of `csr` (the base CSR of a COO tensor or its transpose): the position of its entry in the tensor's
    COO arrays. eid is that position, except for a tensor listed in slot order of `base` whose eids still name a parent's
    entries (output of an edge dropout): there the forward CSR's slots are the positions and the transpose finds them
    through the shared eid."""
    if base.eid_is_slot or not base.slot_order:
        return csr.eid
    pos = th.empty(max(getattr(base, 'parent_nnz', 0), 1), dtype=th.int32, device=base.device)
    pos[base.eid.long()] = th.arange(base.nnz, dtype=th.int32, device=base.device)
    return pos[csr.eid.long()]


def _sparse_from_csr(csr, shape, bwd=None):
    idx = th.stack([csr.rows().long(), csr.indices.long()])
    t = th.sparse_coo_tensor(idx, csr.vals, shape, device=csr.device, check_invariants=False)
    csr.slot_order = True          # this tensor's COO arrays are in CSR slot order
    t._dg_csr = csr
    return t


class GraphAugmentation:
    @staticmethod
    def random_edge_dropout(graph, dropout_rate=0.1):
        """augmentation.py:13-89."""
        if not isinstance(graph, HeteroGraph):
            return graph
        device = graph.device
        perms = {}
        for c in graph.canonical_etypes:
            n = graph.number_of_edges(c)
            if n == 0:
                continue
            perms[c] = (_randperm(n, device), num_keep_edges(n, dropout_rate))
        return graph.edge_dropout(perms)

    @staticmethod
    def random_edge_dropout_sparse(sparse_graph, dropout_rate=0.1):
        """augmentation.py:92-124."""
        if not isinstance(sparse_graph, th.Tensor) or not sparse_graph.is_sparse:
            return sparse_graph
        base = adjacency_csr(sparse_graph)
        if base.slot_order and not base.eid_is_slot:
            # output of an earlier dropout: its edge ids still name the grand-parent's entries;
            # renumber to this tensor's own COO positions (= slots) before drawing a new perm
            # (sized from the parent's entry count kept on the CSR -- no device read)
            pos = th.empty(max(getattr(base, 'parent_nnz', 0), base.nnz, 1), dtype=th.int32, device=base.device)
            pos[base.eid.long()] = th.arange(base.nnz, dtype=th.int32, device=base.device)
            t = base.transpose()
            nb = ops.inherit_layout(ops.CSR(base.indptr, base.indices, pos[base.eid.long()], base.vals, base.n_rows, base.n_cols), base)
            nt = ops.inherit_layout(ops.CSR(t.indptr, t.indices, pos[t.eid.long()], t.vals, t.n_rows, t.n_cols), t)
            nb._t, nt._t = nt, nb
            nb.slot_order = nb.eid_is_slot = True
            sparse_graph._dg_csr = base = nb
        n = base.nnz
        k = num_keep_edges(n, dropout_rate)
        perm = _randperm(n, sparse_graph.device)
        flags = ops.keep_flags(n, [(perm, k, 0)], base.device)
        dropped = ops.csr_dropout(base, flags, k)
        dropped.parent_nnz = dropped._t.parent_nnz = n            # eid values of the result name entries of this tensor
        return _sparse_from_csr(dropped, sparse_graph.shape)

    @staticmethod
    def add_random_edges(graph, add_rate=0.05, self_loops=False, candidates=None):
        """augmentation.py:127-205: per relation add num_add = max(1, int(E * rate)) new edges -- the first num_add
        candidates, in draw order, among at most 10 * num_add uniform (src, dst) draws that are neither existing edges nor
        repeats of an earlier accepted candidate (self-loops excluded when both endpoints are of one node type).

        The reference walks the draws one by one in Python against a `set` of all existing edges; here all 10 * num_add
        candidates are drawn at once on the device and filtered with three sorts / searches -- existing-edge membership by
        binary search in the relation's sorted edge keys, first occurrence by a stable sort of the candidate keys, rank in
        draw order by a prefix sum -- which selects exactly the edges the sequential walk accepts for the same draws. One
        scalar (how many were accepted; it sizes the new edge list) is read back per relation: no host loop.
        `candidates` = {canonical etype: (src, dst)} injects the draws (tests); default: torch's device generator (the
        reference uses Python's `random`)."""
        if not isinstance(graph, HeteroGraph):
            return graph
        out = graph.clone()
        device = out.device
        for c in out.canonical_etypes:
            st, _, dt = c
            n = out.number_of_edges(c)
            n_src, n_dst = out.number_of_nodes(st), out.number_of_nodes(dt)
            if n == 0 or n_src == 0 or n_dst == 0:
                continue
            num_add = max(1, int(n * add_rate))
            m = num_add * 10                                       # augmentation.py:176
            src, dst = out.edges(etype=c)
            have = th.sort(src.long() * n_dst + dst.long()).values
            if candidates is not None and c in candidates:
                cs, cd = (th.as_tensor(x, device=device).long()[:m] for x in candidates[c])
            else:
                cs = th.randint(0, n_src, (m,), device=device)
                cd = th.randint(0, n_dst, (m,), device=device)
            key = cs * n_dst + cd
            pos = th.searchsorted(have, key).clamp_(max=n - 1)
            ok = have[pos] != key                                  # not an existing edge
            if not self_loops and st == dt:
                ok &= cs != cd
            # first occurrence among the candidates that could be accepted at all (a rejected draw does not block a later
            # identical one -- it would be rejected again for the same reason)
            order = th.sort(th.where(ok, key, th.full_like(key, -1)), stable=True).indices
            skey = key[order]
            dup = th.zeros_like(ok)
            dup[order[1:]] = (skey[1:] == skey[:-1]) & ok[order[1:]] & ok[order[:-1]]
            ok &= ~dup
            rank = th.cumsum(ok.to(th.int64), 0)
            take = ok & (rank <= num_add)
            new = key[take]                                        # draw order; the one device -> host read sizes it
            if new.numel():
                out.add_edges(th.div(new, n_dst, rounding_mode='floor'), new % n_dst, etype=c)
        return out

    @staticmethod
    def feature_noise(features, noise_scale=0.1, noise=None):
        """augmentation.py:208-241. `noise` injects the N(0,1) draw (tests)."""
        if features is None:
            return None
        if not isinstance(features, th.Tensor):
            features = th.tensor(features, dtype=th.float32, device='cuda')
        noise = th.randn_like(features) if noise is None else noise
        return th.add(features, noise, alpha=noise_scale)                        # features + noise * scale, one pass

    @staticmethod
    def sparse_graph_noise(graph, noise_scale=0.05, noise=None):
        """augmentation.py:244-273: noise on the stored values, clamped at 0; structure unchanged -- so the CSR and its
        cached transpose are reused as they are and only the two value arrays are regathered (no sort).
        `noise` injects the N(0,1) draw (tests)."""
        if not isinstance(graph, th.Tensor) or not graph.is_sparse:
            return graph
        values = graph._values()
        noise = th.randn_like(values) if noise is None else noise
        noisy = th.clamp(values + noise * noise_scale, min=0.0)
        out = th.sparse_coo_tensor(graph._indices(), noisy, graph.shape, device=graph.device, check_invariants=False)
        base = adjacency_csr(graph)
        # position of slot s's entry in this tensor's COO arrays: s itself when the COO is in slot order, else eid[s] for
        # a base graph (eid = COO position); a dropped graph's COO is always in slot order (_sparse_from_csr)
        def regather(c):
            return noisy.contiguous() if (c is base and base.slot_order) else noisy[_coo_position(c, base).long()].contiguous()
        csr = ops.inherit_layout(ops.CSR(base.indptr, base.indices, base.eid, regather(base), base.n_rows, base.n_cols), base)
        csr.slot_order, csr.eid_is_slot, csr.parent_nnz = base.slot_order, base.eid_is_slot, base.parent_nnz
        if base._t is not None:
            t = base._t
            ct = ops.inherit_layout(ops.CSR(t.indptr, t.indices, t.eid, regather(t), t.n_rows, t.n_cols), t)
            ct.slot_order, ct.eid_is_slot, ct.parent_nnz = t.slot_order, t.eid_is_slot, t.parent_nnz
            csr._t, ct._t = ct, csr
        out._dg_csr = csr
        return out

    @staticmethod
    def feature_masking(features, mask_rate=0.1, u=None):
        """augmentation.py:276-308. `u` injects the U(0,1) draw (tests)."""
        if features is None:
            return None
        if not isinstance(features, th.Tensor):
            features = th.tensor(features, dtype=th.float32, device='cuda')
        u = th.rand_like(features) if u is None else u
        return features * (u > mask_rate)

    @staticmethod
    def mix_up_features(features, alpha=0.2, indices=None, lam=None):
        """augmentation.py:311-337. `indices` / `lam` inject the permutation and the Beta(alpha, alpha) draw (tests)."""
        if features is None or not isinstance(features, th.Tensor):
            return features
        if indices is None:
            indices = th.randperm(features.size(0), device=features.device)
        if lam is None:
            lam = np.random.beta(alpha, alpha)
        return lam * features + (1 - lam) * features[indices]


_GRAPH_KEYS = ('enc_graph', 'dec_graph')
_SPARSE_KEYS = ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph')


def augment_graph_data(graph_data, aug_methods=None, aug_params=None):
    """augmentation.py:402-530 -- same keys, same method names, same parameter names and defaults."""
    if aug_methods is None:
        aug_methods = ['edge_dropout']
    if aug_params is None:
        aug_params = {'edge_dropout_rate': 0.1}
    out = dict(graph_data)
    A = GraphAugmentation
    for method in aug_methods:
        if method == 'edge_dropout':
            rate = aug_params.get('edge_dropout_rate', 0.1)
            for k in _GRAPH_KEYS:
                if isinstance(out.get(k), HeteroGraph):
                    out[k] = A.random_edge_dropout(out[k], rate)
            for k in _SPARSE_KEYS:
                v = out.get(k)
                if isinstance(v, th.Tensor) and v.is_sparse:
                    out[k] = A.random_edge_dropout_sparse(v, rate)
        elif method == 'add_random_edges':
            rate = aug_params.get('add_edge_rate', 0.05)
            for k in _GRAPH_KEYS:
                if isinstance(out.get(k), HeteroGraph):
                    out[k] = A.add_random_edges(out[k], rate)
        elif method == 'feature_noise':
            fs = aug_params.get('feature_noise_scale', 0.1)
            ss = aug_params.get('sim_noise_scale', 0.05)
            for k, s in (('drug_feat', fs), ('disease_feat', fs), ('drug_sim_feat', ss), ('disease_sim_feat', ss)):
                if out.get(k) is not None:
                    out[k] = A.feature_noise(out[k], s)
        elif method == 'graph_noise':
            s = aug_params.get('graph_noise_scale', 0.05)
            for k in _SPARSE_KEYS:
                v = out.get(k)
                if isinstance(v, th.Tensor) and v.is_sparse:
                    out[k] = A.sparse_graph_noise(v, s)
        elif method == 'feature_masking':
            r = aug_params.get('feature_mask_rate', 0.1)
            for k in ('drug_feat', 'disease_feat'):
                if out.get(k) is not None:
                    out[k] = A.feature_masking(out[k], r)
        elif method == 'mix_up':
            a = aug_params.get('mixup_alpha', 0.2)
            for k in ('drug_feat', 'disease_feat'):
                if out.get(k) is not None:
                    out[k] = A.mix_up_features(out[k], a)
    return out
