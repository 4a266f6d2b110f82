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
SELECT_MIN_EDGES = 1 << 20


def _randperm(n, device):
    """The random order whose first `num_keep` entries are kept (augmentation.py:51, :117).

    Eager: `th.randperm` on the graph's device, exactly as the reference -- the kept sets are the reference's for a
    given generator state. Inside a CUDA-graph capture (where the generator's offsets already differ from an eager
    run, and torch's small-n randperm, drawn on the CPU, cannot be recorded at all) only the kept SET is needed: an
    `ops.RandomSubset` marker makes `ops.keep_flags` draw a uniformly random num_keep-subset with a sort-free radix
    select -- the same distribution as randperm[:num_keep] -- for relations of at least SELECT_MIN_EDGES edges.
    DG_EDGE_SAMPLER=randperm / select forces either."""
    mode = os.environ.get('DG_EDGE_SAMPLER', 'auto')
    on_cuda = th.device(device).type == 'cuda'
    capturing = on_cuda and th.cuda.is_current_stream_capturing()
    if on_cuda and (mode == 'select' or (mode == 'auto' and capturing and n >= SELECT_MIN_EDGES)):
        return ops.RandomSubset(n)
    if n < 30000 and capturing:
        return th.argsort(th.rand(n, device=device))
    return th.randperm(n, device=device)


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
            pos = th.empty(int(base.eid.max()) + 1 if base.nnz else 1, dtype=th.int32, device=base.device)
            pos[base.eid.long()] = th.arange(base.nnz, dtype=th.int32, device=base.device)
            t = base.transpose()
            nb = ops.CSR(base.indptr, base.indices, pos[base.eid.long()], base.vals, base.n_rows, base.n_cols)
            nt = ops.CSR(t.indptr, t.indices, pos[t.eid.long()], t.vals, t.n_rows, t.n_cols)
            nb._t, nt._t = nt, nb
            nb.slot_order = nb.eid_is_slot = True
            sparse_graph._dg_csr = base = nb
        n = base.nnz
        k = num_keep_edges(n, dropout_rate)
        perm = _randperm(n, sparse_graph.device)
        flags = ops.keep_flags(n, [(perm, k, 0)], base.device)
        return _sparse_from_csr(ops.csr_dropout(base, flags, k), sparse_graph.shape)

    @staticmethod
    def add_random_edges(graph, add_rate=0.05, self_loops=False):
        """augmentation.py:127-205: add max(1, int(E*rate)) new, distinct, not-yet-present edges per
        relation. Candidates are drawn on device in batches; the first `num_add` valid ones win."""
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
            src, dst = out.edges(etype=c)
            have = src.long() * n_dst + dst.long()
            new = th.empty(0, dtype=th.int64, device=device)
            attempts, max_attempts = 0, num_add * 10               # augmentation.py:176
            while new.numel() < num_add and attempts < max_attempts:
                m = min(max(2 * (num_add - new.numel()), 64), max_attempts - attempts)
                cs = th.randint(0, n_src, (m,), device=device)
                cd = th.randint(0, n_dst, (m,), device=device)
                attempts += m
                ok = th.ones(m, dtype=th.bool, device=device)
                if not self_loops and st == dt:
                    ok &= cs != cd
                key = cs * n_dst + cd
                ok &= ~th.isin(key, have) & ~th.isin(key, new)
                key = key[ok]
                # first occurrence wins, draw order kept
                uniq, inv = th.unique(key, return_inverse=True)
                first = th.full((uniq.numel(),), key.numel(), dtype=th.int64, device=device)
                first.scatter_reduce_(0, inv, th.arange(key.numel(), device=device), reduce='amin')
                new = th.cat([new, key[th.sort(first).values]])[:num_add]
            if new.numel():
                out.add_edges(new // n_dst, new % n_dst, etype=c)
        return out

    @staticmethod
    def feature_noise(features, noise_scale=0.1):
        """augmentation.py:208-241."""
        if features is None:
            return None
        if not isinstance(features, th.Tensor):
            features = th.tensor(features, dtype=th.float32, device='cuda')
        return th.add(features, th.randn_like(features), alpha=noise_scale)      # features + noise * scale, one pass

    @staticmethod
    def sparse_graph_noise(graph, noise_scale=0.05):
        """augmentation.py:244-273: noise on the stored values, clamped at 0; structure unchanged."""
        if not isinstance(graph, th.Tensor) or not graph.is_sparse:
            return graph
        values = graph._values()
        noisy = th.clamp(values + th.randn_like(values) * noise_scale, min=0.0)
        out = th.sparse_coo_tensor(graph._indices(), noisy, graph.shape, device=graph.device, check_invariants=False)
        base = adjacency_csr(graph)
        vals = noisy.contiguous() if base.slot_order else noisy[base.eid.long()].contiguous()
        csr = ops.CSR(base.indptr, base.indices, base.eid, vals, base.n_rows, base.n_cols)
        csr.slot_order = base.slot_order       # transpose is rebuilt on demand with the new values
        out._dg_csr = csr
        return out

    @staticmethod
    def feature_masking(features, mask_rate=0.1):
        """augmentation.py:276-308."""
        if features is None:
            return None
        if not isinstance(features, th.Tensor):
            features = th.tensor(features, dtype=th.float32, device='cuda')
        return features * (th.rand_like(features) > mask_rate)

    @staticmethod
    def mix_up_features(features, alpha=0.2):
        """augmentation.py:311-337."""
        if features is None or not isinstance(features, th.Tensor):
            return features
        indices = th.randperm(features.size(0), device=features.device)
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
