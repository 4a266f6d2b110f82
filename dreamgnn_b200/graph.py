"""Graph handle that replaces `dgl.DGLGraph` on DREAM-GNN's hot path.

It offers the DGL-surface subset the reference actually uses (constructor from
`{(src_type, etype, dst_type): (src, dst)}` + `num_nodes_dict`, `bipartite_from_scipy`, `.etypes`,
`.canonical_etypes` (sorted), `.ntypes`, `number_of_nodes/edges/src_nodes`, `.edges(etype=)`,
`.edges[et].data`, `.nodes[nt].data`, `g[etype]` slices sharing node data, `in/out_degrees`,
`srcdata/dstdata/edata`, `local_scope()`, `update_all(copy_u, sum)`, `apply_edges(udf)`, `.int()`,
`.to()`, `.clone()`, `.device`, `add_edges` -- data_loader.py:448-509, layers.py:174-233, 361-365,
augmentation.py:24-89, 139-205) and, underneath, owns per-destination-type *relation blocks*: one
canonical CSR (rows = destination nodes, columns = `r * N_src + src`) over all R relations that feed a
node type plus its transpose, int32 on device, built lazily by the CUDA kernels and cached. The fused
GCMC layer aggregates all relations of a block in one SpMM launch.

Message passing needs CUDA tensors; structure bookkeeping also works on CPU tensors so that the
loader-side code can run before `.to(device)`.
"""
import contextlib

import numpy as np
import torch as th

from . import ops


class DGLError(Exception):
    """Mirror of dgl.DGLError (layers.py:216)."""


class _CopyU:
    def __init__(self, u, out):
        self.u, self.out = u, out


class _Sum:
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


class function:  # noqa: N801  -- mirrors the `dgl.function` namespace (layers.py:229-232)
    @staticmethod
    def copy_u(u, out):
        return _CopyU(u, out)

    copy_src = copy_u

    @staticmethod
    def sum(msg, out):
        return _Sum(msg, out)


def _as_index(x, dtype, device=None):
    if isinstance(x, th.Tensor):
        t = x.to(dtype)
    else:
        t = th.as_tensor(np.asarray(x), dtype=dtype)
    return t.to(device) if device is not None else t


class RelBlock:
    """All relations into one destination node type as a single sparse operator.

    rows = destination nodes; columns = r * N_src + src_node for relation r, i.e. the messages of all
    relations live in one [R, N_src, D] buffer gathered as [R*N_src, D]. Relation-major (rather than
    interleaved) keeps the rows of the TRANSPOSED block grouped by relation, so the warps of a CTA
    see similar row lengths in the backward SpMM (rel "1" rows are ~100x shorter than rel "0" rows).
    Edge id of edge e of relation r = offsets[r] + e.
    """

    def __init__(self, etypes, src_type, dst_type, n_src, n_dst, csr, offsets):
        self.etypes, self.src_type, self.dst_type = list(etypes), src_type, dst_type
        self.n_src, self.n_dst, self.csr, self.offsets = n_src, n_dst, csr, list(offsets)

    @property
    def num_rel(self):
        return len(self.etypes)


class _TypedView:
    def __init__(self, data):
        self.data = data


class _NodeAccessor:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        return _TypedView(self._g._ndata[ntype])


class _EdgeAccessor:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, etype):
        return _TypedView(self._g._edata[self._g.to_canonical_etype(etype)])

    def __call__(self, etype=None, form='uv', order='eid'):
        return self._g._get_edges(self._g.to_canonical_etype(etype))


class _Gathered:
    def __init__(self, store, index):
        self._store, self._index = store, index

    def __getitem__(self, key):
        return self._store[key][self._index.long()]


class _EdgeBatch:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class HeteroGraph:
    def __init__(self, edges, num_nodes, ndata=None, edata=None, idtype=th.int64, blocks=None):
        self._edges = dict(sorted(edges.items()))        # canonical etypes sorted like DGL
        self._num_nodes = dict(num_nodes)
        self._ndata = ndata if ndata is not None else {nt: {} for nt in self._num_nodes}
        self._edata = edata if edata is not None else {c: {} for c in self._edges}
        self.idtype = idtype
        self._blocks = blocks if blocks is not None else {}
        self._csr = {}
        self._derived = {}          # .int() / .to(device) results, so repeated conversions keep their CSR caches
        self.nodes = _NodeAccessor(self)
        self.edges = _EdgeAccessor(self)

    # ---- schema ---------------------------------------------------------------------------------
    @property
    def canonical_etypes(self):
        return list(self._edges.keys())

    @property
    def etypes(self):
        return [c[1] for c in self._edges]

    @property
    def ntypes(self):
        return sorted(self._num_nodes.keys())

    @property
    def srctypes(self):
        return sorted({c[0] for c in self._edges})

    @property
    def dsttypes(self):
        return sorted({c[2] for c in self._edges})

    def _get_edges(self, c):
        e = self._edges[c]
        if callable(e):                                   # lazily materialised after edge dropout
            e = e()
            self._edges[c] = e
        return e

    @property
    def device(self):
        for c in self._edges:
            e = self._edges[c]
            if not callable(e):
                return e[0].device
        for b in self._blocks.values():
            return b.csr.device
        return th.device('cpu')

    def to_canonical_etype(self, etype):
        if etype is None:
            if len(self._edges) != 1:
                raise DGLError('Edge type name must be specified if there are more than one edge types.')
            return next(iter(self._edges))
        if isinstance(etype, tuple):
            return etype
        hits = [c for c in self._edges if c[1] == etype]
        if len(hits) != 1:
            raise DGLError('Edge type "%s" is ambiguous or does not exist.' % (etype,))
        return hits[0]

    def _one_ntype(self, which):
        if len(self._edges) == 1:
            c = next(iter(self._edges))
            return c[0] if which == 'src' else c[2]
        if len(self._num_nodes) == 1:
            return next(iter(self._num_nodes))
        raise DGLError('Node type name must be specified if there are more than one node types.')

    # ---- sizes ----------------------------------------------------------------------------------
    def number_of_nodes(self, ntype=None):
        if ntype is None:
            return sum(self._num_nodes.values())
        return self._num_nodes[ntype]

    num_nodes = number_of_nodes

    def number_of_edges(self, etype=None):
        if etype is None and len(self._edges) != 1:
            return sum(self.number_of_edges(c) for c in self._edges)
        c = self.to_canonical_etype(etype)
        e = self._edges[c]
        if callable(e):
            return e.count
        return int(e[0].numel())

    num_edges = number_of_edges

    def number_of_src_nodes(self, ntype=None):
        return self._num_nodes[ntype or self._one_ntype('src')]

    def number_of_dst_nodes(self, ntype=None):
        return self._num_nodes[ntype or self._one_ntype('dst')]

    num_src_nodes, num_dst_nodes = number_of_src_nodes, number_of_dst_nodes

    def in_degrees(self, etype=None):
        c = self.to_canonical_etype(etype)
        return th.bincount(self._get_edges(c)[1].long(), minlength=self._num_nodes[c[2]]).to(self.idtype)

    def out_degrees(self, etype=None):
        c = self.to_canonical_etype(etype)
        return th.bincount(self._get_edges(c)[0].long(), minlength=self._num_nodes[c[0]]).to(self.idtype)

    # ---- relation slices ------------------------------------------------------------------------
    def __getitem__(self, key):
        c = self.to_canonical_etype(key)
        nn_ = {c[0]: self._num_nodes[c[0]], c[2]: self._num_nodes[c[2]]}
        g = HeteroGraph({c: self._edges[c]}, nn_, {nt: self._ndata[nt] for nt in nn_}, {c: self._edata[c]},
                        self.idtype)
        g._csr = self._csr                                 # share the per-etype CSR cache
        g._parent = self
        return g

    # ---- feature storage ------------------------------------------------------------------------
    @property
    def srcdata(self):
        return self._ndata[self._one_ntype('src')]

    @property
    def dstdata(self):
        return self._ndata[self._one_ntype('dst')]

    @property
    def edata(self):
        return self._edata[self.to_canonical_etype(None)]

    @contextlib.contextmanager
    def local_scope(self):
        saved_n = {nt: dict(d) for nt, d in self._ndata.items()}
        saved_e = {c: dict(d) for c, d in self._edata.items()}
        try:
            yield
        finally:
            for nt, d in self._ndata.items():
                d.clear()
                d.update(saved_n[nt])
            for c, d in self._edata.items():
                d.clear()
                d.update(saved_e[c])

    # ---- sparse structure -----------------------------------------------------------------------
    def etype_csr(self, etype=None):
        """CSR by destination of one relation (rows = dst nodes, columns = src nodes)."""
        c = self.to_canonical_etype(etype)
        if c not in self._csr:
            src, dst = self._get_edges(c)
            if not src.is_cuda:
                raise RuntimeError('message passing needs the graph on a CUDA device (no CPU fallback)')
            self._csr[c] = ops.CSR.from_coo(dst, src, self._num_nodes[c[2]], self._num_nodes[c[0]])
        return self._csr[c]

    def block(self, dst_type):
        """Relation block of every etype whose destination is `dst_type` (cached)."""
        if dst_type not in self._blocks:
            cets = [c for c in self._edges if c[2] == dst_type]
            if not cets:
                raise DGLError('no relation into node type %r' % (dst_type,))
            src_type = cets[0][0]
            if any(c[0] != src_type for c in cets):
                raise DGLError('relation block needs a single source node type')
            R = len(cets)
            rows, cols, offsets, off = [], [], [], 0
            for r, c in enumerate(cets):
                src, dst = self._get_edges(c)
                if not src.is_cuda:
                    raise RuntimeError('message passing needs the graph on a CUDA device (no CPU fallback)')
                rows.append(dst.to(th.int32))
                cols.append(src.to(th.int32) + r * self._num_nodes[src_type])
                offsets.append(off)
                off += int(src.numel())
            n_src, n_dst = self._num_nodes[src_type], self._num_nodes[dst_type]
            csr = ops.CSR.from_coo(th.cat(rows), th.cat(cols), n_dst, n_src * R)
            self._blocks[dst_type] = RelBlock(cets, src_type, dst_type, n_src, n_dst, csr, offsets)
        return self._blocks[dst_type]

    # ---- message passing (generic DGL-style entry points) ---------------------------------------
    def update_all(self, message_func, reduce_func, etype=None):
        if not (isinstance(message_func, _CopyU) and isinstance(reduce_func, _Sum)):
            raise DGLError('only update_all(copy_u, sum) is implemented (layers.py:229-232)')
        c = self.to_canonical_etype(etype)
        h = self._ndata[c[0]][message_func.u]
        d = h.shape[1]
        if d % 4:                                          # the kernel gathers 16-byte vectors: zero-pad odd widths
            h = th.nn.functional.pad(h, (0, (-d) % 4))
        out = ops.spmm(self.etype_csr(c), h)
        self._ndata[c[2]][reduce_func.out] = out[:, :d] if out.shape[1] != d else out

    def apply_edges(self, func, etype=None):
        c = self.to_canonical_etype(etype)
        src, dst = self._get_edges(c)
        batch = _EdgeBatch(_Gathered(self._ndata[c[0]], src), _Gathered(self._ndata[c[2]], dst), self._edata[c])
        self._edata[c].update(func(batch))

    def pair_graph(self, etype=None):
        """Decoder view of a single-relation graph: pairs in edge order + segment structures."""
        c = self.to_canonical_etype(etype)
        key = ('pairs', c)
        if key not in self._csr:
            src, dst = self._get_edges(c)
            self._csr[key] = ops.PairGraph(src, dst, self._num_nodes[c[0]], self._num_nodes[c[2]])
        return self._csr[key]

    # ---- mutation / conversion ------------------------------------------------------------------
    def add_edges(self, u, v, data=None, etype=None):
        c = self.to_canonical_etype(etype)
        s, d = self._get_edges(c)
        self._edges[c] = (th.cat([s, _as_index(u, s.dtype, s.device)]), th.cat([d, _as_index(v, d.dtype, d.device)]))
        self._csr, self._blocks, self._derived = {}, {}, {}   # structure changed: drop cached CSRs

    def _map(self, idx_fn, feat_fn):
        edges = {c: tuple(idx_fn(t) for t in self._get_edges(c)) for c in self._edges}
        nd = {nt: {k: feat_fn(v) for k, v in d.items()} for nt, d in self._ndata.items()}
        ed = {c: {k: feat_fn(v) for k, v in d.items()} for c, d in self._edata.items()}
        return edges, nd, ed

    def clone(self):
        """Structure is immutable (add_edges rebinds), so clones share index tensors and CSR caches;
        node / edge feature tensors are copied like DGL's clone()."""
        nd = {nt: {k: v.clone() for k, v in d.items()} for nt, d in self._ndata.items()}
        ed = {c: {k: v.clone() for k, v in d.items()} for c, d in self._edata.items()}
        g = HeteroGraph(dict(self._edges), self._num_nodes, nd, ed, self.idtype, dict(self._blocks))
        g._csr = dict(self._csr)
        return g

    def to(self, device, **kwargs):
        device = th.device(device)
        if device.type == 'cuda' and device.index is None:
            device = th.device('cuda', th.cuda.current_device())
        if device == self.device:
            return self
        key = ('to', str(device))
        if key not in self._derived:
            e, nd, ed = self._map(lambda t: t.to(device), lambda t: t.to(device))
            self._derived[key] = HeteroGraph(e, self._num_nodes, nd, ed, self.idtype)
        return self._derived[key]

    def int(self):
        if self.idtype == th.int32:
            return self
        if 'int' not in self._derived:
            e, nd, ed = self._map(lambda t: t.to(th.int32), lambda t: t)
            self._derived['int'] = HeteroGraph(e, self._num_nodes, nd, ed, th.int32)
        return self._derived['int']

    def long(self):
        if self.idtype == th.int64:
            return self
        e, nd, ed = self._map(lambda t: t.to(th.int64), lambda t: t)
        return HeteroGraph(e, self._num_nodes, nd, ed, th.int64)

    def cpu(self):
        return self.to('cpu')

    def __repr__(self):
        return 'HeteroGraph(num_nodes=%r, num_edges=%r)' % (
            self._num_nodes, {c: self.number_of_edges(c) for c in self._edges})

    # ---- edge dropout (augmentation.py:13-89) -----------------------------------------------------
    def edge_dropout(self, perms):
        """New graph keeping, per canonical etype, the edges listed first in its permutation:
        `perms[c] = (perm int64 on device, num_keep)`. Built by compacting this graph's cached CSR blocks
        (forward and transposed) with one keep-flag array per block -- no sort, no host sync. Node
        data (ci / cj) is copied, NOT recomputed (augmentation.py:68-70)."""
        blocks, lazy = {}, {}
        for dt in self.dsttypes:
            base = self.block(dt)
            plist, n_keep = [], 0
            for c, off in zip(base.etypes, base.offsets):
                if c in perms:
                    perm, k = perms[c]
                    plist.append((perm, k, off))
                    n_keep += int(k)
            flags = ops.keep_flags(base.csr.nnz, plist, base.csr.device)
            csr = ops.csr_dropout(base.csr, flags, n_keep)
            blk = RelBlock(base.etypes, base.src_type, base.dst_type, base.n_src, base.n_dst, csr, base.offsets)
            blocks[dt] = blk
            for r, c in enumerate(base.etypes):
                lazy[c] = _LazyEdges(blk, r, int(perms[c][1]) if c in perms else 0, self.idtype)
        nd = {nt: {k: v.clone() for k, v in d.items()} for nt, d in self._ndata.items()}
        return HeteroGraph(lazy, self._num_nodes, nd, None, self.idtype, blocks)


class _LazyEdges:
    """(src, dst) of one relation of a dropped block, expanded from the CSR only if someone asks."""

    def __init__(self, block, r, count, idtype):
        self.block, self.r, self.count, self.idtype = block, r, count, idtype

    def __call__(self):
        csr, n_src = self.block.csr, self.block.n_src
        sel = th.div(csr.indices, n_src, rounding_mode='floor') == self.r
        return ((csr.indices[sel] - self.r * n_src).to(self.idtype), csr.rows()[sel].to(self.idtype))


DGLGraph = HeteroGraph
DGLHeteroGraph = HeteroGraph


def heterograph(data_dict, num_nodes_dict=None, idtype=None, device=None):
    """Mirror of dgl.heterograph (data_loader.py:448, 508; augmentation.py:65)."""
    want = idtype
    for c, (u, v) in data_dict.items():
        if want is None and isinstance(u, th.Tensor) and u.dtype in (th.int32, th.int64):
            want = u.dtype
    want = want or th.int64
    edges = {tuple(c): (_as_index(u, want, device), _as_index(v, want, device)) for c, (u, v) in data_dict.items()}
    if num_nodes_dict is None:
        num_nodes_dict = {}
        for (st, _, dt), (u, v) in edges.items():
            num_nodes_dict[st] = max(num_nodes_dict.get(st, 0), int(u.max()) + 1 if u.numel() else 0)
            num_nodes_dict[dt] = max(num_nodes_dict.get(dt, 0), int(v.max()) + 1 if v.numel() else 0)
    return HeteroGraph(edges, num_nodes_dict, idtype=want)


def bipartite_from_scipy(sp_mat, utype, etype, vtype, eweight_name=None, idtype=None, device=None):
    """Mirror of dgl.bipartite_from_scipy (data_loader.py:507): COO storage order is the edge order."""
    coo = sp_mat.tocoo()
    return heterograph({(utype, etype, vtype): (coo.row, coo.col)},
                       num_nodes_dict={utype: coo.shape[0], vtype: coo.shape[1]}, idtype=idtype, device=device)
