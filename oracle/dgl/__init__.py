"""TEST INFRASTRUCTURE ONLY -- pure-torch stand-in for the subset of DGL that DREAM-GNN touches.

The reference (`/root/reference/{layers,data_loader,augmentation}.py`) imports `dgl` at module
top, and DGL is neither pinned in its requirements.txt nor installable here (no network). This
package re-creates, in plain torch on CPU, the ~25 DGL entry points the reference calls so that
the reference files can be imported UNMODIFIED (put `oracle/` ahead of `/root/reference` on
`sys.path`). Reference + this stand-in is the correctness oracle and golden-vector generator.

PARITY UNPINNED: upstream DGL itself cannot be run here, so the DGL semantics below are
restated from its documented behaviour, not checked against the library:
  * `heterograph` keeps per-relation edge order (edge id = insertion order) and exposes
    `canonical_etypes` sorted lexicographically, `ntypes` sorted;
  * `update_all(copy_u, sum)` == `out.index_add_(0, dst, h[src])`;
  * `apply_edges(udf)` hands the UDF row-gathered src/dst features in edge order;
  * `in_degrees/out_degrees` == bincount over dst/src;
  * a relation slice `g[etype]` shares the parent's node-data dicts;
  * `local_scope()` restores node/edge data dicts on exit;
  * `bipartite_from_scipy` keeps the COO storage order of the scipy matrix;
  * `HeteroGraphConv` loops `canonical_etypes` and aggregates with stack->sum.

Nothing in the product package (`dreamgnn_b200/`) may import this module.
"""
import contextlib

import numpy as np
import torch as th

from . import function  # noqa: F401
from . import function as fn  # noqa: F401  (reference tries `import dgl.fn` first)

__version__ = "0.0-standin"


class DGLError(Exception):
    pass


def _as_index(x, dtype=th.int64, device=None):
    if isinstance(x, th.Tensor):
        t = x.to(dtype)
    else:
        t = th.as_tensor(np.asarray(x), dtype=dtype)
    return t.to(device) if device is not None else t


class _TypedView:
    """`g.nodes[ntype]` / `g.edges[etype]` -> object with a `.data` dict."""

    def __init__(self, data):
        self.data = data


class _NodeAccessor:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        return _TypedView(self._g._ndata[ntype])

    def __call__(self, ntype=None):
        return th.arange(self._g.number_of_nodes(ntype), dtype=self._g.idtype, device=self._g.device)


class _EdgeAccessor:
    """`g.edges(etype=...)` is callable and `g.edges[etype].data` is indexable."""

    def __init__(self, g):
        self._g = g

    def __getitem__(self, etype):
        return _TypedView(self._g._edata[self._g.to_canonical_etype(etype)])

    def __call__(self, etype=None, form='uv', order='eid'):
        c = self._g.to_canonical_etype(etype)
        return self._g._edges[c]


class _EdgeBatch:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class _Gathered:
    """Lazy dict: `edges.src['h']` -> node_data['h'][index]."""

    def __init__(self, store, index):
        self._store, self._index = store, index

    def __getitem__(self, key):
        return self._store[key][self._index.long()]


class DGLGraph:
    def __init__(self, edges, num_nodes, ndata=None, edata=None, idtype=th.int64):
        # edges: {canonical etype: (src, dst)}, already tensors of `idtype`
        self._edges = dict(sorted(edges.items()))
        self._num_nodes = dict(num_nodes)
        self._ndata = ndata if ndata is not None else {nt: {} for nt in self._num_nodes}
        self._edata = edata if edata is not None else {c: {} for c in self._edges}
        self.idtype = idtype
        self.nodes = _NodeAccessor(self)
        self.edges = _EdgeAccessor(self)

    # ---- schema -------------------------------------------------------------------------
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

    @property
    def device(self):
        for s, _ in self._edges.values():
            return s.device
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

    def _single_ntype(self, ntype, which):
        if ntype is not None:
            return ntype
        c = self.to_canonical_etype(None) if len(self._edges) == 1 else None
        if c is None:
            if len(self._num_nodes) == 1:
                return next(iter(self._num_nodes))
            raise DGLError('Node type name must be specified if there are more than one node types.')
        return c[0] if which == 'src' else c[2]

    # ---- sizes --------------------------------------------------------------------------
    def number_of_nodes(self, ntype=None):
        if ntype is None:
            if len(self._num_nodes) == 1:
                return next(iter(self._num_nodes.values()))
            return sum(self._num_nodes.values())
        return self._num_nodes[ntype]

    num_nodes = number_of_nodes

    def number_of_edges(self, etype=None):
        if etype is None and len(self._edges) != 1:
            return sum(int(s.numel()) for s, _ in self._edges.values())
        return int(self._edges[self.to_canonical_etype(etype)][0].numel())

    num_edges = number_of_edges

    def number_of_src_nodes(self, ntype=None):
        return self._num_nodes[self._single_ntype(ntype, 'src')]

    def number_of_dst_nodes(self, ntype=None):
        return self._num_nodes[self._single_ntype(ntype, 'dst')]

    num_src_nodes = number_of_src_nodes
    num_dst_nodes = number_of_dst_nodes

    def in_degrees(self, etype=None):
        c = self.to_canonical_etype(etype)
        return th.bincount(self._edges[c][1].long(), minlength=self._num_nodes[c[2]]).to(self.idtype)

    def out_degrees(self, etype=None):
        c = self.to_canonical_etype(etype)
        return th.bincount(self._edges[c][0].long(), minlength=self._num_nodes[c[0]]).to(self.idtype)

    # ---- relation slices ----------------------------------------------------------------
    def __getitem__(self, key):
        c = self.to_canonical_etype(key)
        nn_ = {c[0]: self._num_nodes[c[0]], c[2]: self._num_nodes[c[2]]}
        nd = {nt: self._ndata[nt] for nt in nn_}            # shared, not copied
        return DGLGraph({c: self._edges[c]}, nn_, nd, {c: self._edata[c]}, self.idtype)

    # ---- feature storage ----------------------------------------------------------------
    @property
    def srcdata(self):
        return self._ndata[self._single_ntype(None, 'src')]

    @property
    def dstdata(self):
        return self._ndata[self._single_ntype(None, 'dst')]

    @property
    def ndata(self):
        return self._ndata[self._single_ntype(None, 'src')]

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

    # ---- message passing ----------------------------------------------------------------
    def update_all(self, message_func, reduce_func, etype=None):
        c = self.to_canonical_etype(etype)
        src, dst = self._edges[c]
        if not (isinstance(message_func, function.CopyU) and isinstance(reduce_func, function.Sum)):
            raise DGLError('stand-in implements only update_all(copy_u, sum)')
        h = self._ndata[c[0]][message_func.u]
        if reduce_func.msg != message_func.out:
            raise DGLError('message field mismatch')
        out = th.zeros((self._num_nodes[c[2]],) + tuple(h.shape[1:]), dtype=h.dtype, device=h.device)
        out = out.index_add(0, dst.long(), h[src.long()])
        self._ndata[c[2]][reduce_func.out] = out

    def apply_edges(self, func, etype=None):
        c = self.to_canonical_etype(etype)
        src, dst = self._edges[c]
        batch = _EdgeBatch(_Gathered(self._ndata[c[0]], src), _Gathered(self._ndata[c[2]], dst),
                           self._edata[c])
        self._edata[c].update(func(batch))

    # ---- structure mutation / conversion ------------------------------------------------
    def add_edges(self, u, v, data=None, etype=None):
        c = self.to_canonical_etype(etype)
        s, d = self._edges[c]
        u = _as_index(u, s.dtype, s.device)
        v = _as_index(v, d.dtype, d.device)
        self._edges[c] = (th.cat([s, u]), th.cat([d, v]))
        for k, f in list(self._edata[c].items()):
            pad = th.zeros((u.numel(),) + tuple(f.shape[1:]), dtype=f.dtype, device=f.device)
            self._edata[c][k] = th.cat([f, pad])

    def _map(self, idx_fn, feat_fn):
        edges = {c: (idx_fn(s), idx_fn(d)) for c, (s, d) in self._edges.items()}
        nd = {nt: {k: feat_fn(v) for k, v in d.items()} for nt, d in self._ndata.items()}
        ed = {c: {k: feat_fn(v) for k, v in d.items()} for c, d in self._edata.items()}
        return edges, nd, ed

    def clone(self):
        e, nd, ed = self._map(lambda t: t.clone(), lambda t: t.clone())
        return DGLGraph(e, self._num_nodes, nd, ed, self.idtype)

    def to(self, device, **kwargs):
        e, nd, ed = self._map(lambda t: t.to(device), lambda t: t.to(device))
        return DGLGraph(e, self._num_nodes, nd, ed, self.idtype)

    def int(self):
        e, nd, ed = self._map(lambda t: t.to(th.int32), lambda t: t)
        return DGLGraph(e, self._num_nodes, nd, ed, th.int32)

    def long(self):
        e, nd, ed = self._map(lambda t: t.to(th.int64), lambda t: t)
        return DGLGraph(e, self._num_nodes, nd, ed, th.int64)

    def cpu(self):
        return self.to(th.device('cpu'))

    def __repr__(self):
        return 'Graph(num_nodes=%r, num_edges=%r)' % (
            self._num_nodes, {c: int(s.numel()) for c, (s, _) in self._edges.items()})


DGLHeteroGraph = DGLGraph


def heterograph(data_dict, num_nodes_dict=None, idtype=None, device=None):
    edges = {}
    want = idtype
    for c, (u, v) in data_dict.items():
        if want is None and isinstance(u, th.Tensor) and u.dtype in (th.int32, th.int64):
            want = u.dtype
    want = want or th.int64
    for c, (u, v) in data_dict.items():
        edges[tuple(c)] = (_as_index(u, want, device), _as_index(v, want, device))
    if num_nodes_dict is None:
        num_nodes_dict = {}
        for (st, _, dt), (u, v) in edges.items():
            num_nodes_dict[st] = max(num_nodes_dict.get(st, 0), int(u.max()) + 1 if u.numel() else 0)
            num_nodes_dict[dt] = max(num_nodes_dict.get(dt, 0), int(v.max()) + 1 if v.numel() else 0)
    return DGLGraph(edges, num_nodes_dict, idtype=want)


def bipartite_from_scipy(sp_mat, utype, etype, vtype, eweight_name=None, idtype=None, device=None):
    coo = sp_mat.tocoo()                                   # a coo_matrix returns itself: order kept
    return heterograph({(utype, etype, vtype): (coo.row, coo.col)},
                       num_nodes_dict={utype: coo.shape[0], vtype: coo.shape[1]},
                       idtype=idtype, device=device)


from . import nn  # noqa: E402,F401
