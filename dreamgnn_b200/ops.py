"""Tensor-level wrappers over the C ABI (include/dreamgnn.h) and the autograd Functions built on
them. Everything here requires CUDA tensors; there is no CPU fallback.

Reference call sites replaced (citations into /root/reference):
  * `CSR.from_coo` / `CSR.transpose`   DGL's lazy COO->CSC/CSR inside update_all (layers.py:229-232)
  * `spmm` / `SpMMFunction`            update_all(copy_u,sum) with cj/ci fused (layers.py:224-234) and
                                       th.spmm(adj, support)+bias (layers.py:312-314); backward = same
                                       kernel on the transposed CSR, atomic-free
  * `decoder_mlp` / `DecoderFunction`  apply_edges(udf_u_mul_e) + lin1/lin2/lin3 (layers.py:360-379)
  * `csr_dropout`                      th.randperm edge dropout + graph rebuild (augmentation.py:35-124)
  * `topk_rows`, `knn_graph_from_neighbors`   np.argpartition + scipy A+A^T+I, D^-1 A (data_loader.py:291-308)
"""
import os

import torch as th

from . import _lib as L

I32 = th.int32
DEC_H1, DEC_H2 = 128, 64
SPMM_ACCUMULATE, SPMM_RELU, SPMM_PREFETCH, SPMM_ROWSPLIT = 1, 2, 4, 8
# few, long rows: CTA per row instead of warp per row (148 SMs x 16 resident warps want >= ~2400 rows otherwise)
SPMM_ROWSPLIT_MAX_ROWS, SPMM_ROWSPLIT_MIN_AVG_LEN = 2368, 128
# Gathered operands that do not stay resident in L2 (126 MB, shared with the index and output streams): the SpMM then
# prefetches the rows it is about to gather into L2 one group ahead (DG_SPMM_PREFETCH flag of the C ABI). Measured at
# syn20m: d=344 launches -22 %, d=768 -16 %, the decoder's segment sums -10 %; L2-resident launches of narrow rows
# (d=128, 94 % hit rate) lose ~20 % to the extra instructions -- hence the switch on operand size and row width.
SPMM_PREFETCH_MIN_BYTES = int(os.environ.get('DG_SPMM_PREFETCH_MIN_MB', '120')) << 20          # any row width
SPMM_PREFETCH_MIN_BYTES_WIDE = int(os.environ.get('DG_SPMM_PREFETCH_MIN_MB_WIDE', '48')) << 20   # rows of >= 1 KiB
# Skewed graphs (a few rows of 10^4...10^5 edges among rows of 10^1...10^2: Zipf popularity, SURVEY.md 8d "secondary stress
# set"): warp-per-row leaves the launch waiting on the one warp that walks the longest row. Measured on the Zipf(1.0)
# variant of syn20m (13.9 M pairs, two 50 000-edge rows, median 34; d=344): 9.17 ms in one launch against cuSPARSE's 2.42 ms.
# A CSR whose longest row exceeds SPMM_SPLIT_MIN_DEG at build time is therefore aggregated in two launches of the same
# kernel: the rows cut into chunks of <= SPMM_SPLIT_T edges (a second indptr over the SAME indices / values), then the
# chunk partials summed per row in chunk order with the epilogue (ci, bias, ReLU) applied there -- deterministic, no
# atomics. Same graph: 1.49 ms at T=128 (1.55 / 1.70 / 1.92 / 2.23 ms at 256 / 512 / 1024 / 2048), i.e. 1.6x cuSPARSE
# (profiles/r02c_library_compare.md). Uniform shapes (every BASELINE config: longest row <= ~1 000) never take this path.
SPMM_SPLIT = os.environ.get('DG_SPMM_SPLIT', '1') != '0'
SPMM_SPLIT_T = int(os.environ.get('DG_SPMM_SPLIT_T', '128'))
SPMM_SPLIT_MIN_DEG = int(os.environ.get('DG_SPMM_SPLIT_MIN_DEG', '4096'))


def _i32(t, name):
    """Narrow an index tensor to int32 with a range check (int64 accepted at the Python boundary)."""
    if t.dtype == I32:
        return t.contiguous()
    if t.dtype != th.int64:
        raise TypeError('%s must be int32 or int64' % name)
    if t.numel() and (int(t.max()) > 0x7fffffff or int(t.min()) < 0):
        raise ValueError('%s out of int32 range' % name)
    return t.to(I32).contiguous()


# ------------------------------------------------------------------------------------------------
# index primitives
# ------------------------------------------------------------------------------------------------
def exclusive_scan_i32(x):
    lib = L.load()
    n = x.numel()
    out = th.empty(n + 1, dtype=I32, device=x.device)
    ws = L.workspace(lib.dg_scan_workspace_bytes(n), x.device)
    L.check(lib.dg_exclusive_scan_i32(L.ptr(x, I32, 'x'), L.ptr(out), n, L.ptr(ws), ws.numel(), L.stream()), 'scan')
    return out


def sort_pairs_u64(keys, vals, key_bits):
    """Stable sort of int64-stored unsigned keys (+ optional int32 payload). Inputs are clobbered."""
    lib = L.load()
    n = keys.numel()
    keys_out = th.empty_like(keys)
    vals_out = th.empty_like(vals) if vals is not None else None
    ws = L.workspace(lib.dg_sort_workspace_bytes(n), keys.device)
    L.check(lib.dg_sort_pairs_u64(L.ptr(keys, th.int64, 'keys'), L.ptr(vals, I32 if vals is not None else None),
                                  L.ptr(keys_out), L.ptr(vals_out), n, int(key_bits), L.ptr(ws), ws.numel(),
                                  L.stream()), 'sort_pairs')
    return keys_out, vals_out


# ------------------------------------------------------------------------------------------------
# CSR container
# ------------------------------------------------------------------------------------------------
class CSR:
    """Canonical CSR (rows ascending, columns ascending inside a row) on device, int32.

    `eid[s]` is the id of the edge stored in slot s -- ids refer to the COO list the *base* graph was
    built from, and survive transposition and dropout compaction, so forward / transposed / dropped
    variants of one graph stay consistent.
    """

    __slots__ = ('indptr', 'indices', 'eid', 'vals', 'n_rows', 'n_cols', '_t', 'slot_order', 'eid_is_slot', 'parent_nnz',
                 'split_T', '_plan')

    def __init__(self, indptr, indices, eid, vals, n_rows, n_cols):
        self.indptr, self.indices, self.eid, self.vals = indptr, indices, eid, vals
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self._t = None
        self.slot_order = False     # True when the owning COO tensor lists its entries in slot order
        self.eid_is_slot = False    # True when eid[s] == s (ids already name this tensor's own entries)
        self.parent_nnz = 0         # edge-dropped CSR: entry count of the graph its eids refer to (sizes id lookups)
        self.split_T = 0            # > 0: rows are aggregated in chunks of this many edges (skewed graphs; see split_plan)
        self._plan = None           # cached split_plan of this CSR

    @property
    def nnz(self):
        return int(self.indices.numel())

    @property
    def device(self):
        return self.indptr.device

    @staticmethod
    def from_coo(row, col, n_rows, n_cols, vals=None, edge_ids=None):
        lib = L.load()
        row, col = _i32(row, 'row'), _i32(col, 'col')
        if row.numel() and (int(row.max()) >= n_rows or int(col.max()) >= n_cols):
            raise ValueError('edge endpoint out of range')
        n = row.numel()
        dev = row.device
        indptr = th.empty(n_rows + 1, dtype=I32, device=dev)
        indices = th.empty(n, dtype=I32, device=dev)
        eid = th.empty(n, dtype=I32, device=dev)
        ws = L.workspace(lib.dg_csr_build_workspace_bytes(n, n_rows), dev)
        L.check(lib.dg_csr_build(L.ptr(row, I32, 'row'), L.ptr(col, I32, 'col'), n, n_rows, n_cols, L.ptr(indptr),
                                 L.ptr(indices), L.ptr(eid), L.ptr(ws), ws.numel(), L.stream()), 'csr_build')
        v = None
        if vals is not None:
            v = vals.contiguous()[eid.long()]
        if edge_ids is not None:
            eid = _i32(edge_ids, 'edge_ids')[eid.long()]
        csr = CSR(indptr, indices, eid, v, n_rows, n_cols)
        if SPMM_SPLIT and n > SPMM_SPLIT_MIN_DEG and n_rows > 0:
            # one host read per graph BUILD (from_coo already reads the index range back); derived structures -- dropout
            # compactions, staged clones, re-valued copies -- inherit the decision and never synchronise
            if int((indptr[1:] - indptr[:-1]).max()) > SPMM_SPLIT_MIN_DEG:
                csr.split_T = SPMM_SPLIT_T
        return csr

    def rows(self):
        """COO row id of every slot."""
        lib = L.load()
        out = th.empty(self.nnz, dtype=I32, device=self.device)
        L.check(lib.dg_csr_expand_rows(L.ptr(self.indptr), self.n_rows, L.ptr(out), L.stream()), 'expand_rows')
        return out

    def transpose(self):
        """Cached canonical CSR of the transposed matrix, carrying edge ids and values along."""
        if self._t is None:
            t = CSR.from_coo(self.indices, self.rows(), self.n_cols, self.n_rows, self.vals, self.eid)
            t._t = self
            self._t = t
        return self._t

    def degrees(self):
        return (self.indptr[1:] - self.indptr[:-1])

    def degree_norm(self):
        """1/sqrt(degree) with 0 -> 0 (data_loader.py:454-457), fp32 [n_rows]."""
        lib = L.load()
        out = th.empty(self.n_rows, dtype=th.float32, device=self.device)
        L.check(lib.dg_degree_norm(L.ptr(self.indptr), self.n_rows, L.ptr(out), L.stream()), 'degree_norm')
        return out


def inherit_layout(new, base):
    """`new` is a structure derived from `base` with the same row population (compaction, clone, re-valued copy):
    it keeps base's chunked-aggregation decision. Returns `new`."""
    new.split_T = base.split_T
    return new


def split_plan(indptr, n_rows, nnz, T):
    """Chunked view of a CSR for rows far longer than the rest (pure index arithmetic on the device, static shapes, no
    host read: capturable, and runs on CPU tensors for the tests).

    Row r of `deg` edges becomes max(1, ceil(deg / T)) consecutive virtual rows of <= T edges each. Returns
      v_indptr    int32 [n_v + 1]   indptr of the virtual rows over the SAME indices / values arrays,
      comb_indptr int32 [n_rows + 1] row r's virtual rows are comb_indptr[r] .. comb_indptr[r + 1] - 1, in edge order,
      n_v         = n_rows + nnz // T, a static upper bound on the number of virtual rows (the unused tail is empty rows).
    """
    n_v = int(n_rows) + int(nnz) // int(T)
    ip = indptr.long()
    deg = ip[1:] - ip[:-1]
    nch = ((deg + (T - 1)) // T).clamp_(min=1)
    vend = th.cumsum(nch, 0)                                   # inclusive: first virtual row AFTER row r
    vbeg = vend - nch
    v = th.arange(n_v + 1, device=indptr.device)
    r = th.searchsorted(vend, v, right=True).clamp_(max=int(n_rows) - 1)
    start = th.minimum(ip[:-1][r] + (v - vbeg[r]) * T, ip[1:][r])      # past the last real chunk: == nnz (empty rows)
    comb = th.zeros(int(n_rows) + 1, dtype=I32, device=indptr.device)
    comb[1:] = vend
    return start.to(I32), comb, n_v


class RandomSubset:
    """Stands in for a permutation in `keep_flags`: "a uniformly random subset of `num_keep` of these `n` edges", drawn
    by the sort-free radix select of `dg_random_subset_flags` when the flags are built."""

    def __init__(self, n):
        self.n = int(n)


def random_subset_flags(n, num_keep, out, rnd=None):
    """out[i] = 1 for a uniformly random `num_keep`-subset of range(n) (uint8 [n], contiguous); graph-capturable.
    `rnd` (int64 [n], 63 random bits each) defaults to a fresh draw from torch's generator."""
    lib = L.load()
    if rnd is None:
        rnd = th.empty(n, dtype=th.int64, device=out.device).random_()
    ws = L.workspace(lib.dg_random_subset_workspace_bytes(), out.device)
    L.check(lib.dg_random_subset_flags(L.ptr(rnd, th.int64, 'rnd'), int(n), int(num_keep), L.ptr(out, th.uint8, 'flags'),
                                       L.ptr(ws), ws.numel(), L.stream()), 'random_subset_flags')
    return out


def keep_flags(n_edges, perms, device):
    """uint8 keep flag per edge id: perms = [(perm int64 | RandomSubset, num_keep, id offset), ...]
    (augmentation.py:48-52: keep the first num_keep entries of each randperm)."""
    lib = L.load()
    flags = th.zeros(n_edges, dtype=th.uint8, device=device)
    for perm, num_keep, offset in perms:
        if isinstance(perm, RandomSubset):
            random_subset_flags(perm.n, num_keep, flags[int(offset):int(offset) + perm.n])
            continue
        L.check(lib.dg_keep_flags_from_perm(L.ptr(perm, th.int64, 'perm'), int(num_keep), int(offset), L.ptr(flags),
                                            L.stream()), 'keep_flags')
    return flags


def csr_compact(csr, flags, n_keep):
    """Order-preserving compaction of `csr` to the edges whose flag (by edge id) is set."""
    lib = L.load()
    dev = csr.device
    indptr = th.empty(csr.n_rows + 1, dtype=I32, device=dev)
    indices = th.empty(n_keep, dtype=I32, device=dev)
    eid = th.empty(n_keep, dtype=I32, device=dev)
    vals = th.empty(n_keep, dtype=th.float32, device=dev) if csr.vals is not None else None
    ws = L.workspace(lib.dg_csr_compact_workspace_bytes(csr.n_rows), dev)
    L.check(lib.dg_csr_compact(L.ptr(csr.indptr), L.ptr(csr.indices), L.ptr(csr.eid), L.ptr(csr.vals), csr.n_rows,
                               L.ptr(flags, th.uint8, 'flags'), L.ptr(indptr), L.ptr(indices), L.ptr(eid),
                               L.ptr(vals), L.ptr(ws), ws.numel(), L.stream()), 'csr_compact')
    return inherit_layout(CSR(indptr, indices, eid, vals, csr.n_rows, csr.n_cols), csr)


def csr_dropout(csr, flags, n_keep):
    """Dropped graph in both orientations from one flag array (forward + transposed stay consistent)."""
    fwd = csr_compact(csr, flags, n_keep)
    bwd = csr_compact(csr.transpose(), flags, n_keep)
    fwd._t, bwd._t = bwd, fwd
    return fwd


# ------------------------------------------------------------------------------------------------
# independent sub-computations as parallel stream branches
# ------------------------------------------------------------------------------------------------
# Off by default. `graphed.GraphedIteration` switches it on while it captures the launch-bound real-dataset shapes:
# the branches then become parallel paths of the CUDA graph (forward, and -- autograd replays every backward node on
# the stream of its forward -- backward too). Same kernels, same order inside each branch: results are bit-identical.
PARALLEL_BRANCHES = False


class parallel_branches:
    """Context manager: `with parallel_branches(True): ...` enables `branches()` forking inside the block."""

    def __init__(self, enable=True):
        self.enable = enable

    def __enter__(self):
        global PARALLEL_BRANCHES
        self.saved = PARALLEL_BRANCHES
        if self.enable:
            PARALLEL_BRANCHES = True

    def __exit__(self, *exc):
        global PARALLEL_BRANCHES
        PARALLEL_BRANCHES = self.saved
        return False


def _each_tensor(x):
    if isinstance(x, th.Tensor):
        yield x
    elif isinstance(x, (tuple, list)):
        for y in x:
            yield from _each_tensor(y)
    elif isinstance(x, dict):
        for y in x.values():
            yield from _each_tensor(y)


# Side streams of the parallel branches. `th.cuda.Stream()` hands out the 32 streams of torch's per-device pool round-robin,
# so a process that records several graphs (or just calls `branches` often) sooner or later gets, for a "parallel" branch,
# the very stream its sibling or the main chain runs on -- the branches then serialise silently (measured: the Cdataset shape
# 1.24 ms / iteration captured first in a process, 1.45 ms captured after two other shapes). The branches therefore draw
# from a fixed set of streams with pairwise distinct handles, in the same order in every iteration.
_SIDE_STREAMS = {}
_SIDE_NEXT = {}
_N_SIDE = 12                       # per iteration: routes 1 + node-type halves 4 + common losses 1 < 12; the 13th is the augmentation's


def _distinct_streams(device):
    lst = _SIDE_STREAMS.get(device)
    if lst is None:
        lst = _SIDE_STREAMS[device] = []
        for _ in range(64):
            s = th.cuda.Stream(device=device)
            if all(s.cuda_stream != t.cuda_stream for t in lst):
                lst.append(s)
            if len(lst) == _N_SIDE + 1:
                break
    return lst


def side_stream(main):
    """The next side stream of this iteration's branches: never the stream `main` itself."""
    lst = _distinct_streams(main.device)[:_N_SIDE]
    i = _SIDE_NEXT.get(main.device, 0)
    for _ in range(len(lst)):
        s = lst[i % len(lst)]
        i += 1
        if s.cuda_stream != main.cuda_stream:
            break
    _SIDE_NEXT[main.device] = i
    return s


def augmentation_stream(main):
    """The stream of the pipelined augmentation branch (held for a whole iteration: its own slot)."""
    s = _distinct_streams(main.device)[-1]
    return s if s.cuda_stream != main.cuda_stream else th.cuda.Stream(device=main.device)


def reset_side_streams(device):
    """Start of an iteration: its branches take the side streams in the same order as every other iteration's."""
    _SIDE_NEXT[th.device(device)] = 0


def branches(fns):
    """Results of the independent closures `fns`, in order. With PARALLEL_BRANCHES the first runs on the current
    stream and each other one on its own side stream forked from / joined to it."""
    if not PARALLEL_BRANCHES or len(fns) < 2 or not th.cuda.is_available():
        return [f() for f in fns]
    main = th.cuda.current_stream()
    sides = [side_stream(main) for _ in fns[1:]]
    out = [None] * len(fns)
    for i, st in enumerate(sides, start=1):
        st.wait_stream(main)
        with th.cuda.stream(st):
            out[i] = fns[i]()
    out[0] = fns[0]()
    for i, st in enumerate(sides, start=1):
        main.wait_stream(st)
        for t in _each_tensor(out[i]):
            t.record_stream(main)
    return out


# ------------------------------------------------------------------------------------------------
# row-streaming helpers (rowops.cu)
# ------------------------------------------------------------------------------------------------
def multi_copy(dsts, srcs):
    """dst.copy_(src) for many same-shaped contiguous CUDA tensor pairs in ONE launch (per 96 pairs)."""
    lib = L.load()
    pairs = [(d, s_) for d, s_ in zip(dsts, srcs) if d.numel()]
    if not pairs:
        return
    arr = (L.CopyItem * len(pairs))()
    for i, (d, s_) in enumerate(pairs):
        if not (d.is_cuda and s_.is_cuda and d.is_contiguous() and s_.is_contiguous() and d.dtype == s_.dtype and d.shape == s_.shape):
            raise ValueError('multi_copy: contiguous CUDA tensors of equal dtype and shape expected')
        arr[i].dst, arr[i].src, arr[i].bytes = d.data_ptr(), s_.data_ptr(), d.numel() * d.element_size()
    L.check(lib.dg_multi_copy(arr, len(pairs), L.stream()), 'multi_copy')


# Zero-initialised int32 tickets lent to kernels that finish in their last CTA (dg_colsum_f32): handed out round-robin from
# one persistent pool per device, so two calls that may be in flight at the same time (parallel stream branches, or nodes
# of one captured graph) never share a slot; every kernel leaves its tickets zero again.
_TICKET_POOL = {}
_TICKET_SLOTS = 8192


def _tickets(device, count):
    """Device address of `count` zeroed tickets, or None (the caller then takes its two-launch form) when the request is
    larger than the pool serves or when the pool would have to be created inside a CUDA-graph capture -- it must outlive
    every graph, so it is only ever allocated eagerly (every capture in this package follows an eager warm-up)."""
    if count > _TICKET_SLOTS // 4:
        return None
    pool = _TICKET_POOL.get(device)
    if pool is None:
        if th.cuda.is_current_stream_capturing():
            return None
        pool = _TICKET_POOL[device] = [th.zeros(_TICKET_SLOTS, dtype=th.int32, device=device), 0]
    if pool[1] + count > _TICKET_SLOTS:
        pool[1] = 0
    addr = pool[0].data_ptr() + 4 * pool[1]
    pool[1] += count
    return addr


def _rows_ok(t):
    return (t.dim() == 2 and t.dtype == th.float32 and t.stride(1) == 1 and t.shape[1] % 4 == 0
            and t.stride(0) % 4 == 0 and t.stride(0) >= t.shape[1] and t.data_ptr() % 16 == 0)


def colsum(x, gate=None, want_masked=False):
    """Deterministic x.sum(0) of a [n, d] fp32 matrix in one pass; with `gate` the sum (and, if `want_masked`,
    the returned matrix) is of x * (gate > 0) -- F.relu's backward fused with the bias gradient.
    Returns sum [d] or (masked [n, d], sum [d]). Shapes the kernel cannot address directly (d % 4 != 0, unaligned rows)
    go through zero-padded aligned copies."""
    if not x.is_cuda:
        raise RuntimeError('dreamgnn_b200.colsum needs CUDA tensors (no CPU fallback)')
    if x.dim() != 2 or x.dtype != th.float32 or (gate is not None and gate.shape != x.shape):
        raise ValueError('colsum: a 2-D fp32 matrix (and a gate of the same shape) expected')
    if x.shape[0] == 0:
        return (x, x.new_zeros(x.shape[1])) if want_masked else x.new_zeros(x.shape[1])
    if not (_rows_ok(x) and (gate is None or _rows_ok(gate))):
        # rows the float4 kernel cannot address (width not a multiple of 4, unaligned or strided rows): the same kernel
        # on zero-padded aligned copies -- never a library reduction
        d0 = x.shape[1]
        aligned = lambda t: th.nn.functional.pad(t, (0, (-d0) % 4)).contiguous()
        res = colsum(aligned(x), None if gate is None else aligned(gate), want_masked)
        return (res[0][:, :d0], res[1][:d0]) if want_masked else res[:d0]
    lib = L.load()
    n, d = x.shape
    out = th.empty(d, dtype=th.float32, device=x.device)
    y = th.empty((n, d), dtype=th.float32, device=x.device) if (want_masked and gate is not None) else None
    ws = L.workspace(lib.dg_colsum_workspace_bytes(n, d), x.device)
    L.check(lib.dg_colsum_f32(x.data_ptr(), x.stride(0), None if gate is None else gate.data_ptr(),
                              0 if gate is None else gate.stride(0), L.ptr(y), d, n, d, L.ptr(out), L.ptr(ws), ws.numel(),
                              _tickets(x.device, (d + 127) // 128), L.stream()), 'colsum')
    if want_masked:
        return (x if y is None else y), out
    return out


class GramCommonLoss(th.autograd.Function):
    """common_loss (utils.py:87-95) in its Gram-matrix form, forward and backward with explicit kernels:

        Z = [z1 | z2]  (rows centred by the column mean, L2-normalised, float64)      dg_colsum + dg_center_normalize
        G = Z^T Z = [[G11, G12], [G21, G22]]                                          one float64 GEMM (library)
        loss = (|G11|^2 + |G22|^2 - 2 |G12|^2) / n^2 = sum(G * G * S) / n^2,  S = [[+1, -1], [-1, +1]]
        dL/dZ = (4 / n^2) Z (G * S);   dc = (gz - z <z, gz>) / |c|;   dx = dc - colmean(dc)

    ~18 launches per call against ~95 for the autograd-traced torch expression; same value (float64 accumulation)."""

    @staticmethod
    def forward(ctx, emb1, emb2):
        lib = L.load()
        n, d = emb1.shape
        dev = emb1.device
        z = th.empty((n, 2 * d), dtype=th.float64, device=dev)
        inv = th.empty((2, n), dtype=th.float64, device=dev)
        for i, e in enumerate((emb1, emb2)):
            cs = colsum(e)
            L.check(lib.dg_center_normalize_f64(e.data_ptr(), e.stride(0), L.ptr(cs), n, d, 1e-12,
                                                z.data_ptr() + i * d * 8, 2 * d, inv.data_ptr() + i * n * 8, L.stream()),
                    'center_normalize')
        g = z.t() @ z                                                     # [2d, 2d] float64
        gs = th.empty((2 * d, 2 * d), dtype=th.float64, device=dev)       # G * S: the off-diagonal blocks negated
        loss = th.empty((), dtype=th.float32, device=dev)
        L.check(lib.dg_gram_common_loss_f64(g.data_ptr(), g.stride(0), d, float(n), L.ptr(gs), L.ptr(loss), L.stream()),
                'gram_common_loss')                                       # loss = sum(G * G * S) / n^2, one launch
        ctx.save_for_backward(z, inv, gs)
        ctx.shape = (n, d)
        return loss

    @staticmethod
    def backward(ctx, gout):
        lib = L.load()
        z, inv, gs = ctx.saved_tensors
        n, d = ctx.shape
        gz = z @ gs                                                       # [n, 2d] float64; dL/dZ = gz * gout * 4 / n^2
        gout = gout.reshape(()).to(th.float32)
        grads = []
        for i in range(2):
            if not ctx.needs_input_grad[i]:
                grads.append(None)
                continue
            dc = th.empty((n, d), dtype=th.float32, device=z.device)
            L.check(lib.dg_center_normalize_bwd_f64(gz.data_ptr() + i * d * 8, 2 * d, z.data_ptr() + i * d * 8, 2 * d,
                                                    inv.data_ptr() + i * n * 8, n, d, L.ptr(dc), d, L.ptr(gout),
                                                    4.0 / float(n) / float(n), L.stream()),
                    'center_normalize_bwd')
            grads.append(dc.sub_(colsum(dc) / float(n)))
        return tuple(grads)


def gram_common_loss(emb1, emb2):
    """Fused common_loss in Gram form for [n, d] fp32 CUDA embeddings (d % 4 == 0)."""
    if not (emb1.is_cuda and emb2.is_cuda):
        raise RuntimeError('dreamgnn_b200.gram_common_loss needs CUDA tensors (no CPU fallback)')
    if emb1.shape != emb2.shape or not (_rows_ok(emb1) and _rows_ok(emb2)):
        raise ValueError('gram_common_loss: two [n, d] fp32 matrices with 16-byte aligned rows expected')
    return GramCommonLoss.apply(emb1, emb2)


# ------------------------------------------------------------------------------------------------
# basis decomposition of the GCMC relation weights (csrc/basis.cu)
# ------------------------------------------------------------------------------------------------
class BasisCombineFunction(th.autograd.Function):
    """W [R, in, Dp] = sum_b att[r, b] * basis[b], message width zero-padded D -> Dp in the same pass (layers.py:120-121)."""

    @staticmethod
    def forward(ctx, att, basis, d_pad):
        lib = L.load()
        att, basis = att.contiguous(), basis.contiguous()
        (R, B), (_, rows, D) = att.shape, basis.shape
        w = th.empty((R, rows, d_pad), dtype=th.float32, device=att.device)
        L.check(lib.dg_basis_combine_fwd_f32(L.ptr(att, th.float32, 'att'), L.ptr(basis, th.float32, 'basis'), R, B, rows, D, d_pad,
                                             L.ptr(w), L.stream()), 'basis_combine_fwd')
        ctx.save_for_backward(att, basis)
        return w

    @staticmethod
    def backward(ctx, dw):
        lib = L.load()
        att, basis = ctx.saved_tensors
        (R, B), (_, rows, D) = att.shape, basis.shape
        dw = dw.contiguous()
        dbasis = th.empty_like(basis)
        datt = th.empty_like(att)
        ws = L.workspace(lib.dg_basis_combine_bwd_workspace_bytes(rows, D), att.device)
        L.check(lib.dg_basis_combine_bwd_f32(L.ptr(att), L.ptr(basis), L.ptr(dw, th.float32, 'dw'), R, B, rows, D, dw.shape[2],
                                             L.ptr(dbasis), L.ptr(datt), L.ptr(ws), ws.numel(), L.stream()), 'basis_combine_bwd')
        return datt, dbasis, None


BASIS_MAX = 4


def basis_combine(att, basis, mult=1):
    """matmul(att [R, B], basis.view(B, -1)).view(R, in, D) zero-padded to a width that is a multiple of `mult`
    ([R, in, Dp]); R, B <= 4 run as one elementwise kernel each way, anything larger as the reference's matmul + pad."""
    if not att.is_cuda:
        raise RuntimeError('dreamgnn_b200.basis_combine needs CUDA tensors (no CPU fallback)')
    (R, B), (_, rows, D) = att.shape, basis.shape
    d_pad = D + (-D) % max(int(mult), 1)
    if R <= BASIS_MAX and B <= BASIS_MAX and att.dtype == th.float32 and basis.dtype == th.float32:
        return BasisCombineFunction.apply(att, basis, d_pad)
    w = th.matmul(att, basis.reshape(B, -1)).view(R, rows, D)
    return th.nn.functional.pad(w, (0, d_pad - D)) if d_pad != D else w


class WeightedLossSum(th.autograd.Function):
    """total = rel + beta * (c1 + c2) (train.py:292-294) for three scalars: two launches forward, and the backward hands
    the upstream gradient through and scales it once (autograd's trace of the expression: 3 + 5 scalar launches)."""

    @staticmethod
    def forward(ctx, rel, c1, c2, beta):
        ctx.beta = float(beta)
        return th.add(rel, c1 + c2, alpha=float(beta))

    @staticmethod
    def backward(ctx, g):
        gb = g * ctx.beta
        return g, gb, gb, None


# ------------------------------------------------------------------------------------------------
# binary cross entropy over the scored pairs (csrc/loss.cu)
# ------------------------------------------------------------------------------------------------
class BCEWithLogitsFunction(th.autograd.Function):
    """mean BCE-with-logits (train.py:291; targets smoothed as train.py:15-23): two launches forward, one backward."""

    @staticmethod
    def forward(ctx, logits, target, smoothing):
        lib = L.load()
        n = logits.numel()
        loss = th.empty((), dtype=th.float32, device=logits.device)
        ws = L.workspace(lib.dg_bce_logits_workspace_bytes(n), logits.device)
        L.check(lib.dg_bce_logits_fwd_f32(L.ptr(logits, th.float32, 'logits'), L.ptr(target, th.float32, 'target'), n,
                                          float(smoothing), L.ptr(loss), L.ptr(ws), ws.numel(), L.stream()), 'bce_logits_fwd')
        ctx.save_for_backward(logits, target)
        ctx.smoothing = float(smoothing)
        return loss

    @staticmethod
    def backward(ctx, gout):
        lib = L.load()
        logits, target = ctx.saved_tensors
        dx = th.empty_like(logits)
        gout = gout.to(th.float32).contiguous()
        L.check(lib.dg_bce_logits_bwd_f32(L.ptr(logits), L.ptr(target), logits.numel(), ctx.smoothing, L.ptr(gout), L.ptr(dx),
                                          L.stream()), 'bce_logits_bwd')
        return dx, None, None


def bce_with_logits(logits, target, smoothing=0.0):
    """mean(binary_cross_entropy_with_logits(logits, target * (1 - s) + s / 2)) for fp32 CUDA tensors of equal shape."""
    if not (isinstance(logits, th.Tensor) and logits.is_cuda and target.is_cuda):
        raise RuntimeError('dreamgnn_b200.bce_with_logits needs CUDA tensors (no CPU fallback)')
    if logits.shape != target.shape or logits.numel() == 0:
        raise ValueError('bce_with_logits: logits %s and target %s must have the same non-empty shape'
                         % (tuple(logits.shape), tuple(target.shape)))
    return BCEWithLogitsFunction.apply(logits.to(th.float32).contiguous(), target.to(th.float32).contiguous(), float(smoothing))


class FusedBCEWithLogitsLoss(th.nn.Module):
    """Drop-in for nn.BCEWithLogitsLoss() (mean reduction) and train.py's LabelSmoothingBCELoss(smoothing)."""

    def __init__(self, smoothing=0.0):
        super().__init__()
        self.smoothing = float(smoothing)

    def forward(self, pred, target):
        return bce_with_logits(pred, target, self.smoothing)


# ------------------------------------------------------------------------------------------------
# fused row kernels (csrc/fused.cu)
# ------------------------------------------------------------------------------------------------
ACT_CODES = {None: 0, 'identity': 0, 'leaky': 1, 'relu': 2}


class _SeedPool:
    """Seeds of one training iteration drawn by ONE launch (`begin_seed_pool`, called at the top of the iteration on its
    main stream) instead of one `randint` launch in front of each of the ~15 fused dropout kernels."""
    SIZE = 32

    def __init__(self, device):
        self.buf = th.randint(0, 2 ** 62, (self.SIZE,), device=device, dtype=th.int64)
        self.next = 0
        self.capturing = th.cuda.is_current_stream_capturing()
        self.stream = th.cuda.current_stream(device)
        self.event = th.cuda.Event()
        self.event.record(self.stream)


_SEED_POOLS = {}


def begin_seed_pool(device):
    """Start of a training iteration: draw its dropout seeds (one launch) and rewind the side streams of its parallel
    branches. Call on the stream the iteration's branches fork from."""
    device = th.device(device)
    if device.type == 'cuda':
        _SEED_POOLS[device] = _SeedPool(device)
        reset_side_streams(device)


def drop_seed_pool(device=None):
    """Forget the pool (after a CUDA-graph capture its buffer belongs to the graph's memory and is rewritten only by replays)."""
    if device is None:
        _SEED_POOLS.clear()
    else:
        _SEED_POOLS.pop(th.device(device), None)


def fresh_seed(device):
    """A dropout seed drawn on the device (no host sync; a captured CUDA graph gets a fresh one on every replay): the next
    slot of the iteration's pool when there is one (same capture state, not used up), else its own `randint` launch."""
    device = th.device(device)
    pool = _SEED_POOLS.get(device)
    if pool is not None and pool.next < pool.SIZE and pool.capturing == th.cuda.is_current_stream_capturing():
        cur = th.cuda.current_stream(device)
        if cur != pool.stream:
            cur.wait_event(pool.event)                 # a parallel branch forked before / beside the draw
        seed = pool.buf[pool.next:pool.next + 1]
        pool.next += 1
        return seed
    return th.randint(0, 2 ** 62, (1,), device=device, dtype=th.int64)


def _seed_args(seed):
    dev = seed if isinstance(seed, th.Tensor) else None
    return (0 if dev is not None else int(seed)), dev


class ActDropoutFunction(th.autograd.Function):
    """y = dropout(act(x)) in one launch; the backward regenerates the keep mask from (seed, row, column group)."""

    @staticmethod
    def forward(ctx, x, act, slope, p, seed):
        lib = L.load()
        out = th.empty((x.shape[0], x.shape[1]), dtype=th.float32, device=x.device)
        sv, sd = _seed_args(seed)
        L.check(lib.dg_act_dropout_f32(x.data_ptr(), x.stride(0), None, 0, L.ptr(out), out.stride(0), x.shape[0], x.shape[1],
                                       act, float(slope), float(p), sv, L.ptr(sd), L.stream()), 'act_dropout')
        ctx.save_for_backward(x, sd)
        ctx.cfg = (act, float(slope), float(p), sv)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, sd = ctx.saved_tensors
        act, slope, p, sv = ctx.cfg
        dy = dy if _rows_ok(dy) else dy.contiguous()
        dx = th.empty((x.shape[0], x.shape[1]), dtype=th.float32, device=x.device)
        L.check(lib.dg_act_dropout_f32(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), L.ptr(dx), dx.stride(0), x.shape[0],
                                       x.shape[1], act, slope, p, sv, L.ptr(sd), L.stream()), 'act_dropout_bwd')
        return dx, None, None, None, None


def act_dropout(x, act=None, slope=0.1, p=0.0, training=True, seed=None):
    """dropout(act(x)) for a 2-D fp32 CUDA matrix, act in {None, 'leaky', 'relu'} (layers.py:134-138, 247, 281-282).
    Rows that the float4 kernel cannot address (width not a multiple of 4, unaligned) go through a zero-padded copy."""
    if not x.is_cuda:
        raise RuntimeError('dreamgnn_b200.act_dropout needs CUDA tensors (no CPU fallback)')
    p = float(p) if training else 0.0
    if act in (None, 'identity') and p == 0.0:
        return x
    if x.dim() != 2 or x.dtype != th.float32:
        raise ValueError('act_dropout: a 2-D fp32 matrix expected')
    if x.shape[0] == 0:
        return x
    if not _rows_ok(x):
        # rows the float4 kernel cannot address: the same kernel on a zero-padded aligned copy (act(0) = 0 for all three)
        d0 = x.shape[1]
        xp = th.nn.functional.pad(x, (0, (-d0) % 4)).contiguous()
        return act_dropout(xp, act, slope, p, True, seed)[:, :d0]
    if seed is None:
        seed = fresh_seed(x.device) if p > 0 else 0
    return ActDropoutFunction.apply(x, ACT_CODES[act], slope, p, seed)


ATT_MAX_HIDDEN, ATT_MAX_WIDTH = 16, 256      # the backward's shared-memory reduction holds 9 x [16, d] floats


class AttentionFunction(th.autograd.Function):
    """Attention.forward (layers.py:324-338) over two views in one kernel each way."""

    @staticmethod
    def forward(ctx, za, zb, w1, b1, w2, p, seed):
        lib = L.load()
        n, d = za.shape
        out = th.empty((n, d), dtype=th.float32, device=za.device)
        beta = th.empty((n, 2), dtype=th.float32, device=za.device)
        w1, b1, w2 = w1.contiguous(), b1.contiguous(), w2.reshape(-1).contiguous()
        sv, sd = _seed_args(seed)
        L.check(lib.dg_attention_fwd_f32(za.data_ptr(), za.stride(0), zb.data_ptr(), zb.stride(0), n, d, L.ptr(w1), L.ptr(b1),
                                         L.ptr(w2), w1.shape[0], float(p), sv, L.ptr(sd), L.ptr(out), out.stride(0), L.ptr(beta),
                                         L.stream()), 'attention_fwd')
        ctx.save_for_backward(za, zb, w1, b1, w2, sd)
        ctx.cfg = (float(p), sv)
        ctx.set_materialize_grads(False)           # an unused beta output arrives as None, not as a zero matrix
        return out, beta

    @staticmethod
    def backward(ctx, dout, dbeta):
        lib = L.load()
        za, zb, w1, b1, w2, sd = ctx.saved_tensors
        p, sv = ctx.cfg
        n, d = za.shape
        h = w1.shape[0]
        if dout is None:
            dout = th.zeros((n, d), dtype=th.float32, device=za.device)
        dout = dout if _rows_ok(dout) else dout.contiguous()
        if dbeta is not None:
            dbeta = dbeta.reshape(n, 2).contiguous()
        dza = th.empty((n, d), dtype=th.float32, device=za.device) if ctx.needs_input_grad[0] else None
        dzb = th.empty((n, d), dtype=th.float32, device=za.device) if ctx.needs_input_grad[1] else None
        params = th.empty(16 * d + 32, dtype=th.float32, device=za.device)
        ws = L.workspace(lib.dg_attention_bwd_workspace_bytes(n, d), za.device)
        L.check(lib.dg_attention_bwd_f32(za.data_ptr(), za.stride(0), zb.data_ptr(), zb.stride(0), n, d, L.ptr(w1), L.ptr(b1),
                                         L.ptr(w2), h, p, sv, L.ptr(sd), dout.data_ptr(), dout.stride(0), L.ptr(dbeta),
                                         L.ptr(dza), d, L.ptr(dzb), d, L.ptr(params), L.ptr(ws), ws.numel(), L.stream()),
                'attention_bwd')
        dw1 = params[:16 * d].view(16, d)[:h]
        db1 = params[16 * d:16 * d + h]
        dw2 = params[16 * d + 16:16 * d + 16 + h].view(1, h)
        return dza, dzb, dw1, db1, dw2, None, None


def attention_eligible(za, zb, w1):
    return (za.is_cuda and za.dtype == th.float32 and za.dim() == 2 and za.shape == zb.shape and za.shape[0] > 0
            and _rows_ok(za) and _rows_ok(zb) and za.shape[1] <= ATT_MAX_WIDTH and w1.shape[0] <= ATT_MAX_HIDDEN)


def attention_fuse(za, zb, w1, b1, w2, p=0.0, training=False, seed=None):
    """(sum_k beta_k z_k [n, d], beta [n, 2]) for the two views za, zb; beta = dropout(softmax_k(w2 . tanh(W1 z_k + b1)))."""
    if not za.is_cuda:
        raise RuntimeError('dreamgnn_b200.attention_fuse needs CUDA tensors (no CPU fallback)')
    p = float(p) if training else 0.0
    if seed is None:
        seed = fresh_seed(za.device) if p > 0 else 0
    return AttentionFunction.apply(za, zb, w1, b1, w2, p, seed)


# ------------------------------------------------------------------------------------------------
# SpMM
# ------------------------------------------------------------------------------------------------
# Optional launch log for bench.py's roofline: when PROFILE is a list every SpMM launch appends
# (tag, nnz, n_rows, n_cols, d, element bytes, valued, start event, end event).
PROFILE = None


def _spmm_raw(csr, x, src_scale=None, dst_scale=None, bias=None, flags=0, out=None, tag='spmm'):
    lib = L.load()
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.contiguous()
    if x.shape[0] != csr.n_cols:
        raise ValueError('spmm: x has %d rows, graph has %d source nodes' % (x.shape[0], csr.n_cols))
    d = x.shape[1]
    if out is None:
        out = th.empty((csr.n_rows, d), dtype=th.float32, device=x.device)
    for nm, t, n in (('src_scale', src_scale, csr.n_cols), ('dst_scale', dst_scale, csr.n_rows), ('bias', bias, d)):
        if t is not None and (t.numel() != n or t.dtype != th.float32):
            raise ValueError('spmm: %s must be fp32 with %d elements' % (nm, n))
    if csr.split_T and csr.nnz and not (flags & SPMM_ACCUMULATE):
        # skewed graph: chunk partials (src_scale applied), then the per-row sum of the chunks with the epilogue
        if csr._plan is None:
            v_indptr, comb_indptr, n_v = split_plan(csr.indptr, csr.n_rows, csr.nnz, csr.split_T)
            chunks = CSR(v_indptr, csr.indices, csr.eid, csr.vals, n_v, csr.n_cols)
            comb = CSR(comb_indptr, th.arange(n_v, dtype=I32, device=x.device), None, None, csr.n_rows, n_v)
            csr._plan = (chunks, comb)
        chunks, comb = csr._plan
        part = _spmm_raw(chunks, x, src_scale, None, None, 0, tag=tag)
        return _spmm_raw(comb, part, None, dst_scale, bias, flags & SPMM_RELU, out=out, tag=tag + '.combine')
    if 0 < csr.n_rows <= SPMM_ROWSPLIT_MAX_ROWS and csr.nnz >= SPMM_ROWSPLIT_MIN_AVG_LEN * csr.n_rows:
        flags |= SPMM_ROWSPLIT
    operand = csr.n_cols * d * x.element_size()
    if operand >= SPMM_PREFETCH_MIN_BYTES or (d * x.element_size() >= 1024 and operand >= SPMM_PREFETCH_MIN_BYTES_WIDE):
        flags |= SPMM_PREFETCH
    args = (L.ptr(csr.indptr), L.ptr(csr.indices), L.ptr(csr.vals), L.ptr(src_scale), L.ptr(dst_scale), L.ptr(bias))
    if PROFILE is not None:
        ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        ev0.record()
    if x.dtype == th.float32:
        rc = lib.dg_spmm_csr_f32(*args, x.data_ptr() if x.is_cuda else L.ptr(x), x.stride(0), L.ptr(out),
                                 out.stride(0), csr.n_rows, d, flags, L.stream())
    elif x.dtype == th.bfloat16:
        rc = lib.dg_spmm_csr_bf16(*args, x.data_ptr() if x.is_cuda else L.ptr(x), x.stride(0), L.ptr(out),
                                  out.stride(0), csr.n_rows, d, flags, L.stream())
    else:
        raise TypeError('spmm: x must be float32 or bfloat16')
    L.check(rc, 'spmm_csr')
    if PROFILE is not None:
        ev1.record()
        PROFILE.append((tag, csr.nnz, csr.n_rows, csr.n_cols, d, x.element_size(), csr.vals is not None, ev0, ev1))
    return out


class SpMMFunction(th.autograd.Function):
    """out = act(dst_scale * (A_vals @ (src_scale * x)) + bias); grad flows to x and bias only."""

    @staticmethod
    def forward(ctx, x, bias, csr, src_scale, dst_scale, relu, tag):
        out = _spmm_raw(csr, x, src_scale, dst_scale, bias, SPMM_RELU if relu else 0, tag=tag)
        ctx.csr, ctx.relu, ctx.has_bias, ctx.tag = csr, relu, bias is not None, tag
        ctx.save_for_backward(src_scale, dst_scale, out if relu else None)
        ctx.x_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        src_scale, dst_scale, out = ctx.saved_tensors
        dout = dout.contiguous()
        want_bias = ctx.has_bias and ctx.needs_input_grad[1]
        dbias = None
        if ctx.relu:                                      # relu mask (+ bias gradient) in one pass over dout
            dout, s = colsum(dout, gate=out, want_masked=True)
            dbias = s if want_bias else None
        elif want_bias:
            dbias = colsum(dout)
        dx = None
        if ctx.needs_input_grad[0]:
            # d x[j] = src_scale[j] * sum_{i : j in row i} vals * dst_scale[i] * dout[i]  -> transposed CSR
            if ctx.x_dtype == th.bfloat16:
                dout = dout.to(th.bfloat16)          # bf16 storage path: gather the gradient rows in bf16 too
            dx = _spmm_raw(ctx.csr.transpose(), dout, dst_scale, src_scale, None, 0, tag=ctx.tag + '.bwd')
            if ctx.x_dtype != th.float32:
                dx = dx.to(ctx.x_dtype)
        return dx, dbias, None, None, None, None, None


def spmm(csr, x, src_scale=None, dst_scale=None, bias=None, relu=False, tag='spmm'):
    if not x.is_cuda:
        raise RuntimeError('dreamgnn_b200.spmm needs CUDA tensors (no CPU fallback)')
    return SpMMFunction.apply(x, bias, csr, src_scale, dst_scale, relu, tag)


# ------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------
class PairGraph:
    """Scored (drug, disease) pairs in label order + the two segment structures the deterministic
    backward needs (CSR by drug and by disease whose `indices` are pair ids).

    `processing_order()` is the optional by-drug walk (DG_DECODER_ORDER=by-drug): the pairs sorted by drug once
    per graph (stable, so the order is a pure function of the pair list), which turns the pd-row gather into a
    re-read of the row the previous pair used and leaves only the ps rows (n_dst x 512 B) as a random gather that
    L2 holds. The kernel scatters the logits back to label order (`perm`), so callers never see the difference."""

    def __init__(self, src, dst, n_src, n_dst):
        self.src, self.dst = _i32(src, 'src'), _i32(dst, 'dst')
        self.n_src, self.n_dst = int(n_src), int(n_dst)
        self._by_src = self._by_dst = self._order = None
        self._slots = {}

    @property
    def n_pairs(self):
        return int(self.src.numel())

    def _segments(self, key, n):
        ids = th.arange(self.n_pairs, dtype=I32, device=key.device)
        c = CSR.from_coo(key, ids, n, max(self.n_pairs, 1))
        return c

    def by_src(self):
        if self._by_src is None:
            self._by_src = self._segments(self.src, self.n_src)
        return self._by_src

    def by_dst(self):
        if self._by_dst is None:
            self._by_dst = self._segments(self.dst, self.n_dst)
        return self._by_dst

    SLOT_GROUP = 16                       # pairs per epilogue thread of the backward kernel (decoder_tc.cu)

    def source_slots(self, by_drug=False):
        """Plan of the source-node segment sum fused into the backward's epilogue: (pair_slot int32 [E], n_slots,
        CSR over slots by source node). In processing order, a new slot starts at every multiple of 16 pairs and wherever
        the source node changes -- label order is sorted by drug inside each label class, so almost every aligned group of
        16 pairs is one slot. Built once per pair graph (one host read of the slot count)."""
        key = bool(by_drug)
        if key not in self._slots:
            src = self.processing_order()[1] if by_drug else self.src
            e = self.n_pairs
            pos = th.arange(e, device=src.device)
            new = (pos % self.SLOT_GROUP) == 0
            new[1:] |= src[1:] != src[:-1]
            slot = (th.cumsum(new.to(th.int32), 0, dtype=th.int32) - 1).contiguous()
            n_slots = int(slot[-1]) + 1
            seg = CSR.from_coo(src[new], th.arange(n_slots, dtype=I32, device=src.device), self.n_src, n_slots)
            self._slots[key] = (slot, n_slots, seg)
        return self._slots[key]

    def processing_order(self):
        """(perm, src_p, dst_p, seg_src, seg_dst): perm[i] = label position of the i-th processed pair (pairs
        sorted by drug, ties in label order), the endpoint arrays in that order, and the two segment CSRs over
        processing positions (seg_src's indices are simply 0..E-1: each drug's pairs are contiguous)."""
        if self._order is None:
            perm = self._segments(self.src, self.n_src).indices
            long_perm = perm.long()
            src_p, dst_p = self.src[long_perm].contiguous(), self.dst[long_perm].contiguous()
            self._order = (perm, src_p, dst_p, self._segments(src_p, self.n_src), self._segments(dst_p, self.n_dst))
        return self._order


def _decoder_by_drug():
    """DG_DECODER_ORDER=by-drug walks the pairs sorted by drug instead of in label order (see
    PairGraph.processing_order; measured neutral while the kernels are not gather-bound, so label order, which
    needs no permutation, is the default)."""
    return os.environ.get('DG_DECODER_ORDER', 'label') == 'by-drug'


class DecoderFunction(th.autograd.Function):
    @staticmethod
    def forward(ctx, pd, ps, w2, b2, w3, b3, pairs, p, seed, save):
        lib = L.load()
        pd, ps, w2, b2, w3, b3 = (t.contiguous() for t in (pd, ps, w2, b2, w3, b3))
        if pd.shape[1] != DEC_H1 or ps.shape[1] != DEC_H1 or tuple(w2.shape) != (DEC_H2, DEC_H1):
            raise ValueError('decoder hidden widths are fixed at 128 / 64 (layers.py:349-351)')
        if pd.shape[0] != pairs.n_src or ps.shape[0] != pairs.n_dst:
            raise ValueError('decoder: node counts do not match the pair graph')
        e = pairs.n_pairs
        # `seed` is a Python int, or a 1-element int64 CUDA tensor (read on device: CUDA-graph friendly)
        seed_dev = seed if isinstance(seed, th.Tensor) else None
        seed_val = 0 if seed_dev is not None else int(seed)
        out = th.empty(e, dtype=th.float32, device=pd.device)
        z2 = th.empty((e, DEC_H2), dtype=th.float32, device=pd.device) if save else None
        by_drug = e and _decoder_by_drug()
        perm, src_p, dst_p = pairs.processing_order()[:3] if by_drug else (None, pairs.src, pairs.dst)
        ctx.by_drug = by_drug
        L.check(lib.dg_decoder_fwd_f32(L.ptr(src_p), L.ptr(dst_p), L.ptr(perm), e, L.ptr(pd, th.float32, 'pd'),
                                       L.ptr(ps, th.float32, 'ps'), L.ptr(w2, th.float32), L.ptr(b2, th.float32),
                                       L.ptr(w3, th.float32), L.ptr(b3, th.float32), float(p), seed_val, L.ptr(seed_dev), L.ptr(out),
                                       L.ptr(z2), L.stream()), 'decoder_fwd')
        ctx.pairs, ctx.p, ctx.seed, ctx.seed_dev = pairs, float(p), seed_val, seed_dev
        ctx.save_for_backward(pd, ps, w2, w3, z2)
        return out.unsqueeze(1)

    @staticmethod
    def backward(ctx, dout):
        lib = L.load()
        pd, ps, w2, w3, z2 = ctx.saved_tensors
        if z2 is None:
            raise RuntimeError('decoder forward was run without saving z2 (inference mode)')
        pairs = ctx.pairs
        e = pairs.n_pairs
        dev = pd.device
        dout = dout.reshape(-1).contiguous()
        dz1 = th.empty((max(e, 1), DEC_H1), dtype=th.float32, device=dev)
        dw2 = th.empty((DEC_H2, DEC_H1), dtype=th.float32, device=dev)
        db2 = th.empty(DEC_H2, dtype=th.float32, device=dev)
        dw3 = th.empty((1, DEC_H2), dtype=th.float32, device=dev)
        db3 = th.empty(1, dtype=th.float32, device=dev)
        ws = L.workspace(lib.dg_decoder_bwd_workspace_bytes(e), dev)
        if ctx.by_drug:
            perm, src_p, dst_p, seg_src, seg_dst = pairs.processing_order()
        else:
            perm, src_p, dst_p, seg_dst = None, pairs.src, pairs.dst, pairs.by_dst()
        # the segment sum by source node rides in the kernel's epilogue (one partial row per run of equal source inside
        # aligned 16-pair groups) unless the SIMT kernels or DG_DECODER_SEG=spmm are selected
        fuse = (e > 0 and ctx.needs_input_grad[0] and os.environ.get('DG_DECODER', 'tc') != 'simt'
                and os.environ.get('DG_DECODER_SEG', 'fused') == 'fused')
        pair_slot = slot_rows = seg_slots = None
        if fuse:
            pair_slot, n_slots, seg_slots = pairs.source_slots(ctx.by_drug)
            slot_rows = th.empty((n_slots, DEC_H1), dtype=th.float32, device=dev)
        L.check(lib.dg_decoder_bwd_f32(L.ptr(src_p), L.ptr(dst_p), L.ptr(perm), e, L.ptr(pd), L.ptr(ps), L.ptr(w2),
                                       L.ptr(w3), ctx.p, ctx.seed, L.ptr(ctx.seed_dev), L.ptr(z2), L.ptr(dout, th.float32, 'dout'),
                                       L.ptr(dz1), L.ptr(dw2), L.ptr(db2), L.ptr(dw3), L.ptr(db3), L.ptr(pair_slot),
                                       L.ptr(slot_rows), L.ptr(ws), ws.numel(), L.stream()), 'decoder_bwd')
        # scatter of dz1 (processing order) into node gradients = segment sums in fixed order (no atomics)
        if fuse:
            dpd = _spmm_raw(seg_slots, slot_rows, tag='decoder.slots')
        elif ctx.needs_input_grad[0]:
            dpd = _spmm_raw(pairs.processing_order()[3] if ctx.by_drug else pairs.by_src(), dz1, tag='decoder.seg')
        else:
            dpd = None
        dps = _spmm_raw(seg_dst, dz1, tag='decoder.seg') if ctx.needs_input_grad[1] else None
        return dpd, dps, dw2, db2, dw3, db3, None, None, None, None


def decoder_saved_mask(out):
    """Diagnostic: the hidden-2 ReLU mask the backward of `out = decoder_mlp(...)` will differentiate through, as one
    int64 per pair (bit j <=> z2[e, j] > 0), read from the state the forward saved on the autograd graph."""
    fn = out.grad_fn
    while fn is not None and not hasattr(fn, 'saved_tensors'):
        fn = fn.next_functions[0][0] if fn.next_functions else None
    if fn is None or len(fn.saved_tensors) < 5 or fn.saved_tensors[4] is None:
        raise RuntimeError('not the output of a decoder forward that saved its state')
    z2 = fn.saved_tensors[4]
    bit = th.arange(DEC_H2, device=z2.device, dtype=th.int64)
    bits = th.empty(z2.shape[0], dtype=th.int64, device=z2.device)
    for c0 in range(0, z2.shape[0], 1 << 22):                 # chunked: the boolean / int64 temporaries are 8x z2
        bits[c0:c0 + (1 << 22)] = ((z2[c0:c0 + (1 << 22)] > 0).to(th.int64) << bit).sum(1)
    return bits


def decoder_mlp(pd, ps, w2, b2, w3, b3, pairs, p=0.0, seed=0, training=False):
    if not pd.is_cuda:
        raise RuntimeError('dreamgnn_b200.decoder_mlp needs CUDA tensors (no CPU fallback)')
    save = th.is_grad_enabled() and any(t.requires_grad for t in (pd, ps, w2, b2, w3, b3))
    return DecoderFunction.apply(pd, ps, w2, b2, w3.reshape(1, -1), b3, pairs, p if training else 0.0, seed, save)


# ------------------------------------------------------------------------------------------------
# kNN graphs
# ------------------------------------------------------------------------------------------------
def topk_rows(sim, k):
    """Row-wise top-k columns of a float64 block (value desc, column asc), returned ascending."""
    lib = L.load()
    if sim.dtype != th.float64 or sim.dim() != 2 or sim.stride(1) != 1:
        raise TypeError('topk_rows: float64 row-major matrix expected')
    out = th.empty((sim.shape[0], k), dtype=I32, device=sim.device)
    L.check(lib.dg_topk_rows_f64(L.ptr(sim) if sim.is_contiguous() else sim.data_ptr(), sim.shape[0], sim.shape[1],
                                 sim.stride(0), int(k), L.ptr(out), L.stream()), 'topk_rows')
    return out


def knn_graph_from_neighbors(nbr):
    """[n,k] neighbour lists -> CSR of D^-1 (A + A^T + I) with fp32 values (row/col sorted)."""
    lib = L.load()
    nbr = _i32(nbr, 'nbr')
    n, k = nbr.shape
    dev = nbr.device
    nnz_max = 2 * n * k + n
    indptr = th.empty(n + 1, dtype=I32, device=dev)
    row = th.empty(nnz_max, dtype=I32, device=dev)
    col = th.empty(nnz_max, dtype=I32, device=dev)
    val = th.empty(nnz_max, dtype=th.float32, device=dev)
    nnz = th.zeros(1, dtype=I32, device=dev)
    ws = L.workspace(lib.dg_knn_graph_workspace_bytes(n, k), dev)
    L.check(lib.dg_knn_graph_from_neighbors(L.ptr(nbr), n, k, L.ptr(indptr), L.ptr(row), L.ptr(col), L.ptr(val),
                                            L.ptr(nnz), L.ptr(ws), ws.numel(), L.stream()), 'knn_graph')
    m = int(nnz.item())                      # one sync per graph build (not on the training path)
    csr = CSR(indptr, col[:m].contiguous(), th.arange(m, dtype=I32, device=dev), val[:m].contiguous(), n, n)
    csr.eid_is_slot = True
    return csr, row[:m].contiguous()


# ------------------------------------------------------------------------------------------------
# dense projections on the tcgen05 tensor cores (csrc/gemm_tc.cu)
# ------------------------------------------------------------------------------------------------
import os as _os

GEMM_MIN_MACS = int(_os.environ.get('DG_GEMM_MIN_MACS', str(1 << 26)))   # below this: the one-launch fp32 small_gemm (the tcgen05 kernel has a ~20 us floor)


def gemm_backend():
    """'tcgen05' (default) or 'cublas' (DG_GEMM=cublas: the library stand-in, for A/B comparisons)."""
    return _os.environ.get('DG_GEMM', 'tcgen05')


def gemm(a, b, trans_a=False, trans_b=False, row_scale=None, precision=0):
    """C[r] = diag(row_scale[r]) * op(A[r]) @ op(B[r])^T on the tensor cores, op(A) [M,K], op(B) [N,K].
    a: [M,K] ([K,M] with trans_a) or batched [R,..]; b: [N,K] ([K,N] with trans_b) or batched (a 2-D operand is shared
    by all batches). Transposed operands are read as stored (MN-major tensor-core operands): no copy.
    fp32 in, fp32 out; precision 0 = 3xTF32 (fp32-level accuracy)."""
    lib = L.load()
    if a.dtype != th.float32 or b.dtype != th.float32:
        raise TypeError('gemm: fp32 operands expected')
    batch = max(a.shape[0] if a.dim() == 3 else 1, b.shape[0] if b.dim() == 3 else 1)
    a, b = a.contiguous(), b.contiguous()
    (K, M) = a.shape[-2:] if trans_a else a.shape[-2:][::-1]
    (Kb, N) = b.shape[-2:] if trans_b else b.shape[-2:][::-1]
    if Kb != K:
        raise ValueError('gemm: inner dimensions differ')
    a_b, b_b = int(a.dim() == 3 and batch > 1), int(b.dim() == 3 and batch > 1)
    out = th.empty((batch, M, N) if (a.dim() == 3 or b.dim() == 3) else (M, N), dtype=th.float32, device=a.device)
    ws = L.workspace(lib.dg_gemm_nt_workspace_bytes(M, N, K, batch, a_b or batch == 1, b_b or batch == 1), a.device)
    if row_scale is not None:
        row_scale = row_scale.reshape(-1).to(th.float32).contiguous()
        if row_scale.numel() != batch * M:
            raise ValueError('gemm: row_scale must have batch*M elements')
    L.check(lib.dg_gemm_f32(L.ptr(a), a.shape[-1], M * K if a_b else 0, int(trans_a), L.ptr(b), b.shape[-1], N * K if b_b else 0,
                            int(trans_b), L.ptr(out), N, M * N, M, N, K, batch, L.ptr(row_scale), int(precision), L.ptr(ws),
                            ws.numel(), L.stream()), 'gemm')
    return out


def gemm_nt(a, b, row_scale=None, precision=0):
    """C[r] = diag(row_scale[r]) * A[r] @ B[r]^T with a: [M,K] or [R,M,K]; b: [N,K] or [R,N,K]."""
    return gemm(a, b, False, False, row_scale, precision)


class ProjectFunction(th.autograd.Function):
    """y[r] = x @ w[r] for the R relation weights at once (x [M,K] shared, w [R,K,N]) -> [R,M,N]."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return gemm(x, w, trans_b=True)                         # w[r] is [K,N]: read MN-major, no transpose

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dw = None
        dy = dy.contiguous()
        if ctx.needs_input_grad[0]:
            # dx = sum_r dy[r] @ w[r]^T = [dy_0 | dy_1 | ..] @ [w_0 | w_1 | ..]^T : one GEMM with K' = R*N
            R, M, N = dy.shape
            dx = gemm_nt(dy.permute(1, 0, 2).reshape(M, R * N), w.permute(1, 0, 2).reshape(w.shape[1], R * N))
        if ctx.needs_input_grad[1]:
            # dw[r] = x^T @ dy[r]: op(A) = x^T with x [M,K] as stored ([K',M'] for this product), op(B[r]) = dy[r]^T with
            # dy[r] [M,N] as stored ([K',N']); long K' = M -> split-K inside. No transposed copies.
            dw = gemm(x, dy, trans_a=True, trans_b=True)
        return dx, dw


class LinearFunction(th.autograd.Function):
    """y = x @ w^T + b (nn.Linear) with the three GEMMs on the tensor cores: both forward operands are K-major as
    stored; dx = dy @ w reads w as stored ([K',N']); dw = dy^T @ x reads dy and x as stored ([K',M'], [K',N'])."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        y = gemm_nt(x, w)
        if b is not None:
            y += b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = gemm(dy, w, trans_b=True) if ctx.needs_input_grad[0] else None
        dw = gemm(dy, x, trans_a=True, trans_b=True) if ctx.needs_input_grad[1] else None
        db = colsum(dy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


SMALL_GEMM = _os.environ.get('DG_SMALL_GEMM', '1') != '0'      # 0: the library (torch / cuBLAS) for the small layers, for A/B runs


def small_gemm(a, b, trans_a=False, trans_b=False, bias=None, reduce_batch=False):
    """C[r] = op(A[r]) @ op(B[r])^T (+ bias) in plain fp32 for the small dense layers (csrc/small_gemm.cu): same operand
    conventions as `gemm` (a: [M,K] or [K,M] with trans_a; b: [N,K] or [K,N] with trans_b; 3-D = batched, a 2-D operand is
    shared), any row stride, one launch (split-K and, with reduce_batch, the sum over the batch are added in the kernel)."""
    lib = L.load()
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError('dreamgnn_b200.small_gemm needs CUDA tensors (no CPU fallback)')
    if a.dtype != th.float32 or b.dtype != th.float32:
        raise TypeError('small_gemm: fp32 operands expected')
    rows_ok = lambda t: t.stride(-1) == 1 and t.stride(-2) >= t.shape[-1]          # any row stride; no broadcast views
    a = a if rows_ok(a) else a.contiguous()
    b = b if rows_ok(b) else b.contiguous()
    batch = max(a.shape[0] if a.dim() == 3 else 1, b.shape[0] if b.dim() == 3 else 1)
    (K, M) = a.shape[-2:] if trans_a else a.shape[-2:][::-1]
    (Kb, N) = b.shape[-2:] if trans_b else b.shape[-2:][::-1]
    if Kb != K:
        raise ValueError('small_gemm: inner dimensions differ (%d vs %d)' % (K, Kb))
    sa = a.stride(0) if (a.dim() == 3 and batch > 1 and a.shape[0] > 1) else 0
    sb = b.stride(0) if (b.dim() == 3 and batch > 1 and b.shape[0] > 1) else 0
    batched_out = (a.dim() == 3 or b.dim() == 3) and not reduce_batch
    out = th.empty((batch, M, N) if batched_out else (M, N), dtype=th.float32, device=a.device)
    if bias is not None:
        bias = bias.reshape(-1).to(th.float32).contiguous()
        if bias.numel() != N:
            raise ValueError('small_gemm: bias must have N elements')
    ws = L.workspace(lib.dg_small_gemm_workspace_bytes(M, N, K, batch), a.device)
    n_tk = int(lib.dg_small_gemm_tickets(M, N, K, batch))
    tk = _tickets(a.device, n_tk) if n_tk else None
    if n_tk and tk is None:
        raise RuntimeError('small_gemm: a split-K product (K >= 2048 with few output tiles) needs the ticket pool, which is only '
                           'allocated outside CUDA-graph capture: run the step once eagerly before capturing it')
    L.check(lib.dg_small_gemm_f32(a.data_ptr(), a.stride(-2), sa, int(trans_a), b.data_ptr(), b.stride(-2), sb, int(trans_b),
                                  L.ptr(bias), L.ptr(out), N, M * N, M, N, K, batch, int(bool(reduce_batch)), L.ptr(ws), ws.numel(),
                                  tk, L.stream()), 'small_gemm')
    return out


class SmallProjectFunction(th.autograd.Function):
    """y[r] = x @ w[r] for the R relation weights at once below the tensor-core kernel's size threshold: one small_gemm launch
    each for y, dx (summed over the relations inside the kernel) and dw."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return small_gemm(x, w, trans_b=True)                                    # w[r] is [K, N] as stored

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy if dy.stride(-1) == 1 else dy.contiguous()
        dx = small_gemm(dy, w, reduce_batch=True) if ctx.needs_input_grad[0] else None         # sum_r dy[r] @ w[r]^T
        dw = small_gemm(x, dy, trans_a=True, trans_b=True) if ctx.needs_input_grad[1] else None  # x^T @ dy[r]
        return dx, dw


class SmallLinearFunction(th.autograd.Function):
    """y = x @ w^T + b below the tensor-core kernel's size threshold: one small_gemm launch each for y (bias in the
    epilogue), dx and dw (split-K added inside the kernel); the bias gradient is the deterministic one-pass column sum.
    (The library: SIMT sgemm + bias epilogue launch forward, sgemm + split-K reduction launches backward.)"""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return small_gemm(x, w, bias=b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy if dy.stride(-1) == 1 else dy.contiguous()
        dx = small_gemm(dy, w, trans_b=True) if ctx.needs_input_grad[0] else None
        dw = small_gemm(dy, x, trans_a=True, trans_b=True) if ctx.needs_input_grad[1] else None
        db = colsum(dy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def linear(x, w, b=None):
    """F.linear(x, w, b) for 2-D x; tcgen05 3xTF32 above GEMM_MIN_MACS, the one-launch fp32 small_gemm below."""
    if not x.is_cuda:
        raise RuntimeError('dreamgnn_b200.linear needs CUDA tensors (no CPU fallback)')
    if (gemm_backend() == 'tcgen05' and x.dim() == 2 and x.dtype == th.float32 and w.dtype == th.float32
            and x.shape[0] * w.shape[0] * w.shape[1] >= GEMM_MIN_MACS):
        return LinearFunction.apply(x, w, b)
    if x.dim() == 2 and x.dtype == th.float32 and w.dtype == th.float32 and gemm_backend() != 'cublas' and SMALL_GEMM:
        return SmallLinearFunction.apply(x, w, b)
    return th.nn.functional.linear(x, w, b)                    # DG_GEMM=cublas (A/B runs), or not a 2-D fp32 product


def project(x, w):
    """x [M,K] @ w [R,K,N] -> [R,M,N]; tcgen05 3xTF32 for the large projections, the fp32 small_gemm below GEMM_MIN_MACS."""
    if not x.is_cuda:
        raise RuntimeError('dreamgnn_b200.project needs CUDA tensors (no CPU fallback)')
    R, K, N = w.shape
    macs = x.shape[0] * K * N * R
    if gemm_backend() == 'tcgen05' and x.dtype == th.float32 and macs >= GEMM_MIN_MACS:
        return ProjectFunction.apply(x, w)
    if gemm_backend() != 'cublas' and SMALL_GEMM and x.dtype == th.float32 and w.dtype == th.float32 and x.dim() == 2:
        return SmallProjectFunction.apply(x, w)
    return th.matmul(x.unsqueeze(0), w)
