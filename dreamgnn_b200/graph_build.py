"""On-device construction of the hot path's graph inputs -- the pieces of the reference's
`DrugDataLoader` that the north-star path owns (citations into /root/reference/data_loader.py):

  * `generate_enc_graph`   _generate_enc_graph + _calc_norm (data_loader.py:400-490)
  * `generate_dec_graph`   _generate_dec_graph (data_loader.py:492-509)
  * `create_similarity_graph`          _create_similarity_graph (data_loader.py:278-310) + utils.normalize /
                                       sparse_mx_to_torch_sparse_tensor (utils.py:11-27)
  * `create_feature_similarity_graph`  _create_feature_similarity_graph (data_loader.py:312-344)

Index work (CSR, degrees, neighbour sets, symmetrisation) and the fp32 normaliser / adjacency values
are bit-exact with the reference; the float64 cosine-similarity GEMM is a plain library GEMM (cuBLAS
DGEMM through torch.mm) whose last-ulp rounding may differ from numpy's BLAS -- neighbour sets agree
on tie-free inputs (hard part 10).
"""
import numpy as np
import torch as th

from . import ops
from .graph import HeteroGraph, heterograph
from .utils import to_etype_name


def _dev_index(x, device):
    if isinstance(x, th.Tensor):
        return x.to(device=device, dtype=th.int64)
    return th.as_tensor(np.asarray(x), dtype=th.int64).to(device)


def _norm_from_degrees(deg):
    """1/sqrt(deg) with 0 -> 0 through the same kernel as the CSR path (bit-exact _calc_norm)."""
    indptr = ops.exclusive_scan_i32(deg.to(th.int32).contiguous())
    out = th.empty(deg.numel(), dtype=th.float32, device=deg.device)
    L = ops.L
    L.check(L.load().dg_degree_norm(L.ptr(indptr), deg.numel(), L.ptr(out), L.stream()), 'degree_norm')
    return out


def generate_enc_graph(rating_pairs, rating_values, num_drug, num_disease, device, symm=True, add_support=True):
    """Encoder heterograph with etypes "<r>" (drug->disease) and "rev-<r>" (disease->drug) per rating
    value, edge order = pair order, and `ci` / `cj` = 1/sqrt(total degree) as [N,1] fp32 node data."""
    device = th.device(device)
    row, col = _dev_index(rating_pairs[0], device), _dev_index(rating_pairs[1], device)
    vals = (rating_values if isinstance(rating_values, th.Tensor) else th.as_tensor(np.asarray(rating_values))).to(device)
    data = {}
    for rating in th.unique(vals).tolist():                    # np.unique order: ascending
        sel = vals == rating
        et = to_etype_name(int(rating) if float(rating).is_integer() else rating)
        data[('drug', et, 'disease')] = (row[sel], col[sel])
        data[('disease', 'rev-%s' % et, 'drug')] = (col[sel], row[sel])
    g = heterograph(data, num_nodes_dict={'drug': num_drug, 'disease': num_disease})
    assert len(row) == sum(g.number_of_edges(et) for et in g.etypes) // 2      # data_loader.py:451
    if add_support:
        drug_blk, dis_blk = g.block('drug'), g.block('disease')
        drug_ci = drug_blk.csr.degree_norm().unsqueeze(1)      # in-degree over all rev-etypes
        dis_ci = dis_blk.csr.degree_norm().unsqueeze(1)
        if symm:
            # out-degree over all etypes = row lengths of the other block's transpose, summed over relations
            drug_out = dis_blk.csr.transpose().degrees().view(dis_blk.num_rel, num_drug).sum(0)
            dis_out = drug_blk.csr.transpose().degrees().view(drug_blk.num_rel, num_disease).sum(0)
            drug_cj = _norm_from_degrees(drug_out).unsqueeze(1)
            dis_cj = _norm_from_degrees(dis_out).unsqueeze(1)
        else:
            drug_cj = th.ones(num_drug, device=device)         # data_loader.py:484-485 (1-D, as there)
            dis_cj = th.ones(num_disease, device=device)
        g.nodes['drug'].data.update({'ci': drug_ci, 'cj': drug_cj})
        g.nodes['disease'].data.update({'ci': dis_ci, 'cj': dis_cj})
    return g


def generate_dec_graph(rating_pairs, num_drug, num_disease, device):
    """Decoder bipartite graph: one etype ('drug','rate','disease'), edges in pair order, duplicates kept."""
    device = th.device(device)
    return heterograph({('drug', 'rate', 'disease'): (_dev_index(rating_pairs[0], device),
                                                      _dev_index(rating_pairs[1], device))},
                       num_nodes_dict={'drug': num_drug, 'disease': num_disease})


def _sparse_from_knn(csr, rows, n):
    idx = th.stack([rows.long(), csr.indices.long()])
    t = th.sparse_coo_tensor(idx, csr.vals, (n, n), device=csr.device, check_invariants=False)
    csr.slot_order = True
    t._dg_csr = csr
    return t


def knn_graph_from_topk(nbr):
    """Neighbour lists [n,k] -> row-normalised symmetric adjacency as a torch sparse COO fp32 tensor
    (sorted by row then column) carrying its CSR sidecar."""
    csr, rows = ops.knn_graph_from_neighbors(nbr)
    return _sparse_from_knn(csr, rows, nbr.shape[0])


def _device_sim(sim_matrix, device):
    sim = th.as_tensor(np.ascontiguousarray(sim_matrix) if not isinstance(sim_matrix, th.Tensor) else sim_matrix)
    return sim.to(device=th.device(device), dtype=th.float64).contiguous()


def _coo_from_keys(keys, vals, n, device):
    """Sorted COO fp32 sparse tensor from int64 keys row * n + col (already unique and ascending)."""
    idx = th.stack([th.div(keys, n, rounding_mode='floor'), keys % n])
    return th.sparse_coo_tensor(idx, vals.to(th.float32), (n, n), device=device, check_invariants=False)


def knn_graph_directed_from_topk(nbr):
    """`self._symm` False (data_loader.py:302 skipped): D^-1 (A + I) of the directed kNN adjacency. Not on the training
    path (one-off loader step): index work in torch on the device, float64 normalisation like utils.normalize."""
    n, k = nbr.shape
    dev = nbr.device
    r = th.arange(n, device=dev, dtype=th.int64).repeat_interleave(k)
    keys = th.cat([r * n + nbr.reshape(-1).long(), th.arange(n, device=dev, dtype=th.int64) * (n + 1)])
    uniq, counts = th.unique(keys, return_counts=True)
    val = counts.double()
    row = th.div(uniq, n, rounding_mode='floor')
    rowsum = th.zeros(n, dtype=th.float64, device=dev).index_add_(0, row, val)
    r_inv = th.where(rowsum > 0, 1.0 / rowsum, th.zeros_like(rowsum))
    return _coo_from_keys(uniq, r_inv[row] * val, n, dev)


def create_similarity_graph(sim_matrix, k, device, symm=True):
    """kNN graph of a given similarity matrix (float64, as scipy.io.loadmat hands it over): data_loader.py:278-310, both
    values of `self._symm`."""
    sim = _device_sim(sim_matrix, device)
    n = sim.shape[0]
    k_actual = min(k, n - 1)
    nbr = ops.topk_rows(sim, k_actual)
    return knn_graph_from_topk(nbr) if symm else knn_graph_directed_from_topk(nbr)


def _max_symmetrize(keys, vals, n):
    """Elementwise max(A, A^T) of a non-negative sparse matrix given as (keys = row * n + col, float64 vals): the scipy
    expression `adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)` (utils.py:133, augmentation.py:395)."""
    tk = (keys % n) * n + th.div(keys, n, rounding_mode='floor')
    allk, allv = th.cat([keys, tk]), th.cat([vals, vals])
    uniq, inv = th.unique(allk, return_inverse=True)
    out = th.zeros(uniq.numel(), dtype=th.float64, device=keys.device).scatter_reduce_(0, inv, allv, reduce='amax', include_self=False)
    return uniq, out


def knn_graph(dis_mat, k, device):
    """utils.knn_graph (utils.py:106-140; dead code in the reference -- imported at data_loader.py:30, never called -- kept
    as part of the kNN builder's option set): binary, max-symmetrised kNN adjacency without self loops or normalisation,
    as a sorted fp32 sparse COO tensor on the device (the reference returns a scipy matrix)."""
    sim = _device_sim(dis_mat, device)
    n = sim.shape[0]
    k_actual = min(k, n - 1)
    dev = sim.device
    if k_actual <= 0:
        i = th.arange(n, device=dev, dtype=th.int64)
        return _coo_from_keys(i * (n + 1), th.ones(n, dtype=th.float64, device=dev), n, dev)
    nbr = ops.topk_rows(sim, k_actual)
    r = th.arange(n, device=dev, dtype=th.int64).repeat_interleave(k_actual)
    keys, vals = _max_symmetrize(r * n + nbr.reshape(-1).long(), th.ones(r.numel(), dtype=th.float64, device=dev), n)
    return _coo_from_keys(keys, vals, n, dev)


def augmented_knn_graph(dis_mat, k, device, dropout_rate=0.1, add_noise=False, noise_scale=0.1, noise=None, keep=None):
    """augmentation.augmented_knn_graph (augmentation.py:341-399; dead code in the reference): kNN graph -> optional value
    noise clipped to [0.01, 1] -> random edge dropout -> max-symmetrise -> + I. `noise` / `keep` inject the draws (tests);
    default: torch's device generator (the reference uses numpy's global one)."""
    base = knn_graph(dis_mat, k, device).coalesce()
    n = base.shape[0]
    dev = base.device
    idx, vals = base.indices(), base.values().double()
    keys = idx[0] * n + idx[1]
    if add_noise or noise is not None:
        z = th.randn(vals.numel(), device=dev, dtype=th.float64) if noise is None else th.as_tensor(noise, device=dev).double()
        vals = th.clamp(vals + z * noise_scale, 0.01, 1.0)
    if keep is not None or dropout_rate > 0:
        if keep is None:
            num_keep = max(1, int(vals.numel() * (1 - dropout_rate)))
            keep = th.randperm(vals.numel(), device=dev)[:num_keep]
        keep = th.as_tensor(keep, device=dev).long()
        keys, vals = keys[keep], vals[keep]
    keys, vals = _max_symmetrize(keys, vals, n)
    allk = th.cat([keys, th.arange(n, device=dev, dtype=th.int64) * (n + 1)])
    allv = th.cat([vals, th.ones(n, dtype=th.float64, device=dev)])
    uniq, inv = th.unique(allk, return_inverse=True)
    out = th.zeros(uniq.numel(), dtype=th.float64, device=dev).index_add_(0, inv, allv)
    idx = th.stack([th.div(uniq, n, rounding_mode='floor'), uniq % n])
    return th.sparse_coo_tensor(idx, out, (n, n), device=dev, check_invariants=False)


def create_feature_similarity_graph(features, k, device, chunk_rows=8192, symm=True):
    """kNN graph of the float64 cosine similarity of `features`; the N x N similarity is produced
    chunk by chunk (never resident as a whole) and reduced to top-k lists on the fly."""
    device = th.device(device)
    f = th.as_tensor(np.ascontiguousarray(features) if not isinstance(features, th.Tensor) else features)
    f = f.to(device=device, dtype=th.float64)
    norms = th.linalg.norm(f, dim=1, keepdim=True)
    norms[norms == 0] = 1e-10
    f = (f / norms).contiguous()
    n = f.shape[0]
    k_actual = min(k, n - 1)
    nbr = th.empty((n, k_actual), dtype=th.int32, device=device)
    ft = f.t().contiguous()
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        nbr[r0:r1] = ops.topk_rows(th.mm(f[r0:r1], ft), k_actual)      # plain library DGEMM (cuBLAS)
    return knn_graph_from_topk(nbr) if symm else knn_graph_directed_from_topk(nbr)


__all__ = ['generate_enc_graph', 'generate_dec_graph', 'create_similarity_graph', 'create_feature_similarity_graph',
           'knn_graph_from_topk', 'knn_graph_directed_from_topk', 'knn_graph', 'augmented_knn_graph', 'HeteroGraph']
