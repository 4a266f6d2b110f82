"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / plain torch) of DREAM-GNN's message-passing
hot path. It is the checker for the CUDA path and the `cpu_baseline` / `--impl reference` leg of
bench.py; the product package never imports it.

Every function cites the reference lines it follows (`/root/reference/<file>:<lines>`). The
arithmetic that lives in DGL (not in the reference tree, version unpinned) is restated from DGL's
documented semantics exactly as in `oracle/dgl` -- PARITY UNPINNED for those pieces (see
oracle/README.md); everything here is pinned against the reference run through that stand-in by
`tests/test_oracle_vs_reference.py` (build container) and `tests/golden/*.npz` (everywhere).

Conventions: a model is a flat dict of tensors keyed exactly like the reference `state_dict()`;
the encoder graph is a dict {etype: (src, dst)} with etypes "0","1","rev-0","rev-1" plus
`ci`/`cj` per node type; kNN graphs are (row, col, val, n) COO quadruples.
"""
import math

import numpy as np
import torch as th
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# integer / indexing work (bit-exact contracts)
# --------------------------------------------------------------------------------------------


def csr_from_pairs(rows, cols, n_rows):
    """Canonical CSR of a COO pair list: rows ascending, columns ascending inside a row, ties by
    edge id. Returns (indptr int64 [n_rows+1], indices int64 [E], eid int64 [E]) where `eid[s]`
    is the position in the input list of the edge stored in slot `s`.

    This is what DGL builds lazily inside `update_all` (COO->CSC for the forward, COO->CSR for the
    backward; layers.py:229-232) up to the order inside a row, which does not affect sums beyond
    fp reassociation; equals scipy.sparse.csr_matrix((1,(rows,cols))) with sorted indices when the
    pairs are distinct.
    """
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    eid = np.lexsort((np.arange(rows.size), cols, rows)).astype(np.int64)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n_rows), out=indptr[1:])
    return indptr, cols[eid], eid


def degree_norm(deg):
    """data_loader.py:454-457 `_calc_norm`: 1/sqrt(deg) in float32 with deg==0 -> 0."""
    x = np.asarray(deg).astype('float32')
    x[x == 0.] = np.inf
    return (1. / np.sqrt(x)).astype(np.float32)


def enc_graph_from_pairs(pairs, values, n_drug, n_dis):
    """data_loader.py:400-490 `_generate_enc_graph` with symm=True, add_support=True.

    Returns {'edges': {etype: (src, dst)}, 'ci': {...}, 'cj': {...}, 'num_nodes': {...}}; edge
    order inside an etype is the order of the pair list (DGL keeps insertion order).
    """
    row = np.asarray(pairs[0], dtype=np.int64)
    col = np.asarray(pairs[1], dtype=np.int64)
    values = np.asarray(values)
    edges = {}
    deg_drug = np.zeros(n_drug, dtype=np.int64)
    deg_dis = np.zeros(n_dis, dtype=np.int64)
    for rating in np.unique(values):
        sel = np.where(values == rating)
        et = '0' if rating == 0 else '1' if rating == 1 else str(rating).replace('.', '_')
        edges[et] = (row[sel], col[sel])
        edges['rev-' + et] = (col[sel], row[sel])
        # drug_ci = in-degree of rev-etype, disease_ci = in-degree of etype; with symm the
        # out-degrees used for cj are the same counts (data_loader.py:467-478)
        deg_drug += np.bincount(row[sel], minlength=n_drug)
        deg_dis += np.bincount(col[sel], minlength=n_dis)
    nd, ns = degree_norm(deg_drug)[:, None], degree_norm(deg_dis)[:, None]
    return {'edges': edges, 'num_nodes': {'drug': n_drug, 'disease': n_dis},
            'ci': {'drug': nd, 'disease': ns}, 'cj': {'drug': nd.copy(), 'disease': ns.copy()}}


def topk_neighbors(sim, k):
    """Neighbour *sets* of data_loader.py:291-293: the `k_actual=min(k,N-1)` largest entries of
    each row of `sim` (self included when it is among them). np.argpartition returns an
    arbitrary member among ties at the boundary; this restatement fixes the tie rule to
    (value descending, index ascending), which is what the CUDA kernel implements, and returns
    each row's neighbours sorted by index.
    """
    sim = np.asarray(sim)
    n = sim.shape[0]
    k_actual = min(k, n - 1)
    order = np.lexsort((np.broadcast_to(np.arange(n), sim.shape), -sim), axis=1)
    return np.sort(order[:, :k_actual], axis=1)


def knn_graph_from_neighbors(nbr, n):
    """data_loader.py:294-308 + utils.py:11-17 given the neighbour lists: A (ones) -> A+A^T
    (values 1/2) -> +I -> D^-1 A in float64 -> float32 COO (utils.py:20-27).
    Returns (row, col, val float32) sorted by (row, col)."""
    k = nbr.shape[1]
    r = np.repeat(np.arange(n, dtype=np.int64), k)
    c = nbr.reshape(-1).astype(np.int64)
    keys = np.concatenate([r * n + c, c * n + r, np.arange(n, dtype=np.int64) * (n + 1)])
    uniq, counts = np.unique(keys, return_counts=True)
    row, col = uniq // n, uniq % n
    val = counts.astype(np.float64)
    rowsum = np.bincount(row, weights=val, minlength=n)
    r_inv = np.power(rowsum, -1.0)
    r_inv[np.isinf(r_inv)] = 0.
    return row, col, (r_inv[row] * val).astype(np.float32)


def similarity_knn_graph(sim, k, symm=True):
    """data_loader.py:278-310 `_create_similarity_graph`."""
    n = sim.shape[0]
    if not symm:
        return knn_graph_directed(topk_neighbors(sim, k), n)
    return knn_graph_from_neighbors(topk_neighbors(sim, k), n)


def knn_graph_directed(nbr, n):
    """data_loader.py:294-308 with `self._symm` False: A (ones at (i, nbr)) -> +I -> D^-1 A -> float32 COO, (row, col) sorted."""
    k = nbr.shape[1]
    r = np.repeat(np.arange(n, dtype=np.int64), k)
    c = nbr.reshape(-1).astype(np.int64)
    keys = np.concatenate([r * n + c, np.arange(n, dtype=np.int64) * (n + 1)])
    uniq, counts = np.unique(keys, return_counts=True)
    row, col = uniq // n, uniq % n
    val = counts.astype(np.float64)
    r_inv = np.power(np.bincount(row, weights=val, minlength=n), -1.0)
    r_inv[np.isinf(r_inv)] = 0.
    return row, col, (r_inv[row] * val).astype(np.float32)


def max_symmetrize(row, col, val, n):
    """`adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)` (utils.py:133, augmentation.py:395) = the
    elementwise maximum of A and A^T for non-negative entries. (row, col, val) sorted by (row, col), float64 values."""
    row, col, val = np.asarray(row, dtype=np.int64), np.asarray(col, dtype=np.int64), np.asarray(val, dtype=np.float64)
    keys = np.concatenate([row * n + col, col * n + row])
    vals = np.concatenate([val, val])
    order = np.lexsort((-vals, keys))                       # per key: largest value first
    keys, vals = keys[order], vals[order]
    first = np.concatenate([[True], keys[1:] != keys[:-1]])
    return keys[first] // n, keys[first] % n, vals[first]


def knn_graph_binary(sim, k):
    """utils.py:106-140 `knn_graph` (dead code in the reference; the kNN builder's option set): ones at the k_actual =
    min(k, n - 1) largest entries of each row, max-symmetrised; no self loops added, no normalisation; float32."""
    n = sim.shape[0]
    k_actual = min(k, n - 1)
    if k_actual <= 0:
        i = np.arange(n, dtype=np.int64)
        return i, i, np.ones(n, dtype=np.float32)
    nbr = topk_neighbors(sim, k)
    r = np.repeat(np.arange(n, dtype=np.int64), k_actual)
    row, col, val = max_symmetrize(r, nbr.reshape(-1), np.ones(r.size), n)
    return row, col, val.astype(np.float32)


def augmented_knn_graph(sim, k, keep=None, noise=None, noise_scale=0.1):
    """augmentation.py:341-399 with the random draws injected: `noise` ~ N(0,1) per stored entry of the kNN graph (in
    (row, col) order; None = add_noise False), `keep` = indices of the entries kept by the dropout (None = rate 0).
    values + noise * scale clipped to [0.01, 1] -> keep -> max-symmetrise -> + I. Returns (row, col, val float64)."""
    n = sim.shape[0]
    row, col, val = knn_graph_binary(sim, k)
    val = val.astype(np.float64)
    if noise is not None:
        val = np.clip(val + np.asarray(noise, dtype=np.float64) * noise_scale, 0.01, 1.0)
    if keep is not None:
        keep = np.asarray(keep)
        row, col, val = row[keep], col[keep], val[keep]
    row, col, val = max_symmetrize(row, col, val, n)
    keys = np.concatenate([row * n + col, np.arange(n, dtype=np.int64) * (n + 1)])
    vals = np.concatenate([val, np.ones(n)])
    uniq, inv = np.unique(keys, return_inverse=True)
    return uniq // n, uniq % n, np.bincount(inv, weights=vals)


def feature_cosine_similarity(features):
    """data_loader.py:331-339: float64 row-normalised features, X X^T."""
    f = np.asarray(features, dtype=np.float64)
    norms = np.linalg.norm(f, axis=1, keepdims=True)
    norms[norms == 0] = 1e-10
    f = f / norms
    return f @ f.T


def dropout_num_keep(num_edges, rate):
    """augmentation.py:48 / :113: max(1, int(E*(1-rate))) in Python double arithmetic."""
    return max(1, int(num_edges * (1 - rate)))


def edge_dropout_keep(perm, rate):
    """augmentation.py:51-62 / :116-121: kept edge ids, in perm order."""
    perm = np.asarray(perm)
    return perm[:dropout_num_keep(perm.size, rate)]


# --------------------------------------------------------------------------------------------
# floating-point layers (functional; autograd supplies the backward)
# --------------------------------------------------------------------------------------------

_ACT = {
    None: lambda x: x,
    'leaky': lambda x: F.leaky_relu(x, 0.1),       # utils.py:61-62
    'relu': F.relu, 'tanh': th.tanh, 'sigmoid': th.sigmoid, 'softsign': F.softsign,
    'gelu': F.gelu, 'elu': F.elu, 'selu': F.selu,
}


def _idx(x):
    return th.as_tensor(np.asarray(x) if not isinstance(x, th.Tensor) else x).long()


def canonical_etype_order(etypes):
    """DGL sorts canonical etypes (src type, etype, dst type) lexicographically; 'disease' < 'drug'
    so the reverse relations come first: rev-0, rev-1, 0, 1."""
    return sorted(etypes, key=lambda et: (('disease' if et.startswith('rev-') else 'drug'), et))


def copy_u_sum(src, dst, n_dst, h):
    """DGL `update_all(copy_u, sum)` (layers.py:229-232)."""
    out = th.zeros((n_dst, h.shape[1]), dtype=h.dtype)
    return out.index_add(0, _idx(dst), h[_idx(src)])


def gcmc_graph_conv(src, dst, n_dst, feat, weight, cj, ci, p=0.0, training=False):
    """layers.py:169-236 `GCMCGraphConv.forward`: (feat@W) * dropout(cj) -> sum over in-edges -> * ci."""
    if weight is not None:
        feat = feat @ weight                                  # layers.py:220-221, 392
    feat = feat * F.dropout(cj, p, training).view(-1, 1)      # layers.py:224-225
    return copy_u_sum(src, dst, n_dst, feat) * ci             # layers.py:227-234


def gcmc_layer(P, prefix, graph, drug_feat, dis_feat, act='leaky', p=0.0, training=False):
    """layers.py:117-143 `GCMCLayer.forward` (agg='sum'); both the shared-dims branch (W = att@basis,
    layers.py:120-127) and the per-etype `conv.mods.<etype>.weight` branch (layers.py:86-97)."""
    att, basis = P[prefix + 'att'], P[prefix + 'basis']
    own = (prefix + 'conv.mods.0.weight') in P
    W = (att @ basis.reshape(basis.shape[0], -1)).view(-1, basis.shape[1], basis.shape[2])
    feats = {'drug': drug_feat, 'disease': dis_feat}
    nn_ = graph['num_nodes']
    outs = {'drug': [], 'disease': []}
    ratings = sorted(et for et in graph['edges'] if not et.startswith('rev-'))
    widx = {et: i for i, et in enumerate(ratings)}
    for et in canonical_etype_order(graph['edges']):          # HeteroGraphConv: sorted canonical etypes
        rev = et.startswith('rev-')
        st, dt = ('disease', 'drug') if rev else ('drug', 'disease')
        base = et[4:] if rev else et
        w = P[prefix + 'conv.mods.%s.weight' % et] if own else W[widx[base]]
        src, dst = graph['edges'][et]
        outs[dt].append(gcmc_graph_conv(src, dst, nn_[dt], feats[st], w,
                                        th.as_tensor(graph['cj'][st]), th.as_tensor(graph['ci'][dt]),
                                        p, training))
    f = _ACT[act]
    drug = F.dropout(f(th.stack(outs['drug'], 0).sum(0)), p, training)       # layers.py:134-135
    dis = F.dropout(f(th.stack(outs['disease'], 0).sum(0)), p, training)     # layers.py:137-138
    drug = F.linear(drug, P[prefix + 'ifc.weight'], P[prefix + 'ifc.bias'])  # layers.py:140
    dis = F.linear(dis, P[prefix + 'ufc.weight'], P[prefix + 'ufc.bias'])    # layers.py:141
    return drug, dis


def coo_spmm(coo, x):
    """th.spmm on a (possibly uncoalesced) COO adjacency (layers.py:312)."""
    row, col, val, n = coo
    out = th.zeros((n, x.shape[1]), dtype=x.dtype)
    return out.index_add(0, _idx(row), x[_idx(col)] * th.as_tensor(val).view(-1, 1))


def graph_convolution(P, prefix, x, coo):
    """layers.py:306-316 `GraphConvolution.forward`: spmm(adj, x@W) + b."""
    out = coo_spmm(coo, x @ P[prefix + 'weight'])
    b = P.get(prefix + 'bias')
    return out + b if b is not None else out


def _relu(pre, mask=None):
    """F.relu, or -- when the checker matches masks -- `pre * mask` with the ReLU mask of the implementation under test: a
    gradient is a discontinuous function of that mask, so an entry whose pre-activation lies within fp32 rounding of zero
    makes the comparison ill-posed unless both sides differentiate through the same mask (tests/shapes.py)."""
    return F.relu(pre) if mask is None else pre * th.as_tensor(mask).to(pre.dtype)


def gcn(P, prefix, x, coo, p=0.0, training=False, mask=None):
    """layers.py:245-249 `GCN.forward`."""
    h = F.dropout(_relu(graph_convolution(P, prefix + 'gc1.', x, coo), mask), p, training)
    return graph_convolution(P, prefix + 'gc2.', h, coo)


def fgcn(P, prefix, drug_graph, drug_sim_feat, dis_graph, dis_sim_feat,
         drug_feature_graph=None, dis_feature_graph=None, p=0.0, training=False, masks=None):
    """layers.py:260-285 `FGCN.forward` -> (emb1, emb2, emb1_sim, emb1_feat, emb2_sim, emb2_feat). `masks` (checker only):
    ReLU masks keyed 'gc1.drug.sim', 'gc1.disease.sim', 'gc1.drug.feat', 'gc1.disease.feat', 'fusion.drug', 'fusion.disease'."""
    m = masks or {}
    e1s = gcn(P, prefix + 'FGCN_drug.', drug_sim_feat, drug_graph, p, training, m.get('gc1.drug.sim'))
    e2s = gcn(P, prefix + 'FGCN_disease.', dis_sim_feat, dis_graph, p, training, m.get('gc1.disease.sim'))
    if drug_feature_graph is None or dis_feature_graph is None:
        return e1s, e2s, e1s, None, e2s, None
    e1f = gcn(P, prefix + 'FGCN_drug.', drug_sim_feat, drug_feature_graph, p, training, m.get('gc1.drug.feat'))
    e2f = gcn(P, prefix + 'FGCN_disease.', dis_sim_feat, dis_feature_graph, p, training, m.get('gc1.disease.feat'))
    f1 = _relu(F.linear(th.cat([e1s, e1f], 1), P[prefix + 'drug_fusion.weight'], P[prefix + 'drug_fusion.bias']), m.get('fusion.drug'))
    f2 = _relu(F.linear(th.cat([e2s, e2f], 1), P[prefix + 'disease_fusion.weight'], P[prefix + 'disease_fusion.bias']),
               m.get('fusion.disease'))
    return F.dropout(f1, p, training), F.dropout(f2, p, training), e1s, e1f, e2s, e2f


def attention(P, prefix, z, p=0.0, training=False):
    """layers.py:334-338 `Attention.forward`: z [N,K,D] -> (sum_k beta_k z_k, beta)."""
    w = F.linear(th.tanh(F.linear(z, P[prefix + 'project.0.weight'], P[prefix + 'project.0.bias'])),
                 P[prefix + 'project.2.weight'])
    beta = F.dropout(th.softmax(w, dim=1), p, training)
    return (beta * z).sum(1), beta


def mlp_decoder(P, prefix, src, dst, drug_feat, dis_feat, p=0.0, training=False, masks=None):
    """layers.py:360-379 `MLPDecoder.forward`: cat(h_drug[src], h_dis[dst]) -> 256-128-64-1 MLP. `masks` (checker only):
    ReLU masks 'dec.z1' [E, 128] / 'dec.z2' [E, 64] of the implementation under test (see `_relu`)."""
    m = masks or {}
    x = th.cat([drug_feat[_idx(src)], dis_feat[_idx(dst)]], 1)
    x = F.dropout(_relu(F.linear(x, P[prefix + 'lin1.weight'], P[prefix + 'lin1.bias']), m.get('dec.z1')), p, training)
    x = F.dropout(_relu(F.linear(x, P[prefix + 'lin2.weight'], P[prefix + 'lin2.bias']), m.get('dec.z2')), p, training)
    return F.linear(x, P[prefix + 'lin3.weight'], P[prefix + 'lin3.bias'])


def net_forward(P, enc_graph, dec_pairs, drug_graph, drug_sim_feat, drug_feat,
                dis_graph, dis_sim_feat, dis_feat, drug_feature_graph=None, dis_feature_graph=None,
                layers=3, act='leaky', dropout=0.0, attention_dropout=0.0, training=False, relu_masks=None):
    """model.py:60-103 `Net.forward` -> (pred [E,1], drug_out, drug_sim_out, dis_out, dis_sim_out). `relu_masks`: see fgcn."""
    drug_out = dis_out = None
    for i in range(layers):                                   # model.py:67-76
        d_o, s_o = gcmc_layer(P, 'TGCN.%d.' % i, enc_graph, drug_feat, dis_feat, act, dropout, training)
        drug_out = d_o if i == 0 else drug_out + d_o / float(i + 1)
        dis_out = s_o if i == 0 else dis_out + s_o / float(i + 1)
        drug_feat, dis_feat = d_o, s_o
    drug_sim_out, dis_sim_out = fgcn(P, 'FGCN.', drug_graph, drug_sim_feat, dis_graph, dis_sim_feat,
                                     drug_feature_graph, dis_feature_graph, dropout, training, relu_masks)[:2]
    drug_feats, _ = attention(P, 'attention.', th.stack([drug_out, drug_sim_out], 1), attention_dropout, training)
    dis_feats, _ = attention(P, 'attention.', th.stack([dis_out, dis_sim_out], 1), attention_dropout, training)
    pred = mlp_decoder(P, 'decoder.', dec_pairs[0], dec_pairs[1], drug_feats, dis_feats, dropout, training, relu_masks)
    return pred, drug_out, drug_sim_out, dis_out, dis_sim_out


def common_loss(emb1, emb2):
    """utils.py:87-95."""
    emb1 = F.normalize(emb1 - emb1.mean(0, keepdim=True), p=2, dim=1)
    emb2 = F.normalize(emb2 - emb2.mean(0, keepdim=True), p=2, dim=1)
    return th.mean((emb1 @ emb1.t() - emb2 @ emb2.t()) ** 2)


def training_loss(outputs, labels, beta=0.001):
    """train.py:286-294: BCE-with-logits over all scored pairs + beta * common losses."""
    pred, drug_out, drug_sim_out, dis_out, dis_sim_out = outputs
    rel = F.binary_cross_entropy_with_logits(pred.squeeze(-1), labels)
    return rel + beta * (common_loss(drug_out, drug_sim_out) + common_loss(dis_out, dis_sim_out))


# --------------------------------------------------------------------------------------------
# per-iteration augmentation (default methods) and one full training iteration
# --------------------------------------------------------------------------------------------


def augment_default(enc_graph, knn_graphs, feats, rate=0.1, noise=0.05, sim_noise=0.05, gen=None):
    """augmentation.py:402-489 with aug_methods=['edge_dropout','feature_noise'] in the reference's
    RNG call order: randperm per sorted etype (augmentation.py:35-62), randperm per kNN COO
    (augmentation.py:107-124), then 4x randn_like (augmentation.py:473-489). ci/cj are copied,
    not recomputed (augmentation.py:68-70)."""
    edges = {}
    for et in canonical_etype_order(enc_graph['edges']):
        src, dst = enc_graph['edges'][et]
        keep = th.randperm(len(src), generator=gen)[:dropout_num_keep(len(src), rate)].numpy()
        edges[et] = (np.asarray(src)[keep], np.asarray(dst)[keep])
    g = dict(enc_graph, edges=edges)
    coos = []
    for row, col, val, n in knn_graphs:
        keep = th.randperm(len(val), generator=gen)[:dropout_num_keep(len(val), rate)].numpy()
        coos.append((np.asarray(row)[keep], np.asarray(col)[keep], np.asarray(val)[keep], n))
    scales = (noise, noise, sim_noise, sim_noise)
    # randn_like keeps the input's memory format (the reference's features are column-major, as
    # scipy.io.loadmat hands them over) and torch fills in memory order -> same here
    noisy = [f + th.empty_like(f).normal_(0.0, 1.0, generator=gen) * s for f, s in zip(feats, scales)]
    return g, coos, noisy


def add_random_edges(src, dst, n_src, n_dst, add_rate, candidates, self_loops=False, same_type=False):
    """augmentation.py:127-205 for one relation, with the random draws injected: `candidates` is the sequence of
    (src, dst) pairs `random.randint` would produce, in draw order. num_add = max(1, int(E * add_rate)) (:152); at most
    10 * num_add attempts (:176), each candidate counts as one attempt (:181-193); a candidate is accepted when it is not
    an existing edge and not already accepted (:188). Returns the accepted (src, dst) in acceptance order -- they are
    appended after the existing edges (:197-201)."""
    num_edges = len(src)
    if num_edges == 0 or n_src == 0 or n_dst == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    num_add = max(1, int(num_edges * add_rate))
    have = set(zip(np.asarray(src).tolist(), np.asarray(dst).tolist()))
    new, taken = [], set()
    for attempt, (s, d) in enumerate(zip(np.asarray(candidates[0]).tolist(), np.asarray(candidates[1]).tolist())):
        if len(new) >= num_add or attempt >= num_add * 10:
            break
        if not self_loops and same_type and s == d:
            continue
        if (s, d) not in have and (s, d) not in taken:
            new.append((s, d))
            taken.add((s, d))
    a = np.asarray(new, dtype=np.int64).reshape(-1, 2)
    return a[:, 0], a[:, 1]


def sparse_graph_noise(values, noise, noise_scale):
    """augmentation.py:244-273 with the N(0,1) draw injected: clamp(values + noise * scale, min=0); indices unchanged."""
    return th.clamp(th.as_tensor(values) + th.as_tensor(noise) * noise_scale, min=0.0)


def feature_masking(features, u, mask_rate):
    """augmentation.py:276-308 with the U(0,1) draw injected: features * (u > mask_rate)."""
    return th.as_tensor(features) * (th.as_tensor(u) > mask_rate)


def mix_up_features(features, indices, lam):
    """augmentation.py:311-337 with the permutation and the Beta(alpha, alpha) coefficient injected."""
    f = th.as_tensor(features)
    return lam * f + (1 - lam) * f[_idx(indices)]


def make_adam_state(P):
    return {k: (th.zeros_like(v), th.zeros_like(v)) for k, v in P.items()}


def train_iteration(P, opt, step, enc_graph, dec_pairs, labels, knn_graphs, feats, cfg, gen=None):
    """One reference training iteration (train.py:250-300): augmentation, forward, loss, backward,
    clip_grad_norm_(1.0), Adam(lr, weight_decay). `P` holds leaf tensors with requires_grad;
    `opt` is a torch optimizer over them. Returns the loss value."""
    drug_feat, dis_feat, drug_sim, dis_sim = feats
    g, coos, (df, sf, dsim, ssim) = augment_default(
        enc_graph, knn_graphs, (drug_feat, dis_feat, drug_sim, dis_sim),
        cfg.get('edge_dropout_rate', 0.1), cfg.get('feature_noise_scale', 0.05), 0.05, gen)
    out = net_forward(P, g, dec_pairs, coos[0], dsim, df, coos[1], ssim, sf, coos[2], coos[3],
                      layers=cfg.get('layers', 3), act=cfg.get('act', 'leaky'),
                      dropout=cfg.get('dropout', 0.3), attention_dropout=cfg.get('attention_dropout', 0.1),
                      training=True)
    if cfg.get('common_loss', True):
        loss = training_loss(out, labels, cfg.get('beta', 0.001))
    else:
        loss = F.binary_cross_entropy_with_logits(out[0].squeeze(-1), labels)
    opt.zero_grad()
    loss.backward()
    th.nn.utils.clip_grad_norm_([p for p in P.values() if p.requires_grad], cfg.get('grad_clip', 1.0))
    opt.step()
    return float(loss.detach())


def evaluate_auc(P, enc_graph, dec_pairs, labels, knn_graphs, feats, cfg):
    """evaluation.py:4-74: eval-mode forward on the split's own graphs, un-augmented inputs, then
    sklearn ROC / PR areas on the raw logits."""
    from sklearn import metrics
    drug_feat, dis_feat, drug_sim, dis_sim = feats
    with th.no_grad():
        pred = net_forward(P, enc_graph, dec_pairs, knn_graphs[0], drug_sim, drug_feat, knn_graphs[1],
                           dis_sim, dis_feat, knn_graphs[2], knn_graphs[3],
                           layers=cfg.get('layers', 3), act=cfg.get('act', 'leaky'))[0]
    y_score, y_true = pred.view(-1).numpy(), np.asarray(labels)
    fpr, tpr, _ = metrics.roc_curve(y_true, y_score)
    precision, recall, _ = metrics.precision_recall_curve(y_true, y_score)
    return metrics.auc(fpr, tpr), metrics.auc(recall, precision)


def xavier_uniform_(t):
    fan_in, fan_out = th.nn.init._calculate_fan_in_and_fan_out(t)
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return t.uniform_(-bound, bound)
