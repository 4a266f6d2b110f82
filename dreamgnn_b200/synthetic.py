"""Seeded synthetic workloads of the shapes BASELINE.json names, generated on device.
There is no network for the real `.mat` files, so every measured configuration is synthetic data of
the named shape with random-initialised weights; bench.py says so in its `data` field.

  * `sparse_workload`  -- the 20M / 400M-association shapes: distinct uniform cells of the drug x
    disease grid (or, with spec['pair_dist'] = 'zipf', Zipf(1.0) drug popularity: the secondary stress set), Bernoulli(pos_rate) labels, row-L2-normalised N(0,1) features (1024 / 768 wide, so
    the unequal-dims branch of GCMCLayer is exercised), k=15 kNN graphs from cosine similarity.
    FGCN's input is the feature matrix (the reference feeds the dense N x N similarity, impossible
    at 100k nodes) -- the one stated departure of these shapes.
  * `dense_workload`   -- lrssl / Gdataset / Cdataset shapes: every cell of the training split is an
    edge of relation "0" or "1" (90 % of the grid), FGCN input = the N x N similarity matrix.
"""
import argparse

import torch as th
import torch.nn.functional as F

from . import graph_build as GB
from .train import TrainState

SHAPES = {
    # name: (n_drug, n_dis, pairs_or_positives, f_drug, f_dis, k)
    'syn20m': dict(kind='sparse', n_drug=100_000, n_dis=50_000, n_pairs=20_000_000, f_drug=1024, f_dis=768, k=15),
    'syn400m': dict(kind='sparse', n_drug=1_000_000, n_dis=500_000, n_pairs=400_000_000, f_drug=1024, f_dis=768, k=15),
    'lrssl': dict(kind='dense', n_drug=763, n_dis=681, n_pos=3051, f_drug=768, f_dis=768, k=4),
    'gdataset': dict(kind='dense', n_drug=593, n_dis=313, n_pos=1933, f_drug=768, f_dis=768, k=4),
    'cdataset': dict(kind='dense', n_drug=663, n_dis=409, n_pos=2532, f_drug=768, f_dis=768, k=4),
}


def scaled(shape, factor):
    """Proportional replica: nodes and pairs scaled by `factor`, widths and k unchanged (same mean degree)."""
    s = dict(SHAPES[shape]) if isinstance(shape, str) else dict(shape)
    if factor == 1:
        return s
    s['n_drug'] = max(int(s['n_drug'] * factor), 64)
    s['n_dis'] = max(int(s['n_dis'] * factor), 64)
    if s['kind'] == 'sparse':
        s['n_pairs'] = max(int(s['n_pairs'] * factor), 1000)
    else:
        s['n_pos'] = max(int(s['n_pos'] * factor * factor), 50)
    return s


def _features(n, f, gen, device):
    return F.normalize(th.randn(n, f, generator=gen, device=device), p=2, dim=1)      # data_loader.py:221-222


def zipf_cells(n_d, n_s, n_draws, gen, device, s=1.0):
    """The secondary stress set of SURVEY.md 8d config 4: `n_draws` (drug, disease) cells with Zipf(s) drug popularity
    (drug of popularity rank r drawn with probability proportional to 1 / r^s; which drug holds which rank is a seeded
    permutation) and uniform diseases; duplicates removed, so fewer than `n_draws` distinct cells come back -- the most
    popular drugs saturate at every disease. Row lengths of the drug-side CSR then span 1 ... n_s."""
    w = th.arange(1, n_d + 1, device=device, dtype=th.float64).pow_(-float(s))
    cdf = th.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    rank = th.searchsorted(cdf, th.rand(n_draws, generator=gen, device=device, dtype=th.float64)).clamp_(max=n_d - 1)
    drug = th.randperm(n_d, generator=gen, device=device)[rank]
    dis = th.randint(0, n_s, (n_draws,), generator=gen, device=device)
    return th.unique(drug * n_s + dis)


def sparse_workload(spec, device, seed=1234, pos_rate=0.01, sim_dim=64):
    device = th.device(device)
    gen = th.Generator(device).manual_seed(seed)
    n_d, n_s = spec['n_drug'], spec['n_dis']
    if spec.get('pair_dist', 'uniform') == 'zipf':
        cells = zipf_cells(n_d, n_s, spec['n_pairs'], gen, device)
    else:
        cells = th.unique(th.randint(0, n_d * n_s, (spec['n_pairs'],), generator=gen, device=device))
    labels = (th.rand(cells.numel(), generator=gen, device=device) < pos_rate).float()
    order = th.argsort(labels, descending=True, stable=True)            # positives first (data_loader.py:170-183)
    cells, labels = cells[order], labels[order].contiguous()
    pairs = ((cells // n_s).to(th.int32), (cells % n_s).to(th.int32))
    del cells, order
    drug_feat, dis_feat = _features(n_d, spec['f_drug'], gen, device), _features(n_s, spec['f_dis'], gen, device)
    emb_d = th.randn(n_d, sim_dim, generator=gen, device=device, dtype=th.float64)
    emb_s = th.randn(n_s, sim_dim, generator=gen, device=device, dtype=th.float64)
    k = spec['k']
    if max(n_d, n_s) > 250_000:
        # exact cosine kNN is an N^2 fp64 GEMM (2 PFLOP at 1M nodes): the 400M-edge shape uses k distinct
        # pseudo-random neighbours per node instead (same degree structure, same adjacency pipeline)
        def rand_knn(n):
            step = th.randint(1, max(n // (k + 1), 2), (n, k), generator=gen, device=device)
            nbr = (th.arange(n, device=device).unsqueeze(1) + th.cumsum(step, 1)) % n
            return GB.knn_graph_from_topk(th.sort(nbr, dim=1).values.to(th.int32))
        graphs = dict(drug_graph=rand_knn(n_d), disease_graph=rand_knn(n_s), drug_feature_graph=rand_knn(n_d),
                      disease_feature_graph=rand_knn(n_s))
        return dict(spec=spec, pairs=pairs, labels=labels, drug_feat=drug_feat, dis_feat=dis_feat,
                    drug_sim_feat=drug_feat, dis_sim_feat=dis_feat, fdim_drug=spec['f_drug'], fdim_disease=spec['f_dis'],
                    **graphs)
    graphs = dict(drug_graph=GB.create_feature_similarity_graph(emb_d, k, device),
                  disease_graph=GB.create_feature_similarity_graph(emb_s, k, device),
                  drug_feature_graph=GB.create_feature_similarity_graph(drug_feat.double(), k, device),
                  disease_feature_graph=GB.create_feature_similarity_graph(dis_feat.double(), k, device))
    return dict(spec=spec, pairs=pairs, labels=labels, drug_feat=drug_feat, dis_feat=dis_feat,
                drug_sim_feat=drug_feat, dis_sim_feat=dis_feat, fdim_drug=spec['f_drug'], fdim_disease=spec['f_dis'],
                **graphs)


def dense_workload(spec, device, seed=0, train_fraction=0.9, sim_rank=32):
    device = th.device(device)
    gen = th.Generator(device).manual_seed(seed)
    n_d, n_s = spec['n_drug'], spec['n_dis']
    total = n_d * n_s
    perm = th.randperm(total, generator=gen, device=device)
    pos, neg = perm[:spec['n_pos']], perm[spec['n_pos']:]
    # one CV fold: 90 % of the positives and 90 % of ALL negatives are training pairs (data_loader.py:146-171)
    pos, neg = pos[:int(pos.numel() * train_fraction)], neg[:int(neg.numel() * train_fraction)]
    cells = th.cat([th.sort(pos).values, th.sort(neg).values])
    labels = th.cat([th.ones(pos.numel(), device=device), th.zeros(neg.numel(), device=device)])
    pairs = ((cells // n_s).to(th.int32), (cells % n_s).to(th.int32))

    def sim(n):
        x = F.normalize(th.randn(n, sim_rank, generator=gen, device=device, dtype=th.float64), dim=1)
        s = (x @ x.t() + 1.0) / 2.0
        s.fill_diagonal_(1.0)
        return s
    sim_d, sim_s = sim(n_d), sim(n_s)
    drug_feat, dis_feat = _features(n_d, spec['f_drug'], gen, device), _features(n_s, spec['f_dis'], gen, device)
    k = spec['k']
    graphs = dict(drug_graph=GB.create_similarity_graph(sim_d, k, device),
                  disease_graph=GB.create_similarity_graph(sim_s, k, device),
                  drug_feature_graph=GB.create_feature_similarity_graph(drug_feat.double(), k, device),
                  disease_feature_graph=GB.create_feature_similarity_graph(dis_feat.double(), k, device))
    return dict(spec=spec, pairs=pairs, labels=labels, drug_feat=drug_feat, dis_feat=dis_feat,
                drug_sim_feat=sim_d.float(), dis_sim_feat=sim_s.float(), fdim_drug=n_d, fdim_disease=n_s, **graphs)


MAT_SEEDS = {'lrssl': 0, 'gdataset': 1, 'cdataset': 2}


def mat_arrays(shape, seed=None, sim_rank=32):
    """A dense shape as a `.mat`-schema dict (data_loader.py:110-129: `didr` [N_dis x N_drug], `drug`, `disease`,
    `drug_embed`, `disease_embed`, `Wrname`), seeded with numpy so that every consumer -- this package's DrugDataLoader on
    the GPU, the reference's on the CPU -- reads the same dataset (SURVEY.md 8d config 1 generator)."""
    import numpy as np
    s = dict(SHAPES[shape]) if isinstance(shape, str) else dict(shape)
    if s['kind'] != 'dense':
        raise ValueError('mat_arrays: dense shapes only')
    rng = np.random.default_rng((MAT_SEEDS.get(shape, 0) if isinstance(shape, str) else 0) if seed is None else seed)
    n_d, n_s = s['n_drug'], s['n_dis']
    cells = rng.choice(n_d * n_s, size=s['n_pos'], replace=False)
    assoc = np.zeros((n_d, n_s), dtype=np.float64)
    assoc[cells // n_s, cells % n_s] = 1.0

    def sim(n):
        x = rng.standard_normal((n, sim_rank))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        m = (x @ x.T + 1.0) / 2.0
        np.fill_diagonal(m, 1.0)
        return m

    names = np.empty((n_d, 1), dtype=object)
    for i in range(n_d):
        names[i, 0] = np.array(['DB%05d' % i])
    return {'didr': assoc.T.copy(), 'drug': sim(n_d), 'disease': sim(n_s),
            'drug_embed': rng.standard_normal((n_d, s['f_drug'])), 'disease_embed': rng.standard_normal((n_s, s['f_dis'])),
            'Wrname': names}


def write_mat(root, shape, loader_name='lrssl', seed=None):
    """Write `<root>/raw_data/drug_data/<loader_name>/<loader_name>.mat` (the relative path DrugDataLoader reads)."""
    import os
    import scipy.io as sio
    d = os.path.join(root, 'raw_data', 'drug_data', loader_name)
    os.makedirs(d, exist_ok=True)
    arrays = mat_arrays(shape, seed)
    sio.savemat(os.path.join(d, loader_name + '.mat'), arrays)
    return arrays


def make_workload(spec, device, seed=1234):
    return sparse_workload(spec, device, seed) if spec['kind'] == 'sparse' else dense_workload(spec, device, seed)


def train_state(w, device):
    spec = w['spec']
    enc = GB.generate_enc_graph(w['pairs'], w['labels'], spec['n_drug'], spec['n_dis'], device)
    dec = GB.generate_dec_graph(w['pairs'], spec['n_drug'], spec['n_dis'], device).int()
    return TrainState(enc.int(), dec, w['labels'], w['drug_graph'], w['disease_graph'], w['drug_feature_graph'],
                      w['disease_feature_graph'], w['drug_feat'], w['dis_feat'], w['drug_sim_feat'], w['dis_sim_feat'])


def model_args(w, device=None, **over):
    """The reference's CLI defaults (train.py:404-448) for the Net constructor."""
    a = dict(layers=3, model_activation='leaky', gcn_agg_units=1024, gcn_out_units=128, dropout=0.3,
             gcn_agg_accum='sum', share_param=True, device=device, nhid1=768, nhid2=128, attention_dropout=0.1,
             rating_vals=[0, 1], src_in_units=w['drug_feat'].shape[1], dst_in_units=w['dis_feat'].shape[1],
             fdim_drug=w['fdim_drug'], fdim_disease=w['fdim_disease'])
    a.update(over)
    return argparse.Namespace(**a)
