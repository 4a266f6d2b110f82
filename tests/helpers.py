"""Shared test helpers: load tests/golden/*.npz into the structures oracle/restate.py uses."""
import os

import numpy as np
import torch as th

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KNN_KEYS = ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph')
CASES = ('tinyA', 'tinyB')
NET_CFG = {'tinyA': dict(layers=3), 'tinyB': dict(layers=2)}


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def enc_graph(g, split='train', prefix=''):
    n_drug, n_dis = g['feat.drug'].shape[0], g['feat.disease'].shape[0]
    edges = {}
    for et in ('0', '1', 'rev-0', 'rev-1'):
        a = g[f'{prefix}{split}.enc.{et}'] if not prefix else g[f'{prefix}enc.{et}']
        edges[et] = (a[0].astype(np.int64), a[1].astype(np.int64))
    return {'edges': edges, 'num_nodes': {'drug': n_drug, 'disease': n_dis},
            'ci': {nt: g[f'{split}.ci.{nt}'] for nt in ('drug', 'disease')},
            'cj': {nt: g[f'{split}.cj.{nt}'] for nt in ('drug', 'disease')}}


def knn_coo(g, key, prefix='knn.'):
    idx, val = g[f'{prefix}{key}.indices'], g[f'{prefix}{key}.values']
    n = g['feat.drug'].shape[0] if key.startswith('drug') else g['feat.disease'].shape[0]
    return idx[0], idx[1], val, n


def params(g, requires_grad=False):
    """state_dict -> tensors; `ifc` aliases `ufc` (share_param, layers.py:61-62)."""
    P = {}
    for k, v in g.items():
        if k.startswith('sd.'):
            P[k[3:]] = th.tensor(v)
    for k in list(P):
        if '.ifc.' in k:
            P[k] = P[k.replace('.ifc.', '.ufc.')]
    if requires_grad:
        for v in set(P.values()):
            v.requires_grad_(True)
    return P


def net_inputs(g, split='train'):
    return dict(
        enc_graph=enc_graph(g, split), dec_pairs=(g[f'{split}.pairs'][0], g[f'{split}.pairs'][1]),
        drug_graph=knn_coo(g, 'drug_graph'), drug_sim_feat=th.tensor(g['mat.drug'], dtype=th.float32),
        drug_feat=th.tensor(g['feat.drug']), dis_graph=knn_coo(g, 'disease_graph'),
        dis_sim_feat=th.tensor(g['mat.disease'], dtype=th.float32), dis_feat=th.tensor(g['feat.disease']),
        drug_feature_graph=knn_coo(g, 'drug_feature_graph'),
        dis_feature_graph=knn_coo(g, 'disease_feature_graph'))


def rel_err(a, b):
    """Norm-wise relative error (SURVEY.md 7 hard part 2: never element-wise rtol)."""
    a = th.as_tensor(a, dtype=th.float64).reshape(-1)
    b = th.as_tensor(b, dtype=th.float64).reshape(-1)
    den = float(b.norm())
    return float((a - b).norm()) / den if den > 0 else float((a - b).norm())


def canon_coo(row, col, val):
    o = np.lexsort((col, row))
    return np.asarray(row)[o], np.asarray(col)[o], np.asarray(val)[o]
