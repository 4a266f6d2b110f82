"""CPU: pin oracle/restate.py against the golden vectors frozen from the reference
(tests/golden/make_golden.py). Runs everywhere, including the GPU box."""
import numpy as np
import pytest
import torch as th

from oracle import restate as R
from tests import helpers as H


@pytest.fixture(scope='module', params=H.CASES)
def gold(request):
    return request.param, H.load_golden(request.param)


def test_enc_graph_and_norms(gold):
    _, g = gold
    for split in ('train', 'test'):
        eg = R.enc_graph_from_pairs(g[f'{split}.pairs'], g[f'{split}.labels'],
                                    g['feat.drug'].shape[0], g['feat.disease'].shape[0])
        for et in ('0', '1', 'rev-0', 'rev-1'):
            np.testing.assert_array_equal(np.stack(eg['edges'][et]), g[f'{split}.enc.{et}'])
        for nt in ('drug', 'disease'):
            np.testing.assert_array_equal(eg['ci'][nt], g[f'{split}.ci.{nt}'])   # bit-exact fp32
            np.testing.assert_array_equal(eg['cj'][nt], g[f'{split}.cj.{nt}'])


def test_csr_matches_scipy(gold):
    import scipy.sparse as sp
    _, g = gold
    src, dst = g['train.enc.0']
    n_dst, n_src = g['feat.disease'].shape[0], g['feat.drug'].shape[0]
    indptr, indices, eid = R.csr_from_pairs(dst, src, n_dst)
    m = sp.csr_matrix((np.ones(len(src)), (dst, src)), shape=(n_dst, n_src))
    m.sort_indices()
    np.testing.assert_array_equal(indptr, m.indptr)
    np.testing.assert_array_equal(indices, m.indices)
    np.testing.assert_array_equal(src[eid], indices)


def test_knn_graphs(gold):
    _, g = gold
    k = int(g['k'])
    sims = {'drug_graph': g['mat.drug'], 'disease_graph': g['mat.disease'],
            'drug_feature_graph': R.feature_cosine_similarity(g['mat.drug_embed']),
            'disease_feature_graph': R.feature_cosine_similarity(g['mat.disease_embed'])}
    for key, sim in sims.items():
        row, col, val = R.similarity_knn_graph(sim, k)
        gr, gc, gv = H.canon_coo(g[f'knn.{key}.indices'][0], g[f'knn.{key}.indices'][1], g[f'knn.{key}.values'])
        np.testing.assert_array_equal(row, gr)
        np.testing.assert_array_equal(col, gc)
        np.testing.assert_array_equal(val, gv)                                    # bit-exact fp32


def test_forward_eval(gold):
    name, g = gold
    P = H.params(g)
    with th.no_grad():
        out = R.net_forward(P, **H.net_inputs(g), **H.NET_CFG[name])
    for nm, t in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), out):
        assert H.rel_err(t, g['fwd.' + nm]) <= 1e-6, nm


def test_gradients(gold):
    name, g = gold
    P = H.params(g, requires_grad=True)
    out = R.net_forward(P, **H.net_inputs(g), **H.NET_CFG[name], training=True)
    loss = R.training_loss(out, th.tensor(g['train.labels']))
    assert abs(float(loss.detach()) - float(g['loss'])) <= 1e-6
    loss.backward()
    for k in g:
        if not k.startswith('grad.'):
            continue
        p = P[k[5:]]
        if not bool(g['hasgrad.' + k[5:]]):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert H.rel_err(p.grad, g[k]) <= 1e-5, k


def test_augmentation_masks(gold):
    _, g = gold
    eg = H.enc_graph(g)
    knn = [H.knn_coo(g, key) for key in H.KNN_KEYS]
    feats = (th.tensor(g['feat.drug']), th.tensor(g['feat.disease']),
             th.tensor(g['mat.drug'], dtype=th.float32), th.tensor(g['mat.disease'], dtype=th.float32))
    gen = th.Generator().manual_seed(123)
    ag, coos, noisy = R.augment_default(eg, knn, feats, 0.1, 0.05, 0.05, gen)
    for i, et in enumerate(R.canonical_etype_order(eg['edges'])):
        np.testing.assert_array_equal(np.stack(ag['edges'][et]), g[f'aug.enc.{et}'])
        keep = R.edge_dropout_keep(g[f'aug.perm.{i}'], 0.1)
        np.testing.assert_array_equal(eg['edges'][et][0][keep], g[f'aug.enc.{et}'][0])
    for key, coo in zip(H.KNN_KEYS, coos):
        np.testing.assert_array_equal(np.stack(coo[:2]), g[f'aug.knn.{key}.indices'])
        np.testing.assert_array_equal(coo[2], g[f'aug.knn.{key}.values'])
    for fk, t in zip(('drug_feat', 'disease_feat', 'drug_sim_feat', 'disease_sim_feat'), noisy):
        np.testing.assert_array_equal(t.numpy(), g['aug.' + fk])


def test_num_keep_double_arithmetic():
    assert R.dropout_num_keep(464896, 0.1) == 418406          # SURVEY.md 7 hard part 3
    assert R.dropout_num_keep(1, 0.9) == 1
    assert R.dropout_num_keep(10, 0.1) == 9


def test_three_training_iterations(gold):
    """Replays the reference's train() (3 iterations, default dropout + augmentation) from the saved
    post-init RNG state: same augmentation masks, same dropout draws, same Adam step."""
    name, g = gold
    P = {}
    for k, v in g.items():
        if k.startswith('train.sd0.'):
            P[k[10:]] = th.tensor(v)
    for k in list(P):
        if '.ifc.' in k:
            P[k] = P[k.replace('.ifc.', '.ufc.')]
    leaves = list({id(v): v for v in P.values()}.values())
    for v in leaves:
        v.requires_grad_(True)
    opt = th.optim.Adam(leaves, lr=0.002, weight_decay=1e-5)
    inp = H.net_inputs(g)
    knn = [inp['drug_graph'], inp['dis_graph'], inp['drug_feature_graph'], inp['dis_feature_graph']]
    feats = (inp['drug_feat'], inp['dis_feat'], inp['drug_sim_feat'], inp['dis_sim_feat'])
    cfg = dict(H.NET_CFG[name], dropout=0.3, attention_dropout=0.1)
    saved = th.get_rng_state()
    try:
        th.set_rng_state(th.tensor(g['train.rng0']))
        for it in range(1, 4):
            loss = R.train_iteration(P, opt, it, inp['enc_graph'], inp['dec_pairs'],
                                     th.tensor(g['train.labels']), knn, feats, cfg)
    finally:
        th.set_rng_state(saved)
    want = float(str(g['train.log'][0]).split('Loss=')[1].split(',')[0])
    assert abs(loss - want) <= 6e-5, (loss, want)
    tinp = H.net_inputs(g, 'test')
    auroc, aupr = R.evaluate_auc(P, tinp['enc_graph'], tinp['dec_pairs'], g['test.labels'], knn, feats, cfg)
    assert abs(auroc - float(g['train.auroc'])) <= 1e-3 and abs(aupr - float(g['train.aupr'])) <= 1e-3
