"""CPU, build container only: oracle/restate.py against the LIVE unmodified reference (run through
oracle/dgl) on a fresh random dataset that is not one of the frozen goldens."""
import argparse
import tempfile

import numpy as np
import pytest
import torch as th

from oracle import ref_runner as rr
from oracle import restate as R
from tests import helpers as H

pytestmark = pytest.mark.skipif(not rr.reference_available(), reason='/root/reference not present')


@pytest.fixture(scope='module')
def live():
    root = tempfile.mkdtemp(prefix='dg_live_')
    rr.write_synthetic_mat(root, 'lrssl', n_drug=41, n_dis=29, n_pos=130, embed_dim=20, sim_rank=8, seed=7)
    mods, ds = rr.load_reference_dataset(root, 'lrssl', k=5)
    return root, mods, ds


def test_graph_construction_live(live):
    _, _, ds = live
    enc, dec, labels = ds.data_cv[3]['train']
    s, d = dec.edges()
    eg = R.enc_graph_from_pairs((s.numpy(), d.numpy()), labels.numpy(), 41, 29)
    for c in enc.canonical_etypes:
        es, ed = enc.edges(etype=c)
        np.testing.assert_array_equal(np.stack(eg['edges'][c[1]]), np.stack([es.numpy(), ed.numpy()]))
    for nt in ('drug', 'disease'):
        np.testing.assert_array_equal(eg['ci'][nt], enc.nodes[nt].data['ci'].numpy())
        np.testing.assert_array_equal(eg['cj'][nt], enc.nodes[nt].data['cj'].numpy())
    for key, sim in (('drug_graph', ds.drug_sim_features), ('disease_graph', ds.disease_sim_features),
                     ('drug_feature_graph', R.feature_cosine_similarity(ds.drug_embed)),
                     ('disease_feature_graph', R.feature_cosine_similarity(ds.disease_embed))):
        t = ds.cv_specific_graphs[3][key]
        row, col, val = R.similarity_knn_graph(sim, 5)
        gr, gc, gv = H.canon_coo(t._indices()[0].numpy(), t._indices()[1].numpy(), t._values().numpy())
        np.testing.assert_array_equal(row, gr)
        np.testing.assert_array_equal(col, gc)
        np.testing.assert_array_equal(val, gv)


def test_forward_backward_live(live):
    root, mods, ds = live
    args = argparse.Namespace(
        layers=2, model_activation='leaky', gcn_agg_units=60, gcn_out_units=12, dropout=0.0,
        gcn_agg_accum='sum', share_param=True, device='cpu', nhid1=24, nhid2=12, attention_dropout=0.0,
        src_in_units=20, dst_in_units=20, fdim_drug=41, fdim_disease=29, rating_vals=[0, 1])
    th.manual_seed(5)
    net = mods['model'].Net(args)
    enc, dec, labels = ds.data_cv[3]['train']
    graphs = ds.cv_specific_graphs[3]
    dsim, ssim = th.FloatTensor(ds.drug_sim_features), th.FloatTensor(ds.disease_sim_features)
    res = net(enc.int(), dec.int(), graphs['drug_graph'], dsim, ds.drug_feature, graphs['disease_graph'],
              ssim, ds.disease_feature, graphs['drug_feature_graph'], graphs['disease_feature_graph'])
    loss = th.nn.BCEWithLogitsLoss()(res[0].squeeze(-1), labels) + 0.001 * (
        mods['utils'].common_loss(res[1], res[2]) + mods['utils'].common_loss(res[3], res[4]))
    loss.backward()

    P = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for k in list(P):
        if '.ifc.' in k:
            P[k] = P[k.replace('.ifc.', '.ufc.')]
    for v in {id(v): v for v in P.values()}.values():
        v.requires_grad_(True)
    s, d = dec.edges()
    eg = R.enc_graph_from_pairs((s.numpy(), d.numpy()), labels.numpy(), 41, 29)

    def coo(t, n):
        return t._indices()[0].numpy(), t._indices()[1].numpy(), t._values().numpy(), n
    out = R.net_forward(P, eg, (s.numpy(), d.numpy()), coo(graphs['drug_graph'], 41), dsim, ds.drug_feature,
                        coo(graphs['disease_graph'], 29), ssim, ds.disease_feature,
                        coo(graphs['drug_feature_graph'], 41), coo(graphs['disease_feature_graph'], 29),
                        layers=2, training=True)
    for a, b in zip(out, res):
        assert H.rel_err(a.detach(), b.detach()) <= 1e-6
    R.training_loss(out, labels).backward()
    for k, p in net.named_parameters():
        if p.grad is not None:
            assert H.rel_err(P[k].grad, p.grad) <= 1e-5, k


def test_knn_option_set_matches_reference():
    """a14 + the symm=False loader branch: oracle restatements vs the reference's own functions (utils.knn_graph and
    augmentation.augmented_knn_graph are dead code there, but part of the kNN builder's option set)."""
    import scipy.sparse as sp
    mods = rr.import_reference()
    rng = np.random.default_rng(5)
    x = rng.standard_normal((40, 6))
    sim = x @ x.T
    for k in (0, 3, 7, 60):
        want = sp.coo_matrix(mods['utils'].knn_graph(sim, k)).tocsr()
        want.sort_indices()
        want = want.tocoo()
        row, col, val = R.knn_graph_binary(sim, k)
        np.testing.assert_array_equal(row, want.row)
        np.testing.assert_array_equal(col, want.col)
        np.testing.assert_array_equal(val, want.data.astype(np.float32))
    # augmented version: replay numpy's global generator for the injected draws
    base_row, base_col, base_val = R.knn_graph_binary(sim, 4)
    np.random.seed(3)
    want = sp.coo_matrix(mods['augmentation'].augmented_knn_graph(sim, 4, dropout_rate=0.2, add_noise=True, noise_scale=0.1)).tocsr()
    want.sort_indices()
    want = want.tocoo()
    np.random.seed(3)
    noise = np.random.normal(0, 1.0, len(base_val))               # normal(0, s, n) == s * normal(0, 1, n) in numpy's generator
    keep = np.random.choice(len(base_val), max(1, int(len(base_val) * 0.8)), replace=False)
    row, col, val = R.augmented_knn_graph(sim, 4, keep=keep, noise=noise, noise_scale=0.1)
    np.testing.assert_array_equal(row, want.row)
    np.testing.assert_array_equal(col, want.col)
    np.testing.assert_allclose(val, want.data, rtol=1e-12, atol=0)
    # loader with symm=False
    loader = object.__new__(mods['data_loader'].DrugDataLoader)
    loader._symm = False
    t = loader._create_similarity_graph(sim, 4)
    gr, gc, gv = H.canon_coo(t._indices()[0].numpy(), t._indices()[1].numpy(), t._values().numpy())
    row, col, val = R.similarity_knn_graph(sim, 4, symm=False)
    np.testing.assert_array_equal(row, gr)
    np.testing.assert_array_equal(col, gc)
    np.testing.assert_array_equal(val, gv)
