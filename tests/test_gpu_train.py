"""GPU: the DrugDataLoader mirror and train() end to end on a synthetic `.mat` of the reference's schema,
against the golden vectors the reference produced from the same file."""
import argparse
import os
import tempfile

import numpy as np
import pytest
import scipy.io as sio
import torch as th

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def loaded():
    from dreamgnn_b200 import _lib
    from dreamgnn_b200.data_loader import DrugDataLoader
    _lib.load()
    g = H.load_golden('tinyA')
    root = tempfile.mkdtemp(prefix='dg_mat_')
    d = os.path.join(root, 'raw_data', 'drug_data', 'lrssl')
    os.makedirs(d)
    names = np.empty((60, 1), dtype=object)
    for i in range(60):
        names[i, 0] = np.array(['DB%05d' % i])
    sio.savemat(os.path.join(d, 'lrssl.mat'), {'didr': g['mat.didr'], 'drug': g['mat.drug'], 'disease': g['mat.disease'],
                                               'drug_embed': g['mat.drug_embed'], 'disease_embed': g['mat.disease_embed'],
                                               'Wrname': names})
    old = os.getcwd()
    os.chdir(root)
    try:
        ds = DrugDataLoader('lrssl', 'cuda:0', symm=True, k=int(g['k']))
    finally:
        os.chdir(old)
    return root, g, ds


def test_loader_matches_reference_fold0(loaded):
    _, g, ds = loaded
    assert ds.num_drug == 60 and ds.num_disease == 45 and len(ds.data_cv) == 10
    for split in ('train', 'test'):
        enc, dec, labels = ds.data_cv[0][split]
        s, d = dec.edges()
        np.testing.assert_array_equal(np.stack([s.cpu().numpy(), d.cpu().numpy()]), g[f'{split}.pairs'])
        np.testing.assert_array_equal(labels.numpy(), g[f'{split}.labels'])
        for nt in ('drug', 'disease'):
            np.testing.assert_array_equal(enc.nodes[nt].data['ci'].cpu().numpy(), g[f'{split}.ci.{nt}'])
    np.testing.assert_allclose(ds.drug_feature.cpu().numpy(), g['feat.drug'], rtol=1e-6, atol=1e-8)   # fp32 normalise: 1 ulp
    for key in H.KNN_KEYS:
        t = ds.cv_specific_graphs[3][key]
        gr, gc, gv = H.canon_coo(g[f'knn.{key}.indices'][0], g[f'knn.{key}.indices'][1], g[f'knn.{key}.values'])
        np.testing.assert_array_equal(t._indices().cpu().numpy(), np.stack([gr, gc]))
        np.testing.assert_array_equal(t._values().cpu().numpy(), gv)


def test_train_runs_and_evaluates(loaded):
    """train() with the reference's flags: finite loss, AUROC/AUPR in range, result files written; and with
    all randomness off the evaluation equals the oracle's on the same weights (<= 1e-3, north_star)."""
    from dreamgnn_b200.train import build_parser, train
    from dreamgnn_b200.evaluation import evaluate
    from dreamgnn_b200.model import Net
    from oracle import restate as R
    root, g, ds = loaded
    args = build_parser().parse_args(['--data_name', 'lrssl', '--train_max_iter', '7', '--train_valid_interval', '3',
                                      '--gcn_agg_units', '105', '--gcn_out_units', '16', '--nhid1', '40', '--nhid2', '16',
                                      '--num_neighbor', '4'])
    args.device, args.save_dir, args.save_id = 'cuda:0', root, 1
    th.manual_seed(77)
    auroc, aupr = train(args, ds, 0)
    assert 0.0 <= auroc <= 1.0 and 0.0 <= aupr <= 1.0
    assert os.path.isfile(os.path.join(root, 'test_metric1.csv')) and os.path.isfile(os.path.join(root, 'best_metric1.csv'))
    log = open(os.path.join(root, 'test_metric1.csv')).read().splitlines()
    assert len(log) == 3 and all(np.isfinite(float(r.split(',')[1])) for r in log[1:])
    # evaluation parity on fixed weights
    margs = argparse.Namespace(**vars(args))
    margs.rating_vals, margs.src_in_units, margs.dst_in_units = [0, 1], 48, 48
    margs.fdim_drug, margs.fdim_disease = 60, 45
    net = Net(margs)
    net.load_state_dict({k[3:]: th.tensor(v) for k, v in g.items() if k.startswith('sd.')})
    net = net.to('cuda:0')
    gr = ds.cv_specific_graphs[0]
    dsim = th.as_tensor(ds.drug_sim_features, dtype=th.float32).cuda()
    ssim = th.as_tensor(ds.disease_sim_features, dtype=th.float32).cuda()
    a, p = evaluate(args, net, {'test': ds.data_cv[0]['test']}, gr['drug_graph'], ds.drug_feature, dsim,
                    gr['disease_graph'], ds.disease_feature, ssim, gr['drug_feature_graph'], gr['disease_feature_graph'])
    inp = H.net_inputs(g, 'test')
    knn = [inp['drug_graph'], inp['dis_graph'], inp['drug_feature_graph'], inp['dis_feature_graph']]
    feats = (inp['drug_feat'], inp['dis_feat'], inp['drug_sim_feat'], inp['dis_sim_feat'])
    ra, rp = R.evaluate_auc(H.params(g), inp['enc_graph'], inp['dec_pairs'], g['test.labels'], knn, feats, dict(layers=3))
    assert abs(a - ra) <= 1e-3 and abs(p - rp) <= 1e-3
    # the device metric code equals the reference's sklearn path on the same logits
    os.environ['DG_EVAL'] = 'sklearn'
    try:
        a_sk, p_sk, (y_score, y_true) = evaluate(args, net, {'test': ds.data_cv[0]['test']}, gr['drug_graph'], ds.drug_feature,
                                                 dsim, gr['disease_graph'], ds.disease_feature, ssim, gr['drug_feature_graph'],
                                                 gr['disease_feature_graph'], return_predictions=True)
    finally:
        del os.environ['DG_EVAL']
    assert abs(a - a_sk) <= 1e-12 and abs(p - p_sk) <= 1e-12 and y_score.shape == y_true.shape


def test_top_novel_predictions(loaded):
    """predict.get_top_novel_predictions (one encoder pass + one fused decoder launch over all unknown pairs) against the
    reference's algorithm: the oracle's Net.forward on the same pairs, sigmoid, sort (train.py:26-151)."""
    from oracle import restate as R
    from dreamgnn_b200.model import Net
    from dreamgnn_b200.predict import get_top_novel_predictions, novel_pair_scores
    root, g, ds = loaded
    margs = argparse.Namespace(model_activation='leaky', gcn_agg_accum='sum', share_param=True, device=None, dropout=0.3,
                               attention_dropout=0.1, rating_vals=[0, 1], layers=3, gcn_agg_units=105, gcn_out_units=16,
                               nhid1=40, nhid2=16, src_in_units=48, dst_in_units=48, fdim_drug=60, fdim_disease=45)
    net = Net(margs)
    net.load_state_dict({k[3:]: th.tensor(v) for k, v in g.items() if k.startswith('sd.')})
    net = net.to('cuda:0')
    args = argparse.Namespace(device='cuda:0', save_dir=root)
    drug_id, dis_id, score = novel_pair_scores(args, net, ds, 0)
    truth = np.asarray(ds.association_matrix)
    want = np.argwhere(truth == 0)
    np.testing.assert_array_equal(np.stack([drug_id.cpu().numpy(), dis_id.cpu().numpy()], 1), want)      # order of the nested loops
    inp = H.net_inputs(g, 'train')
    inp['dec_pairs'] = (want[:, 0], want[:, 1])
    ref = th.sigmoid(R.net_forward(H.params(g), **inp, layers=3)[0].reshape(-1)).detach()
    assert H.rel_err(score.cpu(), ref) <= 1e-5
    df = get_top_novel_predictions(args, net, ds, 0, top_k=50)
    assert list(df.columns) == ['drug_id', 'disease_id', 'score', 'drug_name'] and len(df) == 50
    assert os.path.isfile(os.path.join(root, 'top50_novel_predictions_fold1.csv'))
    order = th.argsort(ref, descending=True)[:50]
    assert set(zip(df.drug_id, df.disease_id)) == set(map(tuple, want[order.numpy()]))
    assert np.all(np.diff(df.score.values) <= 0) and df.drug_name[0] == 'DB%05d' % df.drug_id[0]
    np.testing.assert_allclose(df.score.values, ref[order].numpy(), rtol=1e-5)


def test_cuda_graph_iteration(loaded):
    """One captured training iteration replays correctly: finite, decreasing loss, a fresh dropout / edge-dropout
    draw on every replay, and the same trajectory as eager execution (statistically: same loss level)."""
    import argparse as ap
    from dreamgnn_b200 import synthetic
    from dreamgnn_b200.graphed import GraphedIteration
    from dreamgnn_b200.model import Net
    from dreamgnn_b200.train import train_iteration, aug_params_from_args
    from dreamgnn_b200.utils import common_loss
    dev = th.device('cuda:0')
    spec = dict(kind='dense', n_drug=90, n_dis=70, n_pos=400, f_drug=48, f_dis=48, k=4)
    w = synthetic.make_workload(spec, dev, seed=5)
    state = synthetic.train_state(w, dev)
    margs = synthetic.model_args(w, gcn_agg_units=96, gcn_out_units=16, nhid1=40, nhid2=16)
    curves = {}
    for graphed in (False, True):
        th.manual_seed(9)
        model = Net(margs).to(dev)
        opt = th.optim.Adam(model.parameters(), lr=0.002, weight_decay=1e-5, capturable=graphed)
        if graphed:
            step = GraphedIteration(model, opt, state)
        else:
            fn = th.nn.BCEWithLogitsLoss()
            p = aug_params_from_args(ap.Namespace())
            step = lambda: train_iteration(model, opt, state, fn, ['edge_dropout', 'feature_noise'], p, 0.001, 1.0, common_loss)
            for _ in range(3):
                step()
        curves[graphed] = [float(step().detach()) for _ in range(40)]
    g = curves[True]
    assert all(np.isfinite(g)) and np.mean(g[-10:]) < np.mean(g[:10])
    assert len({round(x, 6) for x in g[:5]}) == 5                     # replays are not identical: fresh random draws
    assert abs(np.mean(g[-10:]) - np.mean(curves[False][-10:])) < 0.05
    with pytest.raises(ValueError, match='capturable'):
        GraphedIteration(model, th.optim.Adam(model.parameters()), state)


def test_cuda_graph_pipelined_augmentation(loaded):
    """The augmentation of iteration i+1 is recorded as a parallel branch of iteration i's graph: every replay
    trains on the result the previous replay staged (bit-identical buffers) and stages a fresh draw; the serial
    capture (pipeline_aug=False) and the pipelined one reach the same loss level."""
    from dreamgnn_b200 import synthetic
    from dreamgnn_b200.graphed import GraphedIteration, _entry_tensors
    from dreamgnn_b200.model import Net
    dev = th.device('cuda:0')
    spec = dict(kind='dense', n_drug=90, n_dis=70, n_pos=400, f_drug=48, f_dis=48, k=4)
    w = synthetic.make_workload(spec, dev, seed=5)
    state = synthetic.train_state(w, dev)
    margs = synthetic.model_args(w, gcn_agg_units=96, gcn_out_units=16, nhid1=40, nhid2=16)
    curves = {}
    for pipelined in (False, True):
        th.manual_seed(9)
        model = Net(margs).to(dev)
        opt = th.optim.Adam(model.parameters(), lr=0.002, weight_decay=1e-5, capturable=True)
        step = GraphedIteration(model, opt, state, pipeline_aug=pipelined)
        assert (step.staged is not None) == pipelined
        if pipelined:
            st = step.staged
            assert set(st.keys) == {'enc_graph', 'drug_graph', 'disease_graph', 'drug_feature_graph',
                                    'disease_feature_graph', 'drug_feat', 'disease_feat', 'drug_sim_feat',
                                    'disease_sim_feat'} and st.nbytes() > 0
            for _ in range(3):
                before = {k: [t.clone() for t in _entry_tensors(st.tree[k])] for k in st.keys}
                step()
                th.cuda.synchronize()
                changed = 0
                for k in st.keys:
                    live, staged = _entry_tensors(step._live[k]), _entry_tensors(st.tree[k])
                    assert len(live) == len(before[k]) == len(staged)
                    for a, b, c in zip(live, before[k], staged):
                        assert th.equal(a, b)                        # trained on what the previous replay staged
                        changed += int(not th.equal(b, c))
                assert changed >= len(st.keys)                        # and staged a fresh draw for the next one
            # the staged structures are valid CSRs of the right size (same invariants the serial path gives)
            blk = st.tree['enc_graph'].block('disease')
            assert int(blk.csr.indptr[-1]) == blk.csr.nnz == blk.csr.transpose().nnz
        curves[pipelined] = [float(step().detach()) for _ in range(40)]
    for c in curves.values():
        assert all(np.isfinite(c)) and np.mean(c[-10:]) < np.mean(c[:10])
    assert abs(np.mean(curves[True][-10:]) - np.mean(curves[False][-10:])) < 0.05
    assert GraphedIteration(model, opt, state).staged is not None          # default: by size -> on for small shapes
