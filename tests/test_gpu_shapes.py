"""GPU parity at the BASELINE.json dataset shapes (configs 1-3): lrssl 763 x 681 (E = 467 641 training pairs), Gdataset
593 x 313, Cdataset 663 x 409 with the CLI-default model (768-dim embeddings -> 341-wide messages padded to 344,
nhid 768 / 128, 3 layers).

The product path is exercised exactly as a user runs it: a `.mat` file of the reference's schema -> `DrugDataLoader`
(device graph builders) -> `Net` -> loss -> backward -> `evaluate`. Comparands:
  * digests of the UNMODIFIED reference at the same shape (tests/golden/shape_*.npz, make_golden_shapes.py);
  * the CPU oracle (oracle/restate.py) run live in float64 = the exact value of the same expression.
Integer / index work and the fp32 normalisers and adjacency values are bit-exact (SHA-256); activations <= 1e-5
norm-wise; gradients within tests/shapes.py:budget (1e-5 against the float64 value, or 2 x the reference's own fp32
deviation from it where that is larger); AUROC / AUPR within 1e-3.
"""
import os
import tempfile

import numpy as np
import pytest
import scipy.io as sio
import torch as th

from oracle import restate as R
from tests import helpers as H
from tests import shapes as S

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
KNN_KEYS = ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph')


@pytest.fixture(scope='module', params=list(S.DATASETS))
def shape(request):
    from dreamgnn_b200 import _lib
    from dreamgnn_b200.data_loader import DrugDataLoader
    _lib.load()
    name = request.param
    ds = S.dataset(name)
    root = tempfile.mkdtemp(prefix='dg_shape_')
    d = os.path.join(root, 'raw_data', 'drug_data', 'lrssl')
    os.makedirs(d)
    sio.savemat(os.path.join(d, 'lrssl.mat'), ds['arrays'])
    old = os.getcwd()
    os.chdir(root)
    try:
        loader = DrugDataLoader('lrssl', 'cuda:0', symm=True, k=S.K_NEIGHBOR)
    finally:
        os.chdir(old)
    enc, knn = S.oracle_graphs(ds)
    return name, ds, loader, enc, knn, S.load_shape_golden(name)


def _net(ds, sd, dropout=0.0, attention_dropout=0.0):
    from dreamgnn_b200.model import Net
    net = Net(S.net_args(ds, dropout, attention_dropout))
    net.load_state_dict(sd)
    return net.to('cuda:0')


def _call(loader, split='train'):
    dev = 'cuda:0'
    enc, dec, labels = loader.data_cv[0][split]
    gr = loader.cv_specific_graphs[0]
    dsim = th.as_tensor(loader.drug_sim_features, dtype=th.float32).to(dev)
    ssim = th.as_tensor(loader.disease_sim_features, dtype=th.float32).to(dev)
    return (enc.int().to(dev), dec.int().to(dev), gr['drug_graph'], dsim, loader.drug_feature, gr['disease_graph'], ssim,
            loader.disease_feature, gr['drug_feature_graph'], gr['disease_feature_graph']), labels.to(dev)


def test_loader_bit_exact(shape):
    """Fold pairs, labels, encoder edge lists, ci / cj and the four kNN graphs equal the reference's (SHA-256)."""
    _, ds, loader, _, _, g = shape
    for split in ('train', 'test'):
        enc, dec, labels = loader.data_cv[0][split]
        s, d = dec.edges()
        assert S.sha(np.stack([s.cpu().numpy(), d.cpu().numpy()]).astype(np.int64)) == str(g[f'hash.{split}.pairs'])
        assert S.sha(labels.cpu().numpy().astype(np.float32)) == str(g[f'hash.{split}.labels'])
        for c in enc.canonical_etypes:
            es, ed = enc.edges(etype=c)
            assert S.sha(np.stack([es.cpu().numpy(), ed.cpu().numpy()]).astype(np.int64)) == str(g[f'hash.{split}.enc.{c[1]}'])
        for nt in ('drug', 'disease'):
            assert S.sha(enc.nodes[nt].data['ci'].cpu().numpy()) == str(g[f'hash.{split}.ci.{nt}'])
            assert S.sha(enc.nodes[nt].data['cj'].cpu().numpy()) == str(g[f'hash.{split}.cj.{nt}'])
    for gk in KNN_KEYS:
        t = loader.cv_specific_graphs[0][gk]
        assert t._values().numel() == int(g[f'meta.knn.{gk}.nnz']), gk
        assert S.sha(t._indices().cpu().numpy().astype(np.int64)) == str(g[f'hash.knn.{gk}.indices']), gk
        assert S.sha(t._values().cpu().numpy().astype(np.float32)) == str(g[f'hash.knn.{gk}.values']), gk
    # features: the device's fp32 row normalisation may differ from the CPU's in the last bit
    np.testing.assert_allclose(loader.drug_feature.cpu().numpy(), ds['drug_feat'].numpy(), rtol=2e-6, atol=1e-9)


def test_net_forward_eval(shape):
    """Net.forward, eval mode, both splits: vs the reference digest and vs the float64 oracle, <= 1e-5 norm-wise."""
    _, ds, loader, enc, knn, g = shape
    sd = S.init_state_dict(ds)
    assert S.state_dict_hash(sd) == str(g['hash.sd'])
    net = _net(ds, sd).eval()
    names = ('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out')
    with th.no_grad():
        out = net(*_call(loader)[0])
    for nm, t in zip(names, out):
        e_s, e_n = S.digest_errors('fwd.' + nm, t, float(g[f'fwd.{nm}.norm']), g[f'fwd.{nm}.samples'])
        assert e_s <= FP32_TOL and e_n <= FP32_TOL, (nm, e_s, e_n)
    for split in ('train', 'test'):
        with th.no_grad():
            out = net(*_call(loader, split)[0])
            _, ref = S.oracle_forward(ds, enc, knn, sd, split=split, dtype=th.float64)
        for nm, a, b in zip(names, out, ref):
            assert H.rel_err(a.cpu(), b) <= FP32_TOL, (split, nm)


def test_net_gradients(shape):
    """Training loss (train.py:286-294, dropout p = 0) and every parameter gradient."""
    from dreamgnn_b200.utils import common_loss
    _, ds, loader, enc, knn, g = shape
    sd = S.init_state_dict(ds)
    net = _net(ds, sd).train()
    call, labels = _call(loader)
    with S.capture_relu_masks() as masks:
        out = net(*call)
    loss = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), labels) + 0.001 * (
        common_loss(out[1], out[2]) + common_loss(out[3], out[4]))
    loss.backward()
    loss_gpu = float(loss.detach())
    # the exact (float64) value of the same expression, differentiated through the ReLU masks the forward under test
    # produced (tests/shapes.py:capture_relu_masks); and, unmatched, for the count of entries whose mask differs at all
    _, loss64, g64 = S.oracle_loss_and_grads(ds, enc, knn, sd, th.float64, relu_masks=masks)
    assert abs(loss_gpu - float(g['loss'])) <= 2e-6 and abs(loss_gpu - loss64) <= 2e-6, (loss_gpu, float(g['loss']), loss64)
    grads = dict(net.named_parameters())
    names = [key[8:] for key in g if key.startswith('hasgrad.')]
    gold = {k: (float(g[f'grad.{k}.norm']), g[f'grad.{k}.samples']) for k in names if bool(g['hasgrad.' + k])}
    _, _, g64_own = S.oracle_loss_and_grads(ds, enc, knn, sd, th.float64)          # the oracle's own masks
    ref_vs_exact = {k: S.digest_errors('grad.' + k, g64_own[k].float(), *gold[k])[0] for k in gold}
    numel = {k: grads[k].numel() for k in gold}
    budget = S.class_budgets(ref_vs_exact, numel)
    rows, failed = [], []
    for k in names:
        p = grads[k]
        if k not in gold:
            zero = p.grad is None or float(p.grad.abs().max()) == 0.0
            assert zero, k
            continue
        e_exact = H.rel_err(p.grad.cpu(), g64[k])
        e_ref = max(S.digest_errors('grad.' + k, p.grad, *gold[k]))
        rows.append('%-34s vs float64 %.2e (budget %.1e)  vs reference digest %.2e  reference vs float64 %.2e'
                    % (k, e_exact, budget[k], e_ref, ref_vs_exact[k]))
        if e_exact > budget[k]:
            failed.append(rows[-1])
    print('\n'.join(rows))
    assert not failed, failed


def test_evaluate_auc(shape):
    """`evaluate` (evaluation.py:4-74) at the reference's weights vs the reference's own AUROC / AUPR; then after 20
    training iterations of the product path (augmentation + dropout on), vs the oracle at the trained weights."""
    import argparse
    from dreamgnn_b200.evaluation import evaluate
    from dreamgnn_b200.train import TrainState, aug_params_from_args, train_iteration
    _, ds, loader, enc, knn, g = shape
    sd = S.init_state_dict(ds)
    net = _net(ds, sd, dropout=0.3, attention_dropout=0.1)
    args = argparse.Namespace(device='cuda:0')
    gr = loader.cv_specific_graphs[0]
    call, labels = _call(loader)
    dsim, ssim = call[3], call[6]

    def ev(split):
        return evaluate(args, net, {'test': loader.data_cv[0][split]}, gr['drug_graph'], loader.drug_feature, dsim,
                        gr['disease_graph'], loader.disease_feature, ssim, gr['drug_feature_graph'],
                        gr['disease_feature_graph'])
    for split in ('train', 'test'):
        a, p = ev(split)
        assert abs(a - float(g[f'eval.{split}.auroc'])) <= 1e-3 and abs(p - float(g[f'eval.{split}.aupr'])) <= 1e-3, split
    state = TrainState(call[0], call[1], labels, call[2], call[5], call[8], call[9], call[4], call[7], dsim, ssim)
    opt = th.optim.Adam(net.parameters(), lr=0.002, weight_decay=1e-5)
    th.manual_seed(5)
    for _ in range(20):
        loss = train_iteration(net, opt, state, th.nn.BCEWithLogitsLoss(), ['edge_dropout', 'feature_noise'],
                               aug_params_from_args(argparse.Namespace()))
    assert np.isfinite(float(loss))
    trained = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    feats = (ds['drug_feat'], ds['dis_feat'], ds['drug_sim'], ds['dis_sim'])
    pairs, lab = ds['split']['test']
    ra, rp = R.evaluate_auc(S.oracle_params(trained), enc['test'], pairs, lab, knn, feats, dict(layers=3))
    a, p = ev('test')
    assert abs(a - ra) <= 1e-3 and abs(p - rp) <= 1e-3, (a, ra, p, rp)
