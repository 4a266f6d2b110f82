"""The drop-in boundary exercised by the REFERENCE'S OWN modules: `dreamgnn_b200.dgl_compat` serves `import dgl`, and the
unmodified reference files (the source tree in the build container, the archive staged by oracle/stage_ref.py on the GPU
box) run on top of this repo's graph handle.

CPU part (no kernels): the reference's `DrugDataLoader` and `augment_graph_data` drive the graph handle's structural
API. GPU part: the reference's own `layers.py` + `model.py` (every `update_all(copy_u, sum)` lands in `dg_spmm_csr_f32`),
then the reference's `model.py` over `dreamgnn_b200.layers` (fused path), each against the goldens the reference produced
on its DGL stand-in; finally the reference's own `train()` loop end to end on the device.
"""
import argparse
import contextlib
import io
import os
import tempfile

import numpy as np
import pytest
import torch as th

from oracle import ref_runner as rr
from tests import helpers as H

pytestmark = pytest.mark.skipif(not rr.reference_available(), reason='no reference tree and no staged archive (oracle/_ref)')
FP32_TOL = 1e-5
DATA = dict(n_drug=60, n_dis=45, n_pos=220, embed_dim=48, sim_rank=12, seed=0)          # = golden case tinyA
NET = dict(layers=3, gcn_agg_units=105, gcn_out_units=16, nhid1=40, nhid2=16)


def _load(device, replace_layers=False):
    from dreamgnn_b200 import dgl_compat
    replace = dgl_compat.build_modules()
    if replace_layers:
        from dreamgnn_b200 import layers
        replace['layers'] = layers
    mods = rr.import_reference(replace=replace)
    root = tempfile.mkdtemp(prefix='dg_dropin_')
    rr.write_synthetic_mat(root, 'lrssl', **DATA)
    with rr.chdir(root), contextlib.redirect_stdout(io.StringIO()):
        ds = mods['data_loader'].DrugDataLoader('lrssl', th.device(device), symm=True, k=4)
    return mods, ds, root


def _args(ds, root, device, dropout=0.0, attention_dropout=0.0):
    a = argparse.Namespace(
        model_activation='leaky', gcn_agg_accum='sum', share_param=True, device=device, dropout=dropout,
        attention_dropout=attention_dropout, train_max_iter=7, train_valid_interval=3, train_lr=0.002, weight_decay=1e-5,
        beta=0.001, train_grad_clip=1.0, save_dir=root, save_id=1, save_model=False, generate_top_predictions=False,
        label_smoothing=0.0, aug_methods=['edge_dropout', 'feature_noise'], edge_dropout_rate=0.1, feature_noise_scale=0.05,
        graph_noise_scale=0.03, add_edge_rate=0.03, feature_mask_rate=0.1, mixup_alpha=0.2, **NET)
    a.src_in_units, a.dst_in_units = ds.drug_feature_shape[1], ds.disease_feature_shape[1]
    a.fdim_drug, a.fdim_disease = ds.drug_feature_shape[0], ds.disease_feature_shape[0]
    a.rating_vals = ds.cv_data_dict[0][2]
    return a


def test_reference_loader_and_augmentation_on_graph_handle():
    from dreamgnn_b200.graph import HeteroGraph
    mods, ds, _ = _load('cpu')
    assert mods['layers'].__name__ == 'layers' and 'dreamgnn_b200' in (mods['layers'].dgl.__doc__ or '')
    g = H.load_golden('tinyA')
    for split in ('train', 'test'):
        enc, dec, labels = ds.data_cv[0][split]
        assert isinstance(enc, HeteroGraph) and isinstance(dec, HeteroGraph)
        s, d = dec.edges()
        np.testing.assert_array_equal(np.stack([s.numpy(), d.numpy()]), g[f'{split}.pairs'])
        for c in enc.canonical_etypes:
            es, ed = enc.edges(etype=c)
            np.testing.assert_array_equal(np.stack([es.numpy(), ed.numpy()]), g[f'{split}.enc.{c[1]}'])
        for nt in ('drug', 'disease'):
            np.testing.assert_array_equal(enc.nodes[nt].data['ci'].numpy(), g[f'{split}.ci.{nt}'])
            np.testing.assert_array_equal(enc.nodes[nt].data['cj'].numpy(), g[f'{split}.cj.{nt}'])
    enc = ds.data_cv[0]['train'][0].int()
    th.manual_seed(123)
    aug = mods['augmentation'].augment_graph_data({'enc_graph': enc}, ['edge_dropout'], {'edge_dropout_rate': 0.1})
    for c in aug['enc_graph'].canonical_etypes:                # same seed, same draws -> the golden kept edges
        es, ed = aug['enc_graph'].edges(etype=c)
        np.testing.assert_array_equal(np.stack([es.numpy(), ed.numpy()]), g[f'aug.enc.{c[1]}'])


@pytest.mark.gpu
@pytest.mark.parametrize('replace_layers', [False, True], ids=['reference-layers', 'b200-layers'])
def test_reference_model_on_b200_graph(replace_layers):
    """The reference's own `model.Net` (+ its own `layers.py`, or this repo's) on the device graph handle: outputs and every
    parameter gradient equal the goldens of the reference-on-stand-in run."""
    from dreamgnn_b200 import _lib
    _lib.load()
    dev = th.device('cuda:0')
    mods, ds, root = _load('cpu', replace_layers)
    g = H.load_golden('tinyA')
    args = _args(ds, root, dev)
    net = mods['model'].Net(args)
    net.load_state_dict({k[3:]: th.tensor(v) for k, v in g.items() if k.startswith('sd.')})
    net = net.to(dev)
    enc, dec, labels = ds.data_cv[0]['train']
    gr = ds.cv_specific_graphs[0]
    call = (enc.int().to(dev), dec.int().to(dev), gr['drug_graph'].to(dev), th.FloatTensor(ds.drug_sim_features).to(dev),
            ds.drug_feature.to(dev), gr['disease_graph'].to(dev), th.FloatTensor(ds.disease_sim_features).to(dev),
            ds.disease_feature.to(dev), gr['drug_feature_graph'].to(dev), gr['disease_feature_graph'].to(dev))
    _lib.reset_launch_count()
    net.eval()
    with th.no_grad():
        out = net(*call)
    for nm, t in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), out):
        assert H.rel_err(t.cpu(), g['fwd.' + nm]) <= FP32_TOL, nm
    net.train()
    out = net(*call)
    loss = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), labels.to(dev)) + args.beta * (
        mods['utils'].common_loss(out[1], out[2]) + mods['utils'].common_loss(out[3], out[4]))
    loss.backward()
    assert abs(float(loss) - float(g['loss'])) <= 1e-5
    assert _lib.launch_count() > 0                       # the aggregation ran in libdreamgnn.so, not in a library
    for k, p in net.named_parameters():
        if bool(g['hasgrad.' + k]):
            assert H.rel_err(p.grad.cpu(), g['grad.' + k]) <= FP32_TOL, k


@pytest.mark.gpu
def test_reference_train_loop_on_b200():
    """The reference's own `train(args, dataset, cv)` (train.py:154-395) -- its loop, augmentation, model, evaluation -- on
    the device through the shim + `dreamgnn_b200.layers`: runs, logs finite losses, returns metrics in range, and its
    first logged loss agrees with the reference-on-stand-in run of the same seed to the dropout noise."""
    from dreamgnn_b200 import _lib
    _lib.load()
    mods, ds, root = _load('cpu', replace_layers=True)
    g = H.load_golden('tinyA')
    args = _args(ds, root, th.device('cuda:0'), dropout=0.3, attention_dropout=0.1)
    args.train_max_iter = 4
    th.manual_seed(77)
    sink = io.StringIO()
    _lib.reset_launch_count()
    with rr.chdir(root), contextlib.redirect_stdout(sink):
        auroc, aupr = mods['train'].train(args, ds, 0)
    log = [ln for ln in sink.getvalue().splitlines() if ln.startswith('Iter=')]
    assert len(log) == 1 and 0.0 <= auroc <= 1.0 and 0.0 <= aupr <= 1.0
    got = float(log[0].split('Loss=')[1].split(',')[0])
    want = float(str(g['train.log'][0]).split('Loss=')[1].split(',')[0])
    assert np.isfinite(got) and abs(got - want) <= 0.05, (got, want)
    assert _lib.launch_count() > 100
