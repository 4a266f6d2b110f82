"""GPU parity tests, module level: the drop-in modules (dreamgnn_b200.layers / model / augmentation /
graph_build) against the golden vectors frozen from the reference and against the CPU oracle."""
import argparse

import numpy as np
import pytest
import torch as th

from oracle import restate as R
from tests import helpers as H

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 2e-2
CFG = {'tinyA': dict(layers=3, gcn_agg_units=105, gcn_out_units=16, nhid1=40, nhid2=16),
       'tinyB': dict(layers=2, gcn_agg_units=96, gcn_out_units=8, nhid1=20, nhid2=8)}


@pytest.fixture(scope='module')
def dev():
    from dreamgnn_b200 import _lib
    _lib.load()
    return th.device('cuda:0')


@pytest.fixture(scope='module', params=H.CASES)
def case(request, dev):
    from dreamgnn_b200 import graph_build as GB
    name = request.param
    g = H.load_golden(name)
    n_d, n_s = g['feat.drug'].shape[0], g['feat.disease'].shape[0]
    k = int(g['k'])
    built = {}
    for split in ('train', 'test'):
        built[split] = (GB.generate_enc_graph(g[f'{split}.pairs'], g[f'{split}.labels'], n_d, n_s, dev),
                        GB.generate_dec_graph(g[f'{split}.pairs'], n_d, n_s, dev))
    knn = {'drug_graph': GB.create_similarity_graph(g['mat.drug'], k, dev),
           'disease_graph': GB.create_similarity_graph(g['mat.disease'], k, dev),
           'drug_feature_graph': GB.create_feature_similarity_graph(g['mat.drug_embed'], k, dev),
           'disease_feature_graph': GB.create_feature_similarity_graph(g['mat.disease_embed'], k, dev)}
    return name, g, built, knn


def _net(g, name, dev, dropout=0.0, attention_dropout=0.0, sd_prefix='sd.'):
    from dreamgnn_b200.model import Net
    args = argparse.Namespace(model_activation='leaky', gcn_agg_accum='sum', share_param=True, device=None,
                              dropout=dropout, attention_dropout=attention_dropout, rating_vals=[0, 1],
                              src_in_units=g['feat.drug'].shape[1], dst_in_units=g['feat.disease'].shape[1],
                              fdim_drug=g['feat.drug'].shape[0], fdim_disease=g['feat.disease'].shape[0], **CFG[name])
    net = Net(args)
    net.load_state_dict({k[len(sd_prefix):]: th.tensor(v) for k, v in g.items() if k.startswith(sd_prefix)})
    return net.to(dev)


def _inputs(g, built, knn, dev, split='train'):
    enc, dec = built[split]
    return (enc, dec, knn['drug_graph'], th.tensor(g['mat.drug'], dtype=th.float32).to(dev),
            th.tensor(g['feat.drug']).to(dev), knn['disease_graph'],
            th.tensor(g['mat.disease'], dtype=th.float32).to(dev), th.tensor(g['feat.disease']).to(dev),
            knn['drug_feature_graph'], knn['disease_feature_graph'])


def test_graph_construction_bit_exact(case):
    """CSR-side structure, ci/cj and the four kNN graphs against the reference's own loader output."""
    _, g, built, knn = case
    for split in ('train', 'test'):
        enc, dec = built[split]
        assert enc.canonical_etypes == [('disease', 'rev-0', 'drug'), ('disease', 'rev-1', 'drug'),
                                       ('drug', '0', 'disease'), ('drug', '1', 'disease')]
        for et in ('0', '1', 'rev-0', 'rev-1'):
            s, d = enc.edges(etype=et)
            np.testing.assert_array_equal(np.stack([s.cpu().numpy(), d.cpu().numpy()]), g[f'{split}.enc.{et}'])
        for nt in ('drug', 'disease'):
            np.testing.assert_array_equal(enc.nodes[nt].data['ci'].cpu().numpy(), g[f'{split}.ci.{nt}'])
            np.testing.assert_array_equal(enc.nodes[nt].data['cj'].cpu().numpy(), g[f'{split}.cj.{nt}'])
        s, d = dec.edges()
        np.testing.assert_array_equal(np.stack([s.cpu().numpy(), d.cpu().numpy()]), g[f'{split}.pairs'])
        # relation-block CSR == oracle CSR of the combined (dst, r*N_src + src) pairs
        blk = enc.block('disease')
        rows = np.concatenate([g[f'{split}.enc.0'][1], g[f'{split}.enc.1'][1]])
        cols = np.concatenate([g[f'{split}.enc.0'][0], g[f'{split}.enc.1'][0] + blk.n_src])
        indptr, indices, eid = R.csr_from_pairs(rows, cols, blk.n_dst)
        np.testing.assert_array_equal(blk.csr.indptr.cpu().numpy(), indptr)
        np.testing.assert_array_equal(blk.csr.indices.cpu().numpy(), indices)
        np.testing.assert_array_equal(blk.csr.eid.cpu().numpy(), eid)
    for key, t in knn.items():
        gr, gc, gv = H.canon_coo(g[f'knn.{key}.indices'][0], g[f'knn.{key}.indices'][1], g[f'knn.{key}.values'])
        idx = t._indices().cpu().numpy()
        np.testing.assert_array_equal(idx[0], gr)
        np.testing.assert_array_equal(idx[1], gc)
        np.testing.assert_array_equal(t._values().cpu().numpy(), gv)          # bit-exact fp32


def test_net_forward_eval(case, dev):
    name, g, built, knn = case
    net = _net(g, name, dev).eval()
    with th.no_grad():
        out = net(*_inputs(g, built, knn, dev))
    for nm, t in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), out):
        assert H.rel_err(t.cpu(), g['fwd.' + nm]) <= FP32_TOL, nm
    with th.no_grad():
        out_t = net(*_inputs(g, built, knn, dev, 'test'))
    inp = H.net_inputs(g, 'test')
    with th.no_grad():
        ref = R.net_forward(H.params(g), **inp, **H.NET_CFG[name])
    for a, b in zip(out_t, ref):
        assert H.rel_err(a.cpu(), b) <= FP32_TOL


def test_parallel_routes_are_bit_identical(case, dev):
    """Net.parallel_routes records the GCMC route and the FGCN route on two streams (forward and, through autograd's
    stream replay, backward): same kernels, same order inside each route -> bit-identical outputs and gradients."""
    from dreamgnn_b200.utils import common_loss
    name, g, built, knn = case
    labels = th.tensor(g['train.labels']).to(dev)
    res = {}
    for flag in (False, True):
        net = _net(g, name, dev).train()
        net.parallel_routes = flag
        with th.cuda.stream(th.cuda.Stream()):
            out = net(*_inputs(g, built, knn, dev))
            loss = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), labels) + 0.001 * (
                common_loss(out[1], out[2]) + common_loss(out[3], out[4]))
            loss.backward()
        th.cuda.synchronize()
        res[flag] = ([o.detach().clone() for o in out], {k: p.grad.clone() for k, p in net.named_parameters()
                                                          if p.grad is not None})
    for a, b in zip(res[False][0], res[True][0]):
        assert th.equal(a, b)
    assert res[False][1].keys() == res[True][1].keys()
    for k in res[False][1]:
        assert th.equal(res[False][1][k], res[True][1][k]), k


def test_net_gradients(case, dev):
    from dreamgnn_b200.utils import common_loss
    name, g, built, knn = case
    net = _net(g, name, dev).train()
    out = net(*_inputs(g, built, knn, dev))
    labels = th.tensor(g['train.labels']).to(dev)
    loss = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), labels) + 0.001 * (
        common_loss(out[1], out[2]) + common_loss(out[3], out[4]))
    assert abs(float(loss) - float(g['loss'])) <= 1e-5
    loss.backward()
    for k, p in net.named_parameters():
        if not bool(g['hasgrad.' + k]):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert H.rel_err(p.grad.cpu(), g['grad.' + k]) <= FP32_TOL, k
    # second backward pass is bit-identical (deterministic, atomic-free backward)
    grads = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    net.zero_grad()
    out = net(*_inputs(g, built, knn, dev))
    (th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), labels) + 0.001 * (
        common_loss(out[1], out[2]) + common_loss(out[3], out[4]))).backward()
    for k, p in net.named_parameters():
        if p.grad is not None:
            assert th.equal(p.grad, grads[k]), k


def test_bf16_message_path(case, dev):
    from dreamgnn_b200 import layers
    name, g, built, knn = case
    net = _net(g, name, dev).eval()
    layers.MESSAGE_DTYPE = th.bfloat16
    try:
        with th.no_grad():
            out = net(*_inputs(g, built, knn, dev))
    finally:
        layers.MESSAGE_DTYPE = th.float32
    for nm, t in zip(('pred', 'drug_out', 'dis_out'), (out[0], out[1], out[3])):
        assert H.rel_err(t.cpu(), g['fwd.' + nm]) <= BF16_TOL, nm


def test_per_relation_api_matches_fused_layer(case, dev):
    """GCMCGraphConv / HeteroGraphConv (the reference's per-etype call path) == fused relation-block path."""
    name, g, built, knn = case
    net = _net(g, name, dev).eval()
    enc = built['train'][0]
    layer = net.TGCN[0]
    x_d, x_s = th.tensor(g['feat.drug']).to(dev), th.tensor(g['feat.disease']).to(dev)
    with th.no_grad():
        fused = layer(enc, x_d, x_s)
        w = layer._relation_weights()
        args = {et: ((w[et] if layer.W_r is not None else None), False) for et in enc.etypes}
        per = layer.conv(enc, {'drug': x_d, 'disease': x_s}, mod_args=args)
        drug = layer.ifc(layer.agg_act(per['drug']))
        dis = layer.ufc(layer.agg_act(per['disease']))
    assert H.rel_err(fused[0].cpu(), drug.cpu()) <= FP32_TOL and H.rel_err(fused[1].cpu(), dis.cpu()) <= FP32_TOL


def test_augmentation_keeps_reference_edge_sets(case, dev):
    """Same permutations as the reference drew (golden, CPU generator) -> same kept-edge sets, ci/cj
    copied not recomputed, kNN graphs dropped consistently in both orientations."""
    from dreamgnn_b200 import ops
    from dreamgnn_b200.augmentation import num_keep_edges, _sparse_from_csr
    from dreamgnn_b200.layers import adjacency_csr
    _, g, built, knn = case
    enc = built['train'][0]
    perms = {}
    for i, c in enumerate(enc.canonical_etypes):
        p = th.tensor(g[f'aug.perm.{i}']).to(dev)
        perms[c] = (p, num_keep_edges(p.numel(), 0.1))
    dropped = enc.edge_dropout(perms)
    for c in enc.canonical_etypes:
        s, d = dropped.edges(etype=c)
        got = np.stack([s.cpu().numpy(), d.cpu().numpy()])
        want = g[f'aug.enc.{c[1]}']
        assert got.shape == want.shape and dropped.number_of_edges(c) == want.shape[1]
        key = lambda a: np.sort(a[0].astype(np.int64) * 100000 + a[1])
        np.testing.assert_array_equal(key(got), key(want))
    for nt in ('drug', 'disease'):
        assert th.equal(dropped.nodes[nt].data['ci'], enc.nodes[nt].data['ci'])
    # the dropped block's transpose is the true transpose of the dropped forward CSR
    blk = dropped.block('drug')
    t = blk.csr.transpose()
    tp, ti, _ = R.csr_from_pairs(blk.csr.indices.cpu().numpy(), blk.csr.rows().cpu().numpy(), blk.csr.n_cols)
    np.testing.assert_array_equal(t.indptr.cpu().numpy(), tp)
    np.testing.assert_array_equal(t.indices.cpu().numpy(), ti)
    for j, key in enumerate(H.KNN_KEYS):
        base = adjacency_csr(knn[key])
        p = th.tensor(g[f'aug.perm.{4 + j}']).to(dev)
        # the reference's perm indexes ITS COO order; ours is the same order (row-major sorted)
        k = num_keep_edges(p.numel(), 0.1)
        t = _sparse_from_csr(ops.csr_dropout(base, ops.keep_flags(base.nnz, [(p, k, 0)], dev), k), knn[key].shape)
        wr, wc, wv = H.canon_coo(g[f'aug.knn.{key}.indices'][0], g[f'aug.knn.{key}.indices'][1], g[f'aug.knn.{key}.values'])
        idx = t._indices().cpu().numpy()
        np.testing.assert_array_equal(idx[0], wr)
        np.testing.assert_array_equal(idx[1], wc)
        np.testing.assert_array_equal(t._values().cpu().numpy(), wv)


def test_augment_graph_data_api(case, dev):
    from dreamgnn_b200.augmentation import augment_graph_data
    from dreamgnn_b200.graph import HeteroGraph
    name, g, built, knn = case
    enc = built['train'][0]
    data = {'enc_graph': enc, **knn, 'drug_feat': th.tensor(g['feat.drug']).to(dev),
            'disease_feat': th.tensor(g['feat.disease']).to(dev),
            'drug_sim_feat': th.tensor(g['mat.drug'], dtype=th.float32).to(dev),
            'disease_sim_feat': th.tensor(g['mat.disease'], dtype=th.float32).to(dev)}
    th.manual_seed(0)
    out = augment_graph_data(data, ['edge_dropout', 'feature_noise'], {'edge_dropout_rate': 0.1, 'feature_noise_scale': 0.05})
    assert isinstance(out['enc_graph'], HeteroGraph)
    for c in enc.canonical_etypes:
        n = enc.number_of_edges(c)
        assert out['enc_graph'].number_of_edges(c) == R.dropout_num_keep(n, 0.1)
    for key in H.KNN_KEYS:
        assert out[key]._values().numel() == R.dropout_num_keep(knn[key]._values().numel(), 0.1)
    noise = (out['drug_feat'] - data['drug_feat']) / 0.05
    assert abs(float(noise.mean())) < 0.1 and abs(float(noise.std()) - 1.0) < 0.1
    # all optional methods run and a model step works on the augmented data
    th.manual_seed(1)
    out = augment_graph_data(data, ['edge_dropout', 'add_random_edges', 'feature_noise', 'graph_noise',
                                    'feature_masking', 'mix_up'],
                             {'edge_dropout_rate': 0.1, 'add_edge_rate': 0.03, 'feature_noise_scale': 0.05,
                              'graph_noise_scale': 0.03, 'feature_mask_rate': 0.1, 'mixup_alpha': 0.2})
    for c in enc.canonical_etypes:
        kept = R.dropout_num_keep(enc.number_of_edges(c), 0.1)
        assert out['enc_graph'].number_of_edges(c) == kept + max(1, int(kept * 0.03))
    net = _net(g, name, dev, dropout=0.3, attention_dropout=0.1).train()
    res = net(out['enc_graph'], built['train'][1], out['drug_graph'], out['drug_sim_feat'], out['drug_feat'],
              out['disease_graph'], out['disease_sim_feat'], out['disease_feat'], out['drug_feature_graph'],
              out['disease_feature_graph'])
    res[0].sum().backward()
    assert all(th.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)


def test_training_matches_oracle_without_randomness(case, dev):
    """Three optimiser steps (BCE + beta*common loss, clip 1.0, Adam lr 2e-3 wd 1e-5; train.py:286-300)
    with dropout and augmentation off: product on GPU vs oracle on CPU."""
    from dreamgnn_b200.utils import common_loss
    name, g, built, knn = case
    net = _net(g, name, dev, sd_prefix='train.sd0.').train()
    opt = th.optim.Adam(net.parameters(), lr=0.002, weight_decay=1e-5)
    labels = th.tensor(g['train.labels']).to(dev)
    P = {k[10:]: th.tensor(v) for k, v in g.items() if k.startswith('train.sd0.')}
    for k in list(P):
        if '.ifc.' in k:
            P[k] = P[k.replace('.ifc.', '.ufc.')]
    leaves = list({id(v): v for v in P.values()}.values())
    for v in leaves:
        v.requires_grad_(True)
    ropt = th.optim.Adam(leaves, lr=0.002, weight_decay=1e-5)
    inp = H.net_inputs(g)
    for _ in range(3):
        out = net(*_inputs(g, built, knn, dev))
        loss = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), labels) + 0.001 * (
            common_loss(out[1], out[2]) + common_loss(out[3], out[4]))
        opt.zero_grad()
        loss.backward()
        th.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        rout = R.net_forward(P, **inp, **H.NET_CFG[name], training=True)
        rloss = R.training_loss(rout, th.tensor(g['train.labels']))
        ropt.zero_grad()
        rloss.backward()
        th.nn.utils.clip_grad_norm_(leaves, 1.0)
        ropt.step()
        assert abs(float(loss) - float(rloss)) <= 2e-5
    sd = net.state_dict()
    for k, v in P.items():
        assert H.rel_err(sd[k].cpu(), v.detach()) <= 1e-4, k


def test_net_parity_with_tensor_core_projections(case, dev):
    """Same golden parity with every projection forced through the tcgen05 3xTF32 GEMM."""
    from dreamgnn_b200 import ops
    from dreamgnn_b200.utils import common_loss
    name, g, built, knn = case
    old = ops.GEMM_MIN_MACS
    ops.GEMM_MIN_MACS = 0
    try:
        net = _net(g, name, dev).train()
        out = net(*_inputs(g, built, knn, dev))
        labels = th.tensor(g['train.labels']).to(dev)
        loss = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), labels) + 0.001 * (
            common_loss(out[1], out[2]) + common_loss(out[3], out[4]))
        loss.backward()
    finally:
        ops.GEMM_MIN_MACS = old
    for nm, t in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), out):
        assert H.rel_err(t.detach().cpu(), g['fwd.' + nm]) <= FP32_TOL, nm
    for k, p in net.named_parameters():
        if bool(g['hasgrad.' + k]):
            assert H.rel_err(p.grad.cpu(), g['grad.' + k]) <= FP32_TOL, k
