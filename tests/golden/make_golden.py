"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference) on CPU through the DGL stand-in (oracle/dgl). Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no tests, fixtures or datasets, so these files are the
pinned answers for the hot path: graph / normaliser / kNN construction (data_loader.py), the
per-iteration augmentation under a fixed CPU seed (augmentation.py), `Net.forward` outputs in
eval mode and all parameter gradients of the training loss with dropout disabled
(model.py, layers.py, utils.py:87-95, train.py:286-294), and three full `train()` iterations.
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_runner as rr  # noqa: E402

CASES = {
    # shared-dims branch (layers.py:75-85), 3 layers, msg width 35 (not a multiple of 4)
    'tinyA': dict(data=dict(n_drug=60, n_dis=45, n_pos=220, embed_dim=48, sim_rank=12, seed=0), k=4,
                  net=dict(layers=3, gcn_agg_units=105, gcn_out_units=16, nhid1=40, nhid2=16)),
    # unequal in-dims -> per-etype weights branch (layers.py:86-97), 2 layers
    'tinyB': dict(data=dict(n_drug=37, n_dis=53, n_pos=150, embed_dim=(40, 24), sim_rank=10, seed=1), k=3,
                  net=dict(layers=2, gcn_agg_units=96, gcn_out_units=8, nhid1=20, nhid2=8)),
}


def make_args(net, dropout, attention_dropout, save_dir):
    return argparse.Namespace(
        model_activation='leaky', gcn_agg_accum='sum', share_param=True, device='cpu',
        dropout=dropout, attention_dropout=attention_dropout,
        train_max_iter=4, train_valid_interval=3, train_lr=0.002, weight_decay=1e-5, beta=0.001,
        train_grad_clip=1.0, save_dir=save_dir, save_id=1, save_model=False,
        generate_top_predictions=False, label_smoothing=0.0,
        aug_methods=['edge_dropout', 'feature_noise'], edge_dropout_rate=0.1, feature_noise_scale=0.05,
        graph_noise_scale=0.03, add_edge_rate=0.03, feature_mask_rate=0.1, mixup_alpha=0.2, **net)


def coo_arrays(t):
    return t._indices().numpy().copy(), t._values().numpy().copy()


def build_case(name, spec):
    root = tempfile.mkdtemp(prefix='dg_golden_')
    arrays = rr.write_synthetic_mat(root, 'lrssl', **spec['data'])
    mods, ds = rr.load_reference_dataset(root, 'lrssl', k=spec['k'])
    out = {}
    out['mat.didr'] = arrays['didr']
    # as scipy.io.loadmat hands them to the reference: float64, column-major (MATLAB order); the
    # layout matters because th.randn_like fills in memory order (augmentation.py:227)
    out['mat.drug'], out['mat.disease'] = ds.drug_sim_features, ds.disease_sim_features
    out['mat.drug_embed'], out['mat.disease_embed'] = ds.drug_embed, ds.disease_embed
    out['k'] = np.int64(spec['k'])
    cv = 0
    for split in ('train', 'test'):
        enc, dec, labels = ds.data_cv[cv][split]
        s, d = dec.edges()
        out[f'{split}.pairs'] = np.stack([s.numpy(), d.numpy()])
        out[f'{split}.labels'] = labels.numpy()
        for c in enc.canonical_etypes:
            es, ed = enc.edges(etype=c)
            out[f'{split}.enc.{c[1]}'] = np.stack([es.numpy(), ed.numpy()])
        for nt in ('drug', 'disease'):
            out[f'{split}.ci.{nt}'] = enc.nodes[nt].data['ci'].numpy()
            out[f'{split}.cj.{nt}'] = enc.nodes[nt].data['cj'].numpy()
    graphs = ds.cv_specific_graphs[cv]
    for gk in ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph'):
        idx, val = coo_arrays(graphs[gk])
        out[f'knn.{gk}.indices'], out[f'knn.{gk}.values'] = idx, val
    out['feat.drug'] = ds.drug_feature.numpy()
    out['feat.disease'] = ds.disease_feature.numpy()

    # ---- Net.forward (eval) and gradients (train mode, all dropout p = 0) -------------------
    args = make_args(spec['net'], dropout=0.0, attention_dropout=0.0, save_dir=root)
    args.src_in_units = ds.drug_feature_shape[1]
    args.dst_in_units = ds.disease_feature_shape[1]
    args.fdim_drug, args.fdim_disease = ds.drug_feature_shape[0], ds.disease_feature_shape[0]
    args.rating_vals = ds.cv_data_dict[cv][2]
    th.manual_seed(2024)
    net = mods['model'].Net(args)
    for k_, v in net.state_dict().items():
        out['sd.' + k_] = v.numpy().copy()
    enc, dec, labels = ds.data_cv[cv]['train']
    enc, dec = enc.int(), dec.int()
    dsim = th.FloatTensor(ds.drug_sim_features)
    ssim = th.FloatTensor(ds.disease_sim_features)
    call = (enc, dec, graphs['drug_graph'], dsim, ds.drug_feature, graphs['disease_graph'], ssim,
            ds.disease_feature, graphs['drug_feature_graph'], graphs['disease_feature_graph'])
    net.eval()
    with th.no_grad():
        res = net(*call)
    for nm, t in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), res):
        out['fwd.' + nm] = t.numpy().copy()
    net.train()
    res = net(*call)
    loss = th.nn.BCEWithLogitsLoss()(res[0].squeeze(-1), labels) + args.beta * (
        mods['utils'].common_loss(res[1], res[2]) + mods['utils'].common_loss(res[3], res[4]))
    loss.backward()
    out['loss'] = np.float64(loss.item())
    for k_, p in net.named_parameters():
        out['grad.' + k_] = (p.grad if p.grad is not None else th.zeros_like(p)).numpy().copy()
        out['hasgrad.' + k_] = np.bool_(p.grad is not None)

    # ---- per-iteration augmentation under a fixed CPU seed (augmentation.py:402-489) ---------
    gd = {'enc_graph': enc, 'drug_graph': graphs['drug_graph'], 'disease_graph': graphs['disease_graph'],
          'drug_feature_graph': graphs['drug_feature_graph'],
          'disease_feature_graph': graphs['disease_feature_graph'],
          'drug_feat': ds.drug_feature, 'disease_feat': ds.disease_feature,
          'drug_sim_feat': dsim, 'disease_sim_feat': ssim}
    th.manual_seed(123)
    # the perms the reference will draw, in its call order (same seed, same sizes)
    perm_sizes = [enc.number_of_edges(c) for c in enc.canonical_etypes] + [
        graphs[g]._values().numel() for g in ('drug_graph', 'disease_graph', 'drug_feature_graph',
                                              'disease_feature_graph')]
    perms = [th.randperm(n) for n in perm_sizes]
    th.manual_seed(123)
    aug = mods['augmentation'].augment_graph_data(
        gd, ['edge_dropout', 'feature_noise'],
        {'edge_dropout_rate': 0.1, 'feature_noise_scale': 0.05})
    for i, p in enumerate(perms):
        out[f'aug.perm.{i}'] = p.numpy()
    for c in aug['enc_graph'].canonical_etypes:
        es, ed = aug['enc_graph'].edges(etype=c)
        out[f'aug.enc.{c[1]}'] = np.stack([es.numpy(), ed.numpy()])
    for gk in ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph'):
        idx, val = coo_arrays(aug[gk])
        out[f'aug.knn.{gk}.indices'], out[f'aug.knn.{gk}.values'] = idx, val
    for fk in ('drug_feat', 'disease_feat', 'drug_sim_feat', 'disease_sim_feat'):
        out[f'aug.{fk}'] = aug[fk].numpy().copy()

    # ---- three full reference training iterations (train.py:154-395), default dropout -------
    targs = make_args(spec['net'], dropout=0.3, attention_dropout=0.1, save_dir=root)
    for nm in ('src_in_units', 'dst_in_units', 'fdim_drug', 'fdim_disease', 'rating_vals'):
        setattr(targs, nm, getattr(args, nm))
    th.manual_seed(77)
    net0 = mods['model'].Net(targs)          # train() builds the same Net from the same seed
    for k_, v in net0.state_dict().items():
        out['train.sd0.' + k_] = v.numpy().copy()
    out['train.rng0'] = th.get_rng_state().numpy()        # RNG state right after model init
    th.manual_seed(77)
    sink = io.StringIO()
    with rr.chdir(root), contextlib.redirect_stdout(sink):
        auroc, aupr = mods['train'].train(targs, ds, cv)
    log = [ln for ln in sink.getvalue().splitlines() if ln.startswith('Iter=')]
    out['train.auroc'], out['train.aupr'] = np.float64(auroc), np.float64(aupr)
    out['train.log'] = np.array(log)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'written:', len(out), 'arrays; loss', out['loss'], 'auroc/aupr', auroc, aupr, log)


if __name__ == '__main__':
    for name, spec in CASES.items():
        build_case(name, spec)
