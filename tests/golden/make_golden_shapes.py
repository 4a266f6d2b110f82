"""Freeze digests of the UNMODIFIED reference (/root/reference through the DGL stand-in in oracle/dgl) at the full
BASELINE.json dataset shapes -- lrssl 763 x 681, Gdataset 593 x 313, Cdataset 663 x 409, CLI-default model (768-dim
embeddings, gcn_agg_units 1024 -> 341-wide messages, nhid 768 / 128, 3 layers). Build container only:

    python tests/golden/make_golden_shapes.py [lrssl gdataset cdataset]

Inputs are regenerated from a seed by tests/shapes.py, so only OUTPUTS are stored, and of the large tensors only a
digest: the Euclidean norm plus 4 096 entries at fixed pseudo-random positions (tests/shapes.py:sample_index). Integer
structures (fold pairs, encoder edge lists, kNN COO indices) and the fp32 normalisers / adjacency values are stored as
SHA-256 hashes -- their contract is bit-exactness. Per shape: `Net.forward` in eval mode (5 outputs), the training loss
and every parameter gradient with all dropout p = 0, and the reference's own `evaluate` (AUROC / AUPR, sklearn) on both
splits at those weights.
"""
import os
import sys
import tempfile

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_runner as rr  # noqa: E402
from tests import shapes as S  # noqa: E402


def build(name):
    spec = S.DATASETS[name]
    root = tempfile.mkdtemp(prefix='dg_shape_')
    rr.write_synthetic_mat(root, 'lrssl', **spec)              # the loader branch is chosen by name; the schema is shared
    mods, ds = rr.load_reference_dataset(root, 'lrssl', k=S.K_NEIGHBOR)
    out = {'meta.n_drug': np.int64(spec['n_drug']), 'meta.n_dis': np.int64(spec['n_dis']), 'meta.k': np.int64(S.K_NEIGHBOR)}
    cv = 0
    for split in ('train', 'test'):
        enc, dec, labels = ds.data_cv[cv][split]
        s, d = dec.edges()
        out[f'hash.{split}.pairs'] = np.array(S.sha(np.stack([s.numpy(), d.numpy()]).astype(np.int64)))
        out[f'hash.{split}.labels'] = np.array(S.sha(labels.numpy().astype(np.float32)))
        out[f'meta.{split}.n_pairs'] = np.int64(labels.numel())
        for c in enc.canonical_etypes:
            es, ed = enc.edges(etype=c)
            out[f'hash.{split}.enc.{c[1]}'] = np.array(S.sha(np.stack([es.numpy(), ed.numpy()]).astype(np.int64)))
        for nt in ('drug', 'disease'):
            out[f'hash.{split}.ci.{nt}'] = np.array(S.sha(enc.nodes[nt].data['ci'].numpy()))
            out[f'hash.{split}.cj.{nt}'] = np.array(S.sha(enc.nodes[nt].data['cj'].numpy()))
    graphs = ds.cv_specific_graphs[cv]
    for gk in ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph'):
        idx, val = graphs[gk]._indices().numpy(), graphs[gk]._values().numpy()
        o = np.lexsort((idx[1], idx[0]))
        out[f'hash.knn.{gk}.indices'] = np.array(S.sha(idx[:, o].astype(np.int64)))
        out[f'hash.knn.{gk}.values'] = np.array(S.sha(val[o].astype(np.float32)))
        out[f'meta.knn.{gk}.nnz'] = np.int64(val.size)
    out['hash.feat.drug'] = np.array(S.sha(ds.drug_feature.numpy()))
    out['hash.feat.disease'] = np.array(S.sha(ds.disease_feature.numpy()))

    import argparse
    args = argparse.Namespace(model_activation='leaky', gcn_agg_accum='sum', share_param=True, device='cpu', dropout=0.0,
                              attention_dropout=0.0, beta=0.001, **S.NET)
    args.src_in_units, args.dst_in_units = ds.drug_feature_shape[1], ds.disease_feature_shape[1]
    args.fdim_drug, args.fdim_disease = ds.drug_feature_shape[0], ds.disease_feature_shape[0]
    args.rating_vals = ds.cv_data_dict[cv][2]
    th.manual_seed(2024)
    net = mods['model'].Net(args)
    out['hash.sd'] = np.array(S.state_dict_hash(net.state_dict()))
    enc, dec, labels = ds.data_cv[cv]['train']
    dsim, ssim = th.FloatTensor(ds.drug_sim_features), th.FloatTensor(ds.disease_sim_features)
    call = (enc.int(), dec.int(), graphs['drug_graph'], dsim, ds.drug_feature, graphs['disease_graph'], ssim,
            ds.disease_feature, graphs['drug_feature_graph'], graphs['disease_feature_graph'])
    net.eval()
    with th.no_grad():
        res = net(*call)
    for nm, t in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), res):
        out[f'fwd.{nm}.norm'], out[f'fwd.{nm}.samples'] = S.digest_entry('fwd.' + nm, t)
    net.train()
    res = net(*call)
    loss = th.nn.BCEWithLogitsLoss()(res[0].squeeze(-1), labels) + args.beta * (
        mods['utils'].common_loss(res[1], res[2]) + mods['utils'].common_loss(res[3], res[4]))
    loss.backward()
    out['loss'] = np.float64(loss.item())
    for k_, p in net.named_parameters():
        out['hasgrad.' + k_] = np.bool_(p.grad is not None)
        if p.grad is not None:
            out[f'grad.{k_}.norm'], out[f'grad.{k_}.samples'] = S.digest_entry('grad.' + k_, p.grad)
    for split in ('train', 'test'):
        auroc, aupr = mods['evaluation'].evaluate(
            args, net, {'test': ds.data_cv[cv][split]}, graphs['drug_graph'], ds.drug_feature, dsim, graphs['disease_graph'],
            ds.disease_feature, ssim, graphs['drug_feature_graph'], graphs['disease_feature_graph'])
        out[f'eval.{split}.auroc'], out[f'eval.{split}.aupr'] = np.float64(auroc), np.float64(aupr)
    np.savez_compressed(os.path.join(HERE, 'shape_%s.npz' % name), **out)
    print(name, 'written:', len(out), 'entries; loss %.6f; test AUROC %.4f AUPR %.4f'
          % (out['loss'], out['eval.test.auroc'], out['eval.test.aupr']))


if __name__ == '__main__':
    for nm in (sys.argv[1:] or list(S.DATASETS)):
        build(nm)
