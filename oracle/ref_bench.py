"""TEST / MEASUREMENT INFRASTRUCTURE -- time the UNMODIFIED reference's own training loop on the host CPU cores.

Used only by bench.py's `cpu_baseline` leg and `bench.py --impl reference`. The reference modules come from
/root/reference in the build container and from the archive staged by oracle/stage_ref.py on the GPU box
(`oracle/_ref/dreamgnn_reference.zip`); `import dgl` resolves to the pure-torch stand-in in oracle/dgl, i.e. this is
"reference code on a DGL stand-in" (real DGL is not installable offline -- SURVEY.md 8c/8d), `--device -1` path.

What is timed is `train.train(args, dataset, cv)` (train.py:154-395) itself, not a restatement: the function runs
unmodified for `warmup + steps + 1` iterations with the periodic evaluation switched off through its own flag
(`train_valid_interval` > `train_max_iter`), and the iteration boundaries are observed by wrapping the one name it calls
first in every iteration (`augment_graph_data`, train.py:268) with a timestamping pass-through. Iteration i lasts from
its augmentation call to the next one's, i.e. augmentation + forward + loss + backward + clip + Adam (train.py:250-300).

Datasets: the dense shapes (lrssl / Gdataset / Cdataset) go through the reference's own `DrugDataLoader` on a synthetic
`.mat` of its schema; the sparse synthetic shapes (syn20m samples) cannot (the loader materialises dense N_d x N_s
masks, data_loader.py:160-161), so a duck-typed dataset object is filled by calling the reference's own
`_generate_enc_graph` / `_generate_dec_graph` on the sampled pairs, with kNN graphs from oracle/restate.py.
"""
import argparse
import contextlib
import io
import os
import tempfile
import time

import numpy as np
import torch as th

from . import ref_runner as rr
from . import restate as R


def available():
    return rr.reference_available()


def source():
    root = rr.reference_root()
    return None if root is None else ('staged archive oracle/_ref/dreamgnn_reference.zip' if root.endswith('.zip') else root)


def _train_args(save_dir, n_iters):
    """The reference's CLI defaults (train.py:404-448) for one timed run."""
    return argparse.Namespace(
        device=th.device('cpu'), save_dir=save_dir, save_id=1, model_activation='leaky', dropout=0.3, gcn_agg_units=1024,
        gcn_agg_accum='sum', gcn_out_units=128, train_max_iter=n_iters + 1, train_grad_clip=1.0,
        train_valid_interval=10 ** 9, gcn_agg_norm_symm=True, nhid1=768, nhid2=128, train_lr=0.002, layers=3,
        share_param=True, num_neighbor=4, beta=0.001, weight_decay=1e-5, attention_dropout=0.1,
        aug_methods=['edge_dropout', 'feature_noise'], edge_dropout_rate=0.1, add_edge_rate=0.03, feature_noise_scale=0.05,
        graph_noise_scale=0.03, feature_mask_rate=0.1, mixup_alpha=0.2, save_model=False, label_smoothing=0.0,
        generate_top_predictions=False, top_k=200)


class _SparseDataset:
    """What train() reads from a DrugDataLoader (train.py:172-204), for a sampled sparse synthetic workload."""


def dense_dataset(mods, mat_arrays, k):
    import scipy.io as sio
    root = tempfile.mkdtemp(prefix='dg_refbench_')
    d = os.path.join(root, 'raw_data', 'drug_data', 'lrssl')
    os.makedirs(d, exist_ok=True)
    sio.savemat(os.path.join(d, 'lrssl.mat'), mat_arrays)
    with rr.chdir(root), contextlib.redirect_stdout(io.StringIO()):
        ds = mods['data_loader'].DrugDataLoader('lrssl', th.device('cpu'), symm=True, k=k)
    return ds, root


def sparse_dataset(mods, pairs, labels, n_drug, n_dis, drug_feat, dis_feat, knn_coo):
    """pairs: (drug ids, disease ids) numpy int64 in label order; labels {0,1}; feats fp32 tensors; knn_coo: the four
    (row, col, val, n) graphs in the order drug / disease / drug_feature / disease_feature."""
    loader_cls = mods['data_loader'].DrugDataLoader
    shell = object.__new__(loader_cls)                 # the reference's own graph builders, without its .mat loading
    shell._num_drug, shell._num_disease, shell._symm = n_drug, n_dis, True
    with contextlib.redirect_stdout(io.StringIO()):
        enc = shell._generate_enc_graph(pairs, labels, add_support=True)
        dec = shell._generate_dec_graph(pairs)
    ds = _SparseDataset()
    ds.drug_feature, ds.disease_feature = drug_feat, dis_feat
    # train() takes src/dst_in_units from shape[1] and the FGCN input width from shape[0] (train.py:170-173); at the
    # synthetic shapes the FGCN input is the feature matrix, so both are the feature width
    ds.drug_feature_shape = (drug_feat.shape[1], drug_feat.shape[1])
    ds.disease_feature_shape = (dis_feat.shape[1], dis_feat.shape[1])
    ds.drug_sim_features, ds.disease_sim_features = drug_feat.numpy(), dis_feat.numpy()
    ds.cv_data_dict = {0: [None, None, np.array([0, 1])]}
    ds.data_cv = {0: {'train': [enc, dec, th.FloatTensor(np.asarray(labels, dtype=np.float32))]}}
    ds.data_cv[0]['test'] = ds.data_cv[0]['train']
    names = ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph')
    ds.cv_specific_graphs = {0: {nm: th.sparse_coo_tensor(np.stack([r, c]), th.as_tensor(v), (n, n))
                                 for nm, (r, c, v, n) in zip(names, knn_coo)}}
    return ds


def time_train(mods, ds, steps, warmup, root=None, threads=None):
    """Seconds per iteration of the reference's `train()` (list of `steps` values) after `warmup` untimed iterations."""
    th.set_num_threads(threads or os.cpu_count() or 1)
    root = root or tempfile.mkdtemp(prefix='dg_refbench_')
    args = _train_args(root, warmup + steps + 1)
    train_mod = mods['train']
    real = train_mod.augment_graph_data
    stamps = []

    def stamped(*a, **kw):
        stamps.append(time.perf_counter())
        return real(*a, **kw)

    train_mod.augment_graph_data = stamped
    try:
        with rr.chdir(root), contextlib.redirect_stdout(io.StringIO()):
            train_mod.train(args, ds, 0)
    finally:
        train_mod.augment_graph_data = real
    per = np.diff(np.asarray(stamps))                   # warmup + steps intervals
    return [float(x) for x in per[warmup:warmup + steps]]


def aggregated_edges(ds, rate=0.1):
    """Edges aggregated by one training iteration: every GCMC SpMM (3 layers x 4 etypes, forward + backward) and every
    FGCN SpMM (4 graphs x 2 layers, forward + backward) over the KEPT edges (augmentation.py:48, :113)."""
    enc = ds.data_cv[0]['train'][0]
    kept = sum(R.dropout_num_keep(enc.number_of_edges(c), rate) for c in enc.canonical_etypes)
    knn_kept = sum(R.dropout_num_keep(g._values().numel(), rate) for g in ds.cv_specific_graphs[0].values()
                   if isinstance(g, th.Tensor))
    return 3 * 2 * kept + 4 * knn_kept
