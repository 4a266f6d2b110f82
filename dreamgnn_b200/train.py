"""Training loop with the reference's flags and `train(args, dataset, cv)` signature
(train.py:154-395 and the argparse block at train.py:402-452), built on the drop-in modules.

`train_iteration` is the unit bench.py times: augmentation (always on, train.py:254-277), forward,
BCE + beta * common losses (train.py:286-294), backward, clip_grad_norm_, Adam step (train.py:297-300).
"""
import argparse
import os
import time

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .augmentation import augment_graph_data
from .model import Net
from .optim import FusedAdam
from .utils import common_loss, common_loss_gram, setup_seed  # noqa: F401

FIXED_SEEDS = [77, 31415, 888, 1001, 9999, 0, 42, 123, 2024, 7]      # train.py:456


class LabelSmoothingBCELoss(nn.Module):
    """train.py:15-23."""

    def __init__(self, smoothing=0.0):
        super().__init__()
        self.smoothing = smoothing

    def forward(self, pred, target):
        return F.binary_cross_entropy_with_logits(pred, target * (1 - self.smoothing) + self.smoothing * 0.5)


def build_parser():
    """The reference's 36 flags with the same names, types, defaults and quirks (train.py:404-450)."""
    p = argparse.ArgumentParser(description='DREAM-GNN (B200-native hot path)')
    p.add_argument('--device', default='0', type=int)
    p.add_argument('--save_dir', type=str)
    p.add_argument('--save_id', type=int)
    p.add_argument('--model_activation', type=str, default='leaky')
    p.add_argument('--dropout', type=float, default=0.3)
    p.add_argument('--gcn_agg_units', type=int, default=1024)
    p.add_argument('--gcn_agg_accum', type=str, default='sum')
    p.add_argument('--gcn_out_units', type=int, default=128)
    p.add_argument('--train_max_iter', type=int, default=18000)
    p.add_argument('--train_grad_clip', type=float, default=1.0)
    p.add_argument('--train_valid_interval', type=int, default=250)
    p.add_argument('--gcn_agg_norm_symm', type=bool, default=True)
    p.add_argument('--nhid1', type=int, default=768)
    p.add_argument('--nhid2', type=int, default=128)
    p.add_argument('--train_lr', type=float, default=0.002)
    p.add_argument('--layers', type=int, default=3)
    p.add_argument('--share_param', default=True, action='store_true')
    p.add_argument('--data_name', default='Gdataset', type=str)
    p.add_argument('--num_neighbor', type=int, default=4)
    p.add_argument('--beta', type=float, default=0.001)
    p.add_argument('--weight_decay', type=float, default=1e-5)
    p.add_argument('--l2_reg_weight', type=float, default=0.0)
    p.add_argument('--attention_dropout', type=float, default=0.1)
    p.add_argument('--embedding_mode', type=str, default='pretrained', choices=['pretrained', 'random'])
    p.add_argument('--use_augmentation', action='store_true', default=False)
    p.add_argument('--aug_methods', type=str, nargs='+', default=['edge_dropout', 'feature_noise'],
                   choices=['edge_dropout', 'add_random_edges', 'feature_noise', 'graph_noise', 'feature_masking',
                            'mix_up'])
    p.add_argument('--edge_dropout_rate', type=float, default=0.1)
    p.add_argument('--add_edge_rate', type=float, default=0.03)
    p.add_argument('--feature_noise_scale', type=float, default=0.05)
    p.add_argument('--graph_noise_scale', type=float, default=0.03)
    p.add_argument('--feature_mask_rate', type=float, default=0.1)
    p.add_argument('--mixup_alpha', type=float, default=0.2)
    p.add_argument('--save_model', action='store_true')
    p.add_argument('--label_smoothing', type=float, default=0.0)
    p.add_argument('--generate_top_predictions', action='store_true', default=False)
    p.add_argument('--top_k', type=int, default=200)
    p.set_defaults(use_gate_attention=False)
    # additions of this implementation (optional; every reference flag above is unchanged)
    p.add_argument('--cuda_graph', action='store_true', default=False,
                   help='capture one training iteration into a CUDA graph and replay it (small, launch-bound datasets)')
    return p


def aug_params_from_args(args):
    """train.py:237-245."""
    g = lambda k, d: getattr(args, k, d)
    return {'edge_dropout_rate': g('edge_dropout_rate', 0.1), 'feature_noise_scale': g('feature_noise_scale', 0.05),
            'graph_noise_scale': g('graph_noise_scale', 0.02), 'add_edge_rate': g('add_edge_rate', 0.03),
            'feature_mask_rate': g('feature_mask_rate', 0.1), 'mixup_alpha': g('mixup_alpha', 0.2)}


class TrainState:
    """Device-resident inputs of one fold's training loop (what train.py:172-204 prepares)."""

    def __init__(self, enc_graph, dec_graph, labels, drug_graph, dis_graph, drug_feature_graph, disease_feature_graph,
                 drug_feat, dis_feat, drug_sim_feat, dis_sim_feat):
        self.enc_graph, self.dec_graph, self.labels = enc_graph, dec_graph, labels
        self.drug_graph, self.dis_graph = drug_graph, dis_graph
        self.drug_feature_graph, self.disease_feature_graph = drug_feature_graph, disease_feature_graph
        # row-major once, here: a Fortran-ordered array from loadmat keeps its strides through `x + noise`, and every
        # GEMM of every iteration would then start with a contiguous copy of its input
        dense = lambda t: t.contiguous() if isinstance(t, th.Tensor) and not t.is_sparse else t
        self.drug_feat, self.dis_feat = dense(drug_feat), dense(dis_feat)
        self.drug_sim_feat, self.dis_sim_feat = dense(drug_sim_feat), dense(dis_sim_feat)


def augment_state(state, aug_methods, aug_params):
    """The per-iteration augmentation of train.py:254-277 on the resident training inputs."""
    return augment_graph_data({
        'enc_graph': state.enc_graph, 'drug_graph': state.drug_graph, 'disease_graph': state.dis_graph,
        'drug_feature_graph': state.drug_feature_graph, 'disease_feature_graph': state.disease_feature_graph,
        'drug_feat': state.drug_feat, 'disease_feat': state.dis_feat,
        'drug_sim_feat': state.drug_sim_feat, 'disease_sim_feat': state.dis_sim_feat}, aug_methods, aug_params)


def train_iteration(model, optimizer, state, rel_loss_fn, aug_methods, aug_params, beta=0.001, grad_clip=1.0,
                    common_loss_fn=common_loss, aug=None):
    """One training iteration exactly as train.py:250-300; returns the (device) loss tensor. `aug` may carry
    an augmentation drawn ahead of time (graphed.GraphedIteration pipelines it beside the previous iteration)."""
    model.train()
    if state.labels.is_cuda:
        ops.begin_seed_pool(state.labels.device)      # this iteration's dropout seeds: one launch
    if aug is None:
        aug = augment_state(state, aug_methods, aug_params)
    pred, drug_out, drug_sim_out, dis_out, dis_sim_out = model(
        aug['enc_graph'], state.dec_graph, aug['drug_graph'], aug['drug_sim_feat'], aug['drug_feat'],
        aug['disease_graph'], aug['disease_sim_feat'], aug['disease_feat'], aug['drug_feature_graph'],
        aug['disease_feature_graph'], False)
    ready = getattr(model, 'routes_ready', None)
    if ready is not None:
        # the common losses need only the route outputs: a side branch forked where those were complete, beside the
        # attention + decoder + BCE (and, in the backward, beside the decoder's backward); joined for the total
        main = th.cuda.current_stream()
        side = ops.side_stream(main)
        side.wait_event(ready)
        with th.cuda.stream(side):
            c_drug, c_dis = common_loss_fn(drug_out, drug_sim_out), common_loss_fn(dis_out, dis_sim_out)
        rel = rel_loss_fn(pred.squeeze(-1), state.labels)
        main.wait_stream(side)
        c_drug.record_stream(main)
        c_dis.record_stream(main)
    else:
        rel = rel_loss_fn(pred.squeeze(-1), state.labels)
        c_drug, c_dis = common_loss_fn(drug_out, drug_sim_out), common_loss_fn(dis_out, dis_sim_out)
    total = ops.WeightedLossSum.apply(rel, c_drug, c_dis, beta)
    optimizer.zero_grad()
    total.backward()
    clip_and_step(model, optimizer, grad_clip)
    return total


def clip_and_step(model, optimizer, grad_clip):
    """train.py:297-300. `optim.FusedAdam` does both in two launches; any other optimizer takes torch's two calls."""
    if hasattr(optimizer, 'clip_and_step'):
        optimizer.clip_and_step(grad_clip)
    else:
        nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
        optimizer.step()


def make_train_state(dataset, cv, dev):
    """The device-resident training inputs of fold `cv` exactly as train.py:172-204 prepares them."""
    cv_data = dataset.data_cv[cv]
    graphs = dataset.cv_specific_graphs[cv]
    return TrainState(
        cv_data['train'][0].int().to(dev), cv_data['train'][1].int().to(dev), cv_data['train'][2].to(dev),
        graphs['drug_graph'].to(dev), graphs['disease_graph'].to(dev), graphs['drug_feature_graph'].to(dev),
        graphs['disease_feature_graph'].to(dev), dataset.drug_feature.to(dev), dataset.disease_feature.to(dev),
        th.as_tensor(dataset.drug_sim_features, dtype=th.float32).to(dev),
        th.as_tensor(dataset.disease_sim_features, dtype=th.float32).to(dev))


def train(args, dataset, cv):
    """Mirror of train.py:154-395: one fold. `dataset` must expose the reference loader's attributes
    (`drug_feature`, `disease_feature`, `*_feature_shape`, `drug_sim_features`, `disease_sim_features`,
    `cv_data_dict`, `data_cv`, `cv_specific_graphs`) with graphs built by dreamgnn_b200.graph_build."""
    args.src_in_units = dataset.drug_feature_shape[1]
    args.dst_in_units = dataset.disease_feature_shape[1]
    args.fdim_drug = dataset.drug_feature_shape[0]
    args.fdim_disease = dataset.disease_feature_shape[0]
    args.rating_vals = dataset.cv_data_dict[cv][2]
    cv_data = dataset.data_cv[cv]
    dev = args.device
    state = make_train_state(dataset, cv, dev)
    train_data_dict, test_data_dict = {'test': cv_data['train']}, {'test': cv_data['test']}
    model = Net(args=args).to(dev)
    # nn.BCEWithLogitsLoss / LabelSmoothingBCELoss (train.py:206-209) as one fused kernel each way
    rel_loss_fn = ops.FusedBCEWithLogitsLoss(smoothing=max(float(getattr(args, 'label_smoothing', 0.0)), 0.0))
    use_graph = bool(getattr(args, 'cuda_graph', False))
    entry_stream = th.cuda.current_stream() if use_graph else None
    if use_graph and entry_stream == th.cuda.default_stream():
        th.cuda.set_stream(th.cuda.Stream())       # graph capture cannot involve the legacy default stream
    try:
        return _train_fold(args, dataset, cv, dev, state, model, rel_loss_fn, use_graph, train_data_dict, test_data_dict)
    finally:
        if entry_stream is not None:
            th.cuda.set_stream(entry_stream)       # leave the caller's stream current again, also when a fold fails


def _train_fold(args, dataset, cv, dev, state, model, rel_loss_fn, use_graph, train_data_dict, test_data_dict):
    """The loop of train.py:211-395 for one fold (split from `train` so that the stream it switched is always restored)."""
    from .evaluation import evaluate
    # graph mode: the learning rate lives in a device tensor, which the captured Adam step reads on every replay and
    # ReduceLROnPlateau updates in place -- a scheduler step takes effect without re-capturing
    lr = th.tensor(float(args.train_lr), device=dev) if use_graph else args.train_lr
    optimizer = FusedAdam(model.parameters(), lr=lr, weight_decay=args.weight_decay)     # th.optim.Adam + clip, fused
    scheduler = th.optim.lr_scheduler.ReduceLROnPlateau(optimizer, 'max', patience=500, factor=0.5)
    aug_methods = getattr(args, 'aug_methods', ['edge_dropout', 'feature_noise'])
    aug_params = aug_params_from_args(args)
    log_path = os.path.join(args.save_dir, 'test_metric%s.csv' % args.save_id)
    log = open(log_path, 'w')
    log.write('iter,loss,train_auroc,train_aupr,test_auroc,test_aupr\n')
    best = dict(aupr=-1.0, auroc=0.0, it=0, train_aupr=0.0, train_auroc=0.0)
    start = time.perf_counter()
    graphed = None
    first_it = 1
    if use_graph:
        from .graphed import GraphedIteration
        graphed = GraphedIteration(model, optimizer, state, rel_loss_fn, aug_methods, aug_params, args.beta,
                                   args.train_grad_clip, warmup=0)
        # the eager warm-up iterations before the capture are real optimiser steps: they count as training iterations,
        # so a fold makes exactly train_max_iter - 1 updates as in the reference loop (train.py:250)
        first_it = 1 + graphed.eager_iterations
    try:
        for it in range(first_it, args.train_max_iter):
            if graphed is not None:
                total = graphed()
            else:
                total = train_iteration(model, optimizer, state, rel_loss_fn, aug_methods, aug_params, args.beta,
                                        args.train_grad_clip)
            if it % args.train_valid_interval == 0:
                ev = lambda d: evaluate(args, model, d, state.drug_graph, state.drug_feat, state.drug_sim_feat,
                                        state.dis_graph, state.dis_feat, state.dis_sim_feat, state.drug_feature_graph,
                                        state.disease_feature_graph)
                tr_auroc, tr_aupr = ev(train_data_dict)
                te_auroc, te_aupr = ev(test_data_dict)
                scheduler.step(te_aupr)
                log.write('%d,%.4f,%.4f,%.4f,%.4f,%.4f\n' % (it, total.item(), tr_auroc, tr_aupr, te_auroc, te_aupr))
                log.flush()
                print('Iter=%5d, Loss=%.4f, Train: AUROC=%.4f, AUPR=%.4f, Test: AUROC=%.4f, AUPR=%.4f'
                      % (it, total.item(), tr_auroc, tr_aupr, te_auroc, te_aupr))
                if te_aupr > best['aupr']:
                    best.update(aupr=te_aupr, auroc=te_auroc, it=it, train_aupr=tr_aupr, train_auroc=tr_auroc)
                    if getattr(args, 'save_model', False):
                        th.save(model.state_dict(), os.path.join(args.save_dir, 'best_model_fold%s.pth' % args.save_id))
    finally:
        log.close()
    print('Running time:', time.strftime('%H:%M:%S', time.gmtime(round(time.perf_counter() - start))))
    with open(os.path.join(args.save_dir, 'best_metric%s.csv' % args.save_id), 'w') as f:
        f.write('iter,train_auroc,train_aupr,test_auroc,test_aupr\n')
        f.write('%d,%.4f,%.4f,%.4f,%.4f\n' % (best['it'], best['train_auroc'], best['train_aupr'], best['auroc'],
                                                best['aupr']))
    if getattr(args, 'save_model', False) and getattr(args, 'generate_top_predictions', False):
        # train.py:369-395: reload the best checkpoint and export the top-K novel pairs
        from .predict import get_top_novel_predictions
        best_model = Net(args=args).to(dev)
        best_model.load_state_dict(th.load(os.path.join(args.save_dir, 'best_model_fold%s.pth' % args.save_id)))
        top = get_top_novel_predictions(args, best_model, dataset, cv, top_k=args.top_k)
        print('Top 5 novel predictions:\n%s' % top.head(5))
    return best['auroc'], best['aupr']
