"""GPU, needs >= 2 devices (gpurun --gpus 2): the row-partitioned path over NCCL equals the single-GPU path
on the same graph and weights -- loss, and every parameter gradient after the flat all-reduce."""
import os
import socket

import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    th.cuda.set_device(rank)
    dev = th.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from dreamgnn_b200 import dist as D, synthetic
    from dreamgnn_b200.model import Net
    from dreamgnn_b200.utils import common_loss_gram
    spec = dict(kind='sparse', n_drug=512, n_dis=384, n_pairs=30000, f_drug=96, f_dis=64, k=5)
    w = synthetic.make_workload(spec, dev, seed=7)                     # same seed -> identical on every rank
    margs = synthetic.model_args(w, gcn_agg_units=96, gcn_out_units=16, nhid1=40, nhid2=16, dropout=0.0,
                                 attention_dropout=0.0)
    th.manual_seed(11)
    model = Net(margs).to(dev)
    part = D.Partition({'drug': spec['n_drug'], 'disease': spec['n_dis']})
    knn = {k: w[k] for k in ('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph')}
    state = D.PartitionedState(part, w['pairs'], w['labels'], knn, w['drug_feat'], w['dis_feat'], w['drug_sim_feat'],
                               w['dis_sim_feat'], dev)
    model.train()
    feats = (state.drug_feat, state.dis_feat, state.drug_sim_feat, state.dis_sim_feat)
    local, bce, common = D.partitioned_loss(model, state, state.enc_graph, state.knn, feats, beta=0.001)
    model.zero_grad()
    local.backward()
    params = list(model.parameters())
    D.all_reduce_gradients(params)
    total = float(D.all_reduce_sum(bce.detach()) + 0.001 * common.detach())
    got = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    # single-GPU reference on the full graph with the same weights
    st = synthetic.train_state(w, dev)
    model.zero_grad()
    out = model(st.enc_graph, st.dec_graph, st.drug_graph, st.drug_sim_feat, st.drug_feat, st.dis_graph,
                st.dis_sim_feat, st.dis_feat, st.drug_feature_graph, st.disease_feature_graph)
    ref = th.nn.BCEWithLogitsLoss()(out[0].squeeze(-1), st.labels) + 0.001 * (
        common_loss_gram(out[1], out[2]) + common_loss_gram(out[3], out[4]))
    ref.backward()
    assert abs(total - float(ref)) <= 1e-5 * max(1.0, abs(float(ref))), (total, float(ref))
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        err = float((got[k] - p.grad).norm() / (p.grad.norm() + 1e-30))
        assert err <= 1e-5, (k, err)
    # one full partitioned training iteration with augmentation runs and agrees across ranks
    opt = th.optim.Adam(model.parameters(), lr=0.002, weight_decay=1e-5)
    loss = D.train_iteration_partitioned(model, opt, state)
    both = [th.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(both, loss.reshape(1))
    assert th.isfinite(loss) and th.equal(both[0], both[1])
    chk = th.cat([p.detach().reshape(-1)[:4] for p in model.parameters()])
    ref_chk = chk.clone()
    dist.broadcast(ref_chk, 0)
    assert th.equal(chk, ref_chk)                                      # replicas stay bit-identical
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(th.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_row_partitioned_equals_single_gpu():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def _cv_worker(rank, world, port, root):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    os.chdir(root)
    from dreamgnn_b200 import cv_shard
    cv_shard.main(['--data_name', 'lrssl', '--train_max_iter', '5', '--train_valid_interval', '2', '--gcn_agg_units', '105',
                   '--gcn_out_units', '16', '--nhid1', '40', '--nhid2', '16', '--num_neighbor', '4', '--seeds', '77', '31415',
                   '--folds', '3'])


@pytest.mark.skipif(th.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_fold_sharded_cv_two_gpus():
    """The fold/seed launcher end to end over NCCL: 2 seeds x 3 folds dealt to 2 GPUs, results gathered on rank 0
    and written in the reference's CSV layout."""
    import tempfile

    import numpy as np
    import scipy.io as sio

    from tests import helpers as H
    g = H.load_golden('tinyA')
    root = tempfile.mkdtemp(prefix='dg_cv2_')
    d = os.path.join(root, 'raw_data', 'drug_data', 'lrssl')
    os.makedirs(d)
    names = np.empty((60, 1), dtype=object)
    for i in range(60):
        names[i, 0] = np.array(['DB%05d' % i])
    sio.savemat(os.path.join(d, 'lrssl.mat'), {'didr': g['mat.didr'], 'drug': g['mat.drug'], 'disease': g['mat.disease'],
                                               'drug_embed': g['mat.drug_embed'], 'disease_embed': g['mat.disease_embed'],
                                               'Wrname': names})
    mp.spawn(_cv_worker, args=(2, _free_port(), root), nprocs=2, join=True)
    summary = open(os.path.join(root, 'seed_experiments', 'summary_results.csv')).read().splitlines()
    assert summary[0] == 'experiment,seed,avg_auroc,avg_aupr' and len(summary) == 4
    rows = open(os.path.join(root, 'seed_experiments', 'seed_77', 'experiment_results.csv')).read().splitlines()
    assert len(rows) == 5 and rows[-1].startswith('average,')
    for r in rows[1:4]:
        a, p = map(float, r.split(',')[1:])
        assert 0.0 <= a <= 1.0 and 0.0 <= p <= 1.0
