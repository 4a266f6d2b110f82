"""CPU, world_size=2 over gloo: the fold/seed sharding launcher deals jobs round-robin, every job runs on
exactly one rank, and rank-ordered results are gathered identically on all ranks."""
import os
import socket
import tempfile

import torch.distributed as dist
import torch.multiprocessing as mp

from dreamgnn_b200 import cv_shard


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    jobs = cv_shard.job_list([77, 31415, 888], 5)
    ran = []

    def run_job(seed, fold):
        ran.append((seed, fold))
        return 0.5 + fold / 100.0 + (seed % 7) / 1000.0, 0.1 + fold / 50.0
    res = cv_shard.run_sharded(jobs, run_job)
    assert ran == jobs[rank::world]
    assert [(s, f) for s, f, _, _ in res] == sorted(jobs)
    for s, f, a, p in res:
        assert abs(a - (0.5 + f / 100.0 + (s % 7) / 1000.0)) < 1e-12 and abs(p - (0.1 + f / 50.0)) < 1e-12
    if rank == 0:
        avgs = cv_shard.write_results(res, [77, 31415, 888], out_dir)
        assert len(avgs) == 3
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_is_a_partition():
    jobs = cv_shard.job_list(range(5), 10)
    for world in (1, 2, 4, 8):
        parts = [cv_shard.shard_jobs(jobs, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(jobs)
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_sharded_cv_two_ranks_gloo():
    out_dir = tempfile.mkdtemp(prefix='dg_cv_')
    mp.spawn(_worker, args=(2, _free_port(), out_dir), nprocs=2, join=True)
    summary = open(os.path.join(out_dir, 'summary_results.csv')).read().splitlines()
    assert summary[0] == 'experiment,seed,avg_auroc,avg_aupr' and summary[-1].startswith('overall,NA,')
    rows = open(os.path.join(out_dir, 'seed_77', 'experiment_results.csv')).read().splitlines()
    assert rows[0] == 'fold,auroc,aupr' and len(rows) == 7 and rows[-1].startswith('average,')


def test_single_process_needs_no_process_group():
    res = cv_shard.run_sharded([(1, 0), (1, 1)], lambda s, f: (0.9, 0.8), rank=0, world=1)
    assert res == [(1, 0, 0.9, 0.8), (1, 1, 0.9, 0.8)]
