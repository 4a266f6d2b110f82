"""Fold / seed sharding across the GPUs of one box (SURVEY.md 8e, config 2 of BASELINE.json).

The reference runs its 10 seeds x 10 folds sequentially in one process (train.py:471, 500). Each
(seed, fold) job is independent -- own `Net` from `setup_seed(seed)`, own fold graphs -- so the job list is
dealt round-robin to one process per GPU with NO collective on the training path; the only communication is
one gather of (seed, fold, auroc, aupr) per job to rank 0, which writes the reference's result files
(train.py:522-527, 550-556).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        -m dreamgnn_b200.cv_shard --data_name lrssl [reference train.py flags] [--seeds 77 31415 ...] [--folds 10]
"""
import os

import numpy as np
import torch as th
import torch.distributed as dist


def job_seed(seed, fold):
    """RNG seed of one (seed, fold) job. The reference seeds once per experiment and lets the generators run on through
    the 10 folds (train.py:471-505), which a sharded run cannot replay; here fold 0 starts from `setup_seed(seed)` exactly
    as the reference's first fold does, and every later fold from its own deterministic seed, so no two folds of a seed
    share initial weights or dropout / augmentation streams, on any rank count."""
    return int(seed) if fold == 0 else (int(seed) * 1000003 + int(fold)) % (2 ** 31 - 1)


def job_list(seeds, n_folds):
    return [(s, f) for s in seeds for f in range(n_folds)]


def shard_jobs(jobs, rank, world):
    """Round-robin: job i goes to rank i % world (folds of one seed spread over all GPUs)."""
    return jobs[rank::world]


def gather_results(local, group=None):
    """local: list of (seed, fold, auroc, aupr). Returns the merged, job-ordered list on every rank.
    Works with gloo (CPU tensors) and nccl (tensors on the rank's GPU)."""
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(local)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = th.device('cuda', th.cuda.current_device()) if backend == 'nccl' else th.device('cpu')
    count = th.tensor([len(local)], dtype=th.int64, device=dev)
    counts = [th.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    m = max(int(c.item()) for c in counts)
    buf = th.zeros((max(m, 1), 4), dtype=th.float64, device=dev)
    if local:
        buf[:len(local)] = th.tensor(local, dtype=th.float64, device=dev)
    bufs = [th.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    out = []
    for c, b in zip(counts, bufs):
        for row in b[:int(c.item())].cpu().tolist():
            out.append((int(row[0]), int(row[1]), row[2], row[3]))
    return sorted(out)


def write_results(results, seeds, out_dir='seed_experiments'):
    """experiment_results.csv per seed + summary_results.csv (train.py:522-527, 550-556)."""
    os.makedirs(out_dir, exist_ok=True)
    avgs = []
    for s in seeds:
        rows = [(f, a, p) for (seed, f, a, p) in results if seed == s]
        if not rows:
            continue
        d = os.path.join(out_dir, 'seed_%d' % s)
        os.makedirs(d, exist_ok=True)
        avg_a, avg_p = float(np.mean([r[1] for r in rows])), float(np.mean([r[2] for r in rows]))
        with open(os.path.join(d, 'experiment_results.csv'), 'w') as fh:
            fh.write('fold,auroc,aupr\n')
            for f, a, p in sorted(rows):
                fh.write('%d,%.4f,%.4f\n' % (f + 1, a, p))
            fh.write('average,%.4f,%.4f\n' % (avg_a, avg_p))
        avgs.append((s, avg_a, avg_p))
    with open(os.path.join(out_dir, 'summary_results.csv'), 'w') as fh:
        fh.write('experiment,seed,avg_auroc,avg_aupr\n')
        for i, (s, a, p) in enumerate(avgs):
            fh.write('%d,%d,%.4f,%.4f\n' % (i + 1, s, a, p))
        if avgs:
            fh.write('overall,NA,%.4f,%.4f\n' % (np.mean([a for _, a, _ in avgs]), np.mean([p for _, _, p in avgs])))
    return avgs


def run_sharded(jobs, run_job, rank=None, world=None):
    """Run this rank's share of `jobs` with `run_job(seed, fold) -> (auroc, aupr)` and gather everything."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    local = []
    for seed, fold in shard_jobs(jobs, rank, world):
        auroc, aupr = run_job(seed, fold)
        local.append((seed, fold, float(auroc), float(aupr)))
    return gather_results(local)


def main(argv=None):
    from .data_loader import DrugDataLoader
    from .train import FIXED_SEEDS, build_parser, train
    from .utils import setup_seed
    parser = build_parser()
    parser.add_argument('--seeds', type=int, nargs='+', default=FIXED_SEEDS)
    parser.add_argument('--folds', type=int, default=10)
    args = parser.parse_args(argv)
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not th.cuda.is_available():
        raise RuntimeError('dreamgnn_b200 needs CUDA devices (no CPU fallback)')
    th.cuda.set_device(local)
    args.device = 'cuda:%d' % local
    if world > 1:
        dist.init_process_group('nccl', device_id=th.device(args.device))
    # KFold uses a fixed random_state, so folds are identical across seeds: build the loader once per process. Seed
    # first (the reference seeds before it builds its loader, train.py:471-476): a `.mat` without drug_embed /
    # disease_embed falls back to np.random embeddings, which must be the same on every rank and every run.
    setup_seed(args.seeds[0])
    dataset = DrugDataLoader(args.data_name, args.device, symm=args.gcn_agg_norm_symm, k=args.num_neighbor,
                             use_augmentation=args.use_augmentation, n_folds=args.folds)

    def run_job(seed, fold):
        setup_seed(job_seed(seed, fold))
        args.save_dir = os.path.join('seed_experiments', 'seed_%d' % seed)
        os.makedirs(args.save_dir, exist_ok=True)
        args.save_id = fold + 1
        return train(args, dataset, fold)

    results = run_sharded(job_list(args.seeds, args.folds), run_job, rank, world)
    if rank == 0:
        for s, a, p in write_results(results, args.seeds):
            print('seed %d: avg AUROC %.4f, avg AUPR %.4f' % (s, a, p))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
