"""CPU, world_size=2 over gloo: host-side logic of the row-partitioned path -- partition index maps, and
the collectives' autograd adjoints (all-gather <-> reduce-scatter, all-reduce <-> all-reduce) that make the
per-rank backward passes add up to the gradient of the global objective."""
import os
import socket

import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from dreamgnn_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_index_maps():
    for world in (1, 2, 4):
        for rank in range(world):
            part = D.Partition({'drug': 16, 'disease': 8}, rank=rank, world=world)
            ids = th.arange(16)
            own = part.owned(ids, 'drug')
            assert int(own.sum()) == 16 // world
            assert th.equal(part.local(ids[own], 'drug'), th.arange(16 // world))
            # gathered buffer of R=2 relation-major chunks: rank-major, then relation, then local id
            n_loc = part.n_loc['drug']
            for r in (0, 1):
                g = part.gathered_index(ids, 'drug', 2, r)
                want = (ids // n_loc) * 2 * n_loc + r * n_loc + ids % n_loc
                assert th.equal(g, want)
            assert th.equal(part.gathered_index(ids, 'drug'), ids)          # R = 1: identity (contiguous blocks)
    try:
        D.Partition({'drug': 10}, rank=0, world=4)
        assert False
    except ValueError:
        pass


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    th.manual_seed(0)
    full = th.randn(world * 3, 4, dtype=th.float64)
    w = th.randn(4, 2, dtype=th.float64)
    # global objective: f(X) = sum_p || gather(X)[p-th selection] @ w ||^2 with a rank-dependent selection
    sel = [th.tensor([0, 3, 5]), th.tensor([1, 2, 4, 5])][rank]
    x = full[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    gathered = D.all_gather_rows(x)
    assert th.allclose(gathered.detach(), full)
    loss_p = ((gathered[sel] @ w) ** 2).sum()
    total = D.all_reduce_sum(loss_p.detach().clone())
    loss_p.backward()
    # reference: gradient of the summed objective w.r.t. this rank's rows
    xf = full.clone().requires_grad_(True)
    ref = sum(((xf[s] @ w) ** 2).sum() for s in (th.tensor([0, 3, 5]), th.tensor([1, 2, 4, 5])))
    ref.backward()
    assert th.allclose(x.grad, xf.grad[rank * 3:(rank + 1) * 3])
    assert th.allclose(total, ref.detach())
    # differentiable all-reduce: y = sum_p a_p ; objective sum_q g_q(y) -> d/da_p = sum_q g_q'(y)
    a = th.tensor([float(rank + 1)], dtype=th.float64, requires_grad=True)
    y = D.all_reduce_sum(a)
    ((rank + 2.0) * y ** 2).sum().backward()
    ysum = sum(r + 1.0 for r in range(world))
    assert th.allclose(a.grad, th.tensor([sum(2 * (q + 2.0) * ysum for q in range(world))], dtype=th.float64))
    # flat gradient all-reduce
    p1, p2 = th.nn.Parameter(th.zeros(3)), th.nn.Parameter(th.zeros(2, 2))
    p1.grad, p2.grad = th.full((3,), float(rank + 1)), th.full((2, 2), float(10 * (rank + 1)))
    D.all_reduce_gradients([p1, p2])
    assert th.equal(p1.grad, th.full((3,), 3.0)) and th.equal(p2.grad, th.full((2, 2), 30.0))
    dist.barrier()
    dist.destroy_process_group()


def test_collective_adjoints_two_ranks_gloo():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
