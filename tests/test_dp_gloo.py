"""Data-parallel host logic on CPU: world_size 2 over gloo."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from geniconet_b200.dp import GradBuckets, shard_sample_ids, broadcast_parameters
    torch.manual_seed(100 + rank)                       # deliberately different initial weights per rank
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                                torch.nn.Linear(32, 4))
    broadcast_parameters(model)
    w0 = [p.detach().clone() for p in model.parameters()]
    buckets = GradBuckets(model.parameters(), world, bucket_bytes=2048)     # several buckets
    assert len(buckets.buckets) > 1
    ids = shard_sample_ids(3, rank, world, 5)
    g = torch.Generator().manual_seed(0)
    data = torch.randn(64, 16, generator=g)
    x = data[[i % 64 for i in ids]]
    import copy
    twin = copy.deepcopy(model)                         # same weights, no communication: the purely local gradient
    for p in twin.parameters():
        p.grad = None
    twin(x).pow(2).mean().backward()
    local = [p.grad.detach().clone() for p in twin.parameters()]
    for it in range(2):                                 # two steps: the reset/re-arm path is exercised
        buckets.reset()
        loss = model(x).pow(2).mean()
        loss.backward()                                 # buckets are all-reduced from the autograd hooks
        buckets.finish()
    avg = [p.grad.detach().clone() for p in model.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, [t.numpy() for t in local])
    q.put((rank, ids, [t.numpy() for t in w0], [t.numpy() for t in avg], gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_buckets_average_over_two_ranks():
    import numpy as np
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ids0, w0, avg0, gath0), (r1, ids1, w1, avg1, _) = res
    assert set(ids0).isdisjoint(ids1) and ids0 == list(range(30, 35)) and ids1 == list(range(35, 40))
    for a, b in zip(w0, w1):
        assert np.array_equal(a, b)                      # broadcast made the replicas identical
    for k in range(len(avg0)):
        assert np.allclose(avg0[k], avg1[k], atol=1e-7)
        want = (gath0[0][k] + gath0[1][k]) / 2           # post-allreduce == mean of the per-rank gradients
        assert np.allclose(avg0[k], want, atol=1e-6)


def test_single_rank_is_a_noop():
    from geniconet_b200.dp import GradBuckets
    m = torch.nn.Linear(4, 4)
    b = GradBuckets(m.parameters(), 1)
    b.reset()
    m(torch.ones(2, 4)).sum().backward()
    g = m.weight.grad.clone()
    b.finish()
    assert torch.equal(g, m.weight.grad) and b.total_bytes() == 4 * (16 + 4)
