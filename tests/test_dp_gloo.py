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


class _LateChain(torch.autograd.Function):
    """Stand-in for geniconet_b200.fused._Chain: ONE Function over several parameters that hands its gradients to the
    data-parallel sink while its backward is still running and returns them to autograd only at the end."""

    @staticmethod
    def forward(ctx, x, w1, w2, holder):
        ctx.save_for_backward(x, w1, w2)
        ctx.holder = holder
        return (x @ w1.t()).relu() @ w2.t()

    @staticmethod
    def backward(ctx, dy):
        from geniconet_b200 import fused
        x, w1, w2 = ctx.saved_tensors
        h = (x @ w1.t()).relu()
        dw2 = dy.t() @ h
        if fused._grad_sink is not None:
            fused._grad_sink([(ctx.holder[1], dw2)])              # the "last block" is differentiated first
        dh = (dy @ w2) * (h > 0)
        dw1 = dh.t() @ x
        if fused._grad_sink is not None:
            fused._grad_sink([(ctx.holder[0], dw1)])
        return dh @ w1, dw1, dw2, None


def _worker_early(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from geniconet_b200.dp import GradBuckets
    torch.manual_seed(7)
    w1, w2 = torch.nn.Parameter(torch.randn(32, 16) * 0.2), torch.nn.Parameter(torch.randn(4, 32) * 0.2)
    tail = torch.nn.Linear(4, 2)                                 # an ordinary module after the chain: arrives through autograd's hook
    params = [w1, w2] + list(tail.parameters())
    buckets = GradBuckets(params, world, bucket_bytes=64)        # one bucket per parameter
    launched = []
    orig = buckets._launch
    buckets._launch = lambda bi: (launched.append(bi), orig(bi))[1]
    x = torch.randn(6, 16, generator=torch.Generator().manual_seed(50 + rank))
    outs = []
    for it in range(2):
        buckets.reset()
        launched.clear()
        tail(_LateChain.apply(x, w1, w2, (w1, w2))).pow(2).mean().backward()
        order = list(launched)
        buckets.finish()
        outs.append([p.grad.detach().clone().numpy() for p in params])
    # the purely local gradient, no communication
    lw1, lw2 = w1.detach().clone().requires_grad_(True), w2.detach().clone().requires_grad_(True)
    import copy
    ltail = copy.deepcopy(tail)
    for p in ltail.parameters():
        p.grad = None
    ltail((x @ lw1.t()).relu() @ lw2.t()).pow(2).mean().backward()
    local = [lw1.grad.numpy(), lw2.grad.numpy()] + [p.grad.numpy() for p in ltail.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    q.put((rank, outs, gathered, order, [buckets._bucket_of[p] for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_early_gradients_from_a_fused_chain_are_averaged_once():
    """Gradients handed over from INSIDE a chain's backward (fused.set_grad_sink) start their bucket's all-reduce before autograd
    has assigned p.grad; the later post-accumulate hook must not count them twice, and the result is still the rank mean."""
    import numpy as np
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_early, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, outs0, gath, order, bucket_of), (_, outs1, _, _, _) = res
    for step in range(2):
        for k in range(len(outs0[step])):
            want = (gath[0][k] + gath[1][k]) / 2
            assert np.allclose(outs0[step][k], want, atol=1e-6) and np.allclose(outs1[step][k], want, atol=1e-6), (step, k)
    # the tail module's buckets go first (autograd reaches it first), then w2's, then w1's -- each exactly once
    assert len(order) == len(set(order)) == len(set(bucket_of))
    assert order.index(bucket_of[1]) < order.index(bucket_of[0])


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


def test_bucket_layout_keeps_sibling_weights_adjacent():
    """fused.weight_pairs: the two sibling convolutions' weight gradients come out of one wgrad GEMM as ONE tensor; the bucket
    lays each pair out side by side (never across a bucket boundary) so that kernel can write in place (grad_dest)."""
    from geniconet_b200.dp import GradBuckets
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (6, 10, 4, 8, 12)]
    b = GradBuckets(ps, 2, bucket_bytes=64, adjacent=[(ps[3], ps[0]), (ps[1],)])
    assert sorted(id(q) for q in b.params) == sorted(id(q) for q in ps)
    d = b.grad_dest((ps[3], ps[0]))
    assert d is not None and d.numel() == 14
    bi, off = b._slot[id(ps[3])]
    assert b._slot[id(ps[0])] == (bi, off + 8)
    views = dict((id(q), v) for _, qs, vs in b.buckets for q, v in zip(qs, vs))
    assert d.data_ptr() == views[id(ps[3])].data_ptr() and d[8:].data_ptr() == views[id(ps[0])].data_ptr()
    assert b.grad_dest((ps[0], ps[3])) is None and b.grad_dest((ps[1],)).numel() == 10
    assert GradBuckets(ps, 1, adjacent=[(ps[3], ps[0])]).grad_dest((ps[3], ps[0])) is None      # single rank: nothing to exchange
