"""geniconet_b200.optim.Adam (gin_adam_step: one launch for the whole parameter list) against torch.optim.Adam, the optimizer
the reference trains with (run.py:446, 250).  Both sides evaluate the same formula in fp32 with the bias corrections taken in
fp64, so they differ by a few ulp of the LARGEST term of each sum: rtol 1e-5 plus, for the moments, an absolute term of 2e-7
times the gradient scale (squared for exp_avg_sq) -- exp_avg is a signed moving average and its small entries are differences
of much larger ones (first GPU run: 3.6e-8 absolute on an entry of 1.5e-3 with gradients of order 1)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(7,), (4096,), (4097,), (32, 3, 7), (3,), (256, 256, 7), (100003,), (64, 33)]


def _params(seed, misaligned=True):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in SHAPES]
    if misaligned:                                           # a parameter whose storage starts 4 bytes off a 16-byte boundary
        base = torch.randn(1 + 5001, generator=g).cuda()
        ps.append(torch.nn.Parameter(base[1:]))
        assert ps[-1].data_ptr() % 16 == 4
    return ps


def _scale(i):
    return 10.0 ** ((i % 5) - 3)


def _grads(ps, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    return [torch.randn(p.shape, generator=g).cuda() * _scale(i) for i, p in enumerate(ps)]


def _close(a, b, what, atol=1e-8):
    torch.testing.assert_close(a, b, rtol=1e-5, atol=atol, msg=lambda m: '%s: %s' % (what, m))


@pytest.mark.parametrize('weight_decay', [0.0, 0.01])
def test_matches_torch_adam(weight_decay):
    from geniconet_b200.optim import Adam
    ours, ref = _params(0), _params(0)
    o = Adam(ours, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    r = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    for step in range(12):
        gs = _grads(ours, step)
        for p, q, g in zip(ours, ref, gs):
            p.grad = g.clone()                               # a new gradient tensor every step: the table is rebuilt each time
            q.grad = g.clone()
        o.step()
        r.step()
    torch.cuda.synchronize()
    for i, (p, q) in enumerate(zip(ours, ref)):
        _close(p, q, 'parameter %d %s' % (i, tuple(p.shape)))
        eff = _scale(i) + weight_decay * float(q.detach().abs().max())      # size of g + weight_decay * p
        _close(o.state[p]['exp_avg'], r.state[q]['exp_avg'], 'exp_avg %d' % i, atol=2e-7 * eff)
        _close(o.state[p]['exp_avg_sq'], r.state[q]['exp_avg_sq'], 'exp_avg_sq %d' % i, atol=2e-7 * eff ** 2)
        assert float(o.state[p]['step']) == 12.0 == float(r.state[q]['step'])


def test_parameters_without_gradient_are_skipped():
    from geniconet_b200.optim import Adam
    ps = _params(1, misaligned=False)
    before = [p.detach().clone() for p in ps]
    o = Adam(ps, lr=1e-2)
    for i, p in enumerate(ps):
        p.grad = torch.ones_like(p) if i % 2 == 0 else None
    o.step()
    torch.cuda.synchronize()
    for i, (p, b) in enumerate(zip(ps, before)):
        if i % 2 == 0:
            _close(p, b - 1e-2, 'first Adam step moves every element by lr')      # m/(sqrt(v)+eps) = 1 after bias correction
        else:
            assert torch.equal(p, b) and len(o.state[p]) == 0


def test_state_dict_is_interchangeable_with_torch():
    from geniconet_b200.optim import Adam
    ours, ref = _params(2), _params(2)
    o = Adam(ours, lr=3e-4)
    for step in range(3):
        for p, g in zip(ours, _grads(ours, step)):
            p.grad = g
        o.step()
    r = torch.optim.Adam(ref, lr=3e-4, capturable=True)
    with torch.no_grad():
        for p, q in zip(ours, ref):
            q.copy_(p)
    # deepcopy: Optimizer.load_state_dict keeps tensors that already have the right dtype and device BY REFERENCE (a checkpoint
    # that went through torch.save / torch.load is a copy anyway); two optimizers sharing exp_avg would both update it
    r.load_state_dict(copy.deepcopy(o.state_dict()))
    o2 = Adam(_params(2), lr=3e-4)
    with torch.no_grad():
        for p, q in zip(ours, o2.param_groups[0]['params']):
            q.copy_(p)
    o2.load_state_dict(copy.deepcopy(r.state_dict()))        # and back again
    third = o2.param_groups[0]['params']
    for step in range(3, 6):
        gs = _grads(ours, step)
        for p, q, s, g in zip(ours, ref, third, gs):
            p.grad, q.grad, s.grad = g, g.clone(), g.clone()
        o.step(); r.step(); o2.step()
    torch.cuda.synchronize()
    for i, (p, q, s) in enumerate(zip(ours, ref, third)):
        _close(p, q, 'torch continues our state, parameter %d' % i)
        _close(s, q, 'we continue torch state, parameter %d' % i)


def test_cuda_graph_replay_with_device_learning_rate():
    """Captured once, replayed with new gradients and a learning rate that changes on the device (run.py:252-254: CyclicLR)."""
    from geniconet_b200.optim import Adam
    ours, ref = _params(3), _params(3)
    lr_dev = torch.tensor(1e-3, device='cuda')
    o = Adam(ours, lr=lr_dev)
    r = torch.optim.Adam(ref, lr=1e-3)
    static = [torch.zeros_like(p) for p in ours]
    for p, g in zip(ours, static):
        p.grad = g
    o.prepare()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            o.step()
    torch.cuda.current_stream().wait_stream(side)
    for step in range(6):
        lr = 1e-3 * (1 + step % 3)
        lr_dev.fill_(lr)
        for group in r.param_groups:
            group['lr'] = lr
        for s, q, g in zip(static, ref, _grads(ours, step)):
            s.copy_(g)
            q.grad = g.clone()
        graph.replay()
        r.step()
    torch.cuda.synchronize()
    for i, (p, q) in enumerate(zip(ours, ref)):
        _close(p, q, 'parameter %d after 6 replays' % i)
        assert float(o.state[p]['step']) == 6.0
    # an eager step after the replays (different table, same state) keeps working
    for p, q, g in zip(ours, ref, _grads(ours, 99)):
        p.grad, q.grad = g.clone(), g.clone()
    lr_dev.fill_(1e-3)
    for group in r.param_groups:
        group['lr'] = 1e-3
    o.step(); r.step()
    graph.replay()                                           # and the graph still updates from ITS gradient tensors
    for q, s in zip(ref, static):
        q.grad = s.clone()
    r.step()
    torch.cuda.synchronize()
    for i, (p, q) in enumerate(zip(ours, ref)):
        _close(p, q, 'parameter %d after mixing eager steps and replays' % i)


def test_rejects_what_it_cannot_do():
    from geniconet_b200.optim import Adam
    with pytest.raises(ValueError):
        Adam([torch.nn.Parameter(torch.zeros(4, device='cuda'))], lr=-1.0)
    cpu = torch.nn.Parameter(torch.zeros(4))
    cpu.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match='no CPU path'):
        Adam([cpu]).step()
    half = torch.nn.Parameter(torch.zeros(4, device='cuda', dtype=torch.float16))
    half.grad = torch.ones_like(half)
    with pytest.raises(RuntimeError):
        Adam([half]).step()
