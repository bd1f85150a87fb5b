"""GPU parity (run with -m gpu): every hot-path op through the C ABI vs the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import icocnn_ref
import oracle_models as om

pytestmark = pytest.mark.gpu

# fp32 CUDA-core path vs fp32 oracle
RTOL32, ATOL32 = 1e-4, 1e-5


def _conv_pair(cin, cout, stride, level, cm, impl):
    from geniconet_b200.ico_conv import IcoConvS2S
    torch.manual_seed(1)
    ref = icocnn_ref.IcoConvS2S(cin, cout, stride, True, level, cm)
    mod = IcoConvS2S(cin, cout, stride, True, level, cm, impl=impl).cuda()
    mod.load_state_dict(ref.state_dict())
    return ref, mod


def _check_conv(cin, cout, stride, level, cm, B, impl, rtol, atol, layout='cl'):
    ref, mod = _conv_pair(cin, cout, stride, level, cm, impl)
    n = 2 ** level
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, cin, 5 * n, 2 * n, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    xc = x.cuda()
    if layout == 'cl':
        xc = xc.contiguous(memory_format=torch.channels_last)
    xc.requires_grad_(True)
    yc = mod(xc)
    assert yc.shape == yr.shape
    yc.backward(gy.cuda())
    torch.cuda.synchronize()

    def close(a, b, what):
        a = a.detach().cpu()
        scale = b.abs().max().item()
        err = (a - b).abs().max().item()
        assert err <= atol * max(1.0, scale) + rtol * scale, '%s: max err %.3e (scale %.3e)' % (what, err, scale)
    close(yc, yr.detach(), 'fwd')
    close(xc.grad, xr.grad, 'dgrad')
    close(mod.weight.grad, ref.weight.grad, 'wgrad')
    close(mod.bias.grad, ref.bias.grad, 'bgrad')


SIMT_CASES = [
    # (cin, cout, stride, level, corner_mode, B, layout)
    (3, 64, 1, 5, 'average', 2, 'nchw'),      # encoder.0 (models.py:104), NCHW xyz input
    (3, 16, 1, 2, 'zeros', 5, 'nchw'),
    (8, 12, 1, 3, 'average', 3, 'cl'),
    (8, 12, 2, 3, 'average', 3, 'cl'),
    (5, 7, 2, 2, 'zeros', 6, 'nchw'),          # ragged channels, partial sample group
    (16, 8, 1, 1, 'average', 19, 'cl'),        # level 1: 16 samples per tile group, ragged batch
    (64, 64, 1, 4, 'average', 2, 'cl'),
    (64, 128, 2, 4, 'average', 2, 'cl'),
]


@pytest.mark.parametrize('case', SIMT_CASES)
def test_hexconv_simt_matches_oracle(case):
    cin, cout, stride, level, cm, B, layout = case
    _check_conv(cin, cout, stride, level, cm, B, 'simt', RTOL32, ATOL32, layout)


NARROW_CASES = [
    # the xyz input layer (models.py:104) through the warp-level kernels of gin_narrow.cuh (impl 'auto', Cin == 3)
    (3, 64, 1, 5, 'average', 3, 'nchw'),
    (3, 64, 1, 5, 'zeros', 2, 'cl'),
    (3, 64, 1, 2, 'average', 5, 'nchw'),       # level 2: several samples per tile group, ragged batch
    (3, 64, 1, 1, 'average', 19, 'nchw'),
    (3, 128, 1, 3, 'average', 3, 'nchw'),
    (3, 64, 2, 3, 'average', 4, 'nchw'),       # stride 2 uses the same tables
]


@pytest.mark.parametrize('case', NARROW_CASES)
def test_hexconv_narrow_matches_oracle(case):
    cin, cout, stride, level, cm, B, layout = case
    _check_conv(cin, cout, stride, level, cm, B, 'auto', RTOL32, ATOL32, layout)


def test_hexconv_empty_batch():
    from geniconet_b200.ico_conv import IcoConvS2S
    mod = IcoConvS2S(8, 8, 1, True, 2, 'average', impl='simt').cuda()
    y = mod(torch.zeros(0, 8, 20, 8, device='cuda'))
    assert tuple(y.shape) == (0, 8, 20, 8)


def test_hexconv_rejects_cpu_and_bad_shapes():
    from geniconet_b200.ico_conv import IcoConvS2S
    mod = IcoConvS2S(8, 8, 1, True, 2, 'average')
    with pytest.raises(RuntimeError):
        mod(torch.zeros(1, 8, 20, 8))
    with pytest.raises(ValueError):
        mod.cuda()(torch.zeros(1, 8, 40, 16, device='cuda'))
    with pytest.raises(ValueError):
        IcoConvS2S(8, 8, 3, True, 2, 'average')
    with pytest.raises(ValueError):
        IcoConvS2S(8, 8, 1, True, 2, 'mirror')


@pytest.mark.parametrize('level,cm,C,B', [(2, 'average', 256, 3), (3, 'zeros', 8, 2), (4, 'average', 128, 2), (1, 'average', 4, 5)])
def test_upsample_matches_oracle(level, cm, C, B):
    from geniconet_b200.ico_conv import IcoUpsampleS2S
    ref = icocnn_ref.IcoUpsampleS2S(C, level, cm)
    mod = IcoUpsampleS2S(C, level, cm)
    n = 2 ** level
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, 5 * n, 2 * n, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    yc = mod(xc)
    yc.backward(gy.cuda())
    assert torch.allclose(yc.cpu(), yr.detach(), rtol=1e-6, atol=1e-6)
    assert torch.allclose(xc.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('level,B,factors', [(2, 3, (0.6, 0.2, 0.2)), (3, 2, (1., 0., 0.)), (5, 2, (0.6, 0.2, 0.2))])
@pytest.mark.parametrize('layout', ['nchw', 'cl'])
def test_p2p_loss_matches_oracle(level, B, factors, layout):
    from geniconet_b200 import losses
    n = 2 ** level
    V = 10 * 4 ** level + 2
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, 5 * n, 2 * n, generator=g) * 0.5
    tgt = torch.randn(B, 9, V, generator=g)
    xr = x.clone().requires_grad_(True)
    lr, parts = om.ref_p2p_loss(level, xr, tgt, *factors)
    lr.backward()
    crit = losses.P2P_Loss(level, *factors)
    xc = x.cuda()
    if layout == 'cl':
        xc = xc.contiguous(memory_format=torch.channels_last)
    xc.requires_grad_(True)
    lc = crit(xc, tgt.cuda())
    lc.backward()
    got = crit.get_last_losses()
    assert abs(lc.item() - lr.item()) <= 1e-5 * max(1, abs(lr.item()))
    for a, b in zip(got[:3], parts):
        assert abs(a - b.item()) <= 2e-5 * max(1, abs(b.item()))
    gerr = (xc.grad.cpu() - xr.grad).abs().max().item()
    assert gerr <= 1e-4 * xr.grad.abs().max().item() + 1e-9, gerr


def test_kld_and_reparam():
    from geniconet_b200 import losses
    from geniconet_b200.reparam import reparameterize
    g = torch.Generator().manual_seed(5)
    mu = torch.randn(3, 512, 20, 8, generator=g)
    lv = torch.randn(3, 512, 20, 8, generator=g) * 0.5
    mur, lvr = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    kr = om.ref_kld(mur, lvr)
    kr.backward()
    muc = mu.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    lvc = lv.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    kc = losses._KLDFn.apply(muc, lvc)
    kc.backward()
    assert abs(kc.item() - kr.item()) <= 1e-5 * abs(kr.item())
    assert torch.allclose(muc.grad.cpu(), mur.grad, rtol=1e-4, atol=1e-9)
    assert torch.allclose(lvc.grad.cpu(), lvr.grad, rtol=1e-4, atol=1e-9)
    # reparameterisation: z reproduces from the returned eps; eps is N(0,1); (seed, offset) is deterministic
    muc.grad = lvc.grad = None
    z, eps = reparameterize(muc, lvc, seed=42, offset=3, return_eps=True)
    z2, eps2 = reparameterize(muc, lvc, seed=42, offset=3, return_eps=True)
    z3, eps3 = reparameterize(muc, lvc, seed=42, offset=4, return_eps=True)
    assert torch.equal(eps, eps2) and torch.equal(z, z2) and not torch.equal(eps, eps3)
    e = eps.detach().cpu()
    zr = e * torch.exp(0.5 * lvr.detach()) + mur.detach()
    assert torch.allclose(z.detach().cpu(), zr, rtol=1e-5, atol=1e-6)
    assert abs(e.mean().item()) < 5e-3 and abs(e.std().item() - 1) < 5e-3 and abs((e ** 4).mean().item() - 3) < 0.05
    gz = torch.randn(z.shape, generator=g)
    z.backward(gz.cuda())
    mur.grad = lvr.grad = None
    (e * torch.exp(0.5 * lvr) + mur).backward(gz)
    assert torch.allclose(muc.grad.cpu(), mur.grad, rtol=1e-5, atol=1e-7)
    assert torch.allclose(lvc.grad.cpu(), lvr.grad, rtol=1e-4, atol=1e-7)


def test_output2vertices():
    from geniconet_b200.losses import output2vertices
    from oracle import ico_geometry_ref as geo
    x = torch.randn(2, 3, 160, 64)
    rings = torch.from_numpy(geo.pole_rings(5))
    flat = x.reshape(2, 3, -1)
    ref = torch.cat((flat, flat[:, :, rings].mean(-1)), dim=2).transpose(1, 2)
    got = output2vertices(5, x.cuda())
    assert torch.allclose(got.cpu(), ref, rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------- tcgen05 path
# 16-bit operands / fp32 accumulate: vs the fp32 oracle rtol 2e-2 of the tensor scale; vs an oracle fed operands
# rounded the way the kernels round them the only difference is summation order: 2e-3.  Forward-side operands (x, forward W)
# are fp16 by default, gradient-side operands (dy, dgrad W) bf16 (csrc/gin_common.cuh: fwd_fp16).
TC_CASES = [
    # (cin, cout, stride, level, corner_mode, B)  -- every (Cin,Cout,stride) of models.py at reduced batch
    (64, 64, 1, 3, 'average', 2),
    (64, 128, 2, 4, 'average', 2),
    (128, 128, 1, 3, 'average', 3),
    (128, 256, 2, 3, 'average', 2),
    (256, 256, 1, 2, 'average', 5),        # level 2: 4 samples per tile group, ragged batch
    (256, 256, 2, 3, 'average', 4),
    (256, 128, 1, 3, 'zeros', 2),
    (128, 64, 1, 4, 'average', 1),
    (256, 512, 2, 3, 'average', 3),
    (512, 256, 1, 3, 'average', 1),
]


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _fwd_round(t):
    from geniconet_b200 import _lib
    return t.to(_lib.forward_operand_dtype()).to(torch.float32)


@pytest.mark.parametrize('case', TC_CASES)
def test_hexconv_tc_matches_oracle(case):
    from geniconet_b200.ico_conv import IcoConvS2S
    cin, cout, stride, level, cm, B = case
    torch.manual_seed(2)
    ref = icocnn_ref.IcoConvS2S(cin, cout, stride, True, level, cm)
    mod = IcoConvS2S(cin, cout, stride, True, level, cm, impl='tc').cuda()
    mod.load_state_dict(ref.state_dict())
    n = 2 ** level
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, cin, 5 * n, 2 * n, generator=g)
    # reference A: true fp32
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    full = dict(y=yr.detach(), dx=xr.grad.clone(), dw=ref.weight.grad.clone(), db=ref.bias.grad.clone())
    # reference B: operands rounded where the kernels round them -- forward: x and W in the forward format (module refq);
    # dgrad: bf16(dy) and bf16(W), wgrad: bf16(x) and bf16(dy) (module refd: dx depends on W and dy only, dW on x and dy only)
    with torch.no_grad():
        refq = icocnn_ref.IcoConvS2S(cin, cout, stride, True, level, cm)
        refq.load_state_dict(ref.state_dict())
        refq.weight.copy_(_fwd_round(ref.weight))
        refd = icocnn_ref.IcoConvS2S(cin, cout, stride, True, level, cm)
        refd.load_state_dict(ref.state_dict())
        refd.weight.copy_(_bf16_round(ref.weight))
    with torch.no_grad():
        yq = refq(_fwd_round(x))
    xd = _bf16_round(x).requires_grad_(True)
    refd(xd).backward(_bf16_round(gy))
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    yc = mod(xc)
    yc.backward(gy.cuda())
    torch.cuda.synchronize()

    def rel(a, b):
        return ((a.detach().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    assert rel(yc, full['y']) < 2e-2 and rel(yc, yq.detach()) < 2e-3, ('fwd', rel(yc, full['y']), rel(yc, yq.detach()))
    assert rel(xc.grad, full['dx']) < 2e-2 and rel(xc.grad, xd.grad) < 2e-3, ('dgrad', rel(xc.grad, full['dx']), rel(xc.grad, xd.grad))
    assert rel(mod.weight.grad, full['dw']) < 2e-2 and rel(mod.weight.grad, refd.weight.grad) < 2e-3, \
        ('wgrad', rel(mod.weight.grad, full['dw']), rel(mod.weight.grad, refd.weight.grad))
    assert rel(mod.bias.grad, full['db']) < 1e-4


def test_reparam_draws_fresh_noise_under_graph_replay():
    """A captured training step is replayed with frozen scalar arguments; the Philox offset therefore also takes a device-side
    step counter (gin_reparam_fwd_step), so every replay draws new eps."""
    from geniconet_b200 import reparam
    mu = torch.zeros(4, 8, 20, 8, device='cuda')
    lv = torch.zeros_like(mu)
    reparam.manual_seed(5)
    reparam.reparameterize(mu, lv)                      # eager warm-up (creates the counter outside the capture)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            z = reparam.reparameterize(mu, lv)
    outs = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        outs.append(z.clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    assert abs(outs[2].std().item() - 1.0) < 0.05


@pytest.mark.parametrize('from_y', [0, 1])
@pytest.mark.parametrize('C,level,B,two', [(64, 3, 3, False), (128, 2, 5, True), (256, 2, 2, True)])
def test_fused_bn_act_matches_torch(C, level, B, two, from_y):
    """gin_bn_stats / gin_bn_act_fwd / gin_bn_act_bwd(_pair) against torch's own BatchNorm2d (training) + add + ReLU + autograd
    (models.py:37-39, 59-61): statistics, running buffers, the bf16 operand copy incl. pole rows, dgamma / dbeta and dy.
    from_y = 1: the backward re-evaluates the ReLU mask from y and the BatchNorm constants instead of reading the activation copy."""
    from geniconet_b200 import _lib, fused
    L = _lib.lib
    torch.manual_seed(0)
    n = 2 ** level
    P = 10 * 4 ** level
    dev = 'cuda'
    ld = 2 * C if two else C                                   # y1 is a column slice of a wider matrix in the `two` case
    ybuf = torch.randn(B * P, ld, device=dev) * 2 + 0.5
    y2 = torch.randn(B * P, C, device=dev) if two else None
    bns = [torch.nn.BatchNorm2d(C).to(dev).train() for _ in range(2)]
    for bn in bns:
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.uniform_(-0.3, 0.3)
    refbn = [torch.nn.BatchNorm2d(C).to(dev).train() for _ in range(2)]
    for a, b in zip(refbn, bns):
        a.load_state_dict(b.state_dict())
    col0 = C if two else 0

    def as_map(t):                                             # [B*P, C] -> [B, C, 5n, 2n]
        return t.view(B, 5 * n, 2 * n, C).permute(0, 3, 1, 2)
    # ---- torch reference
    y1r = ybuf[:, col0:col0 + C].clone().requires_grad_(True)
    y2r = y2.clone().requires_grad_(True) if two else None
    z = refbn[0](as_map(y1r))
    if two:
        z = z + refbn[1](as_map(y2r))
    out_r = torch.relu(z)
    gout = torch.randn_like(out_r)
    out_r.backward(gout)
    # ---- fused kernels
    stat1 = fused._bn_stats(ybuf, col0, ld, B * P, C, bns[0])
    stat2 = fused._bn_stats(y2, 0, C, B * P, C, bns[1]) if two else None
    out_b, out_f, out_w = fused._bn_act(ybuf, col0, ld, stat1, y2, 0, C, stat2, B, level, C, want_b=True, want_f=True)
    torch.cuda.synchronize()
    ref_flat = out_r.detach().permute(0, 2, 3, 1).reshape(B * P, C)
    assert torch.allclose(out_f, ref_flat, rtol=1e-4, atol=1e-4)
    out_b16 = out_b.view(_lib.forward_operand_dtype())           # the container is 16 bits wide; the format is the forward operand's
    assert torch.allclose(out_b16[:B * P].float(), ref_flat, rtol=1e-2, atol=1e-2)
    ring = [[k * n * 2 * n for k in range(5)], [k * n * 2 * n + (n - 1) * 2 * n + 2 * n - 1 for k in range(5)]]
    poles = torch.stack([ref_flat.view(B, P, C)[:, ring[p]].mean(1) for p in (0, 1)], 1).reshape(2 * B, C)     # [B][pole] rows
    assert torch.allclose(out_b16[B * P:].float(), poles, rtol=1e-2, atol=1e-2)
    assert torch.allclose(out_w.float(), out_b16.float(), rtol=1e-2, atol=1e-2)          # the bf16 twin (wgrad operand, ReLU mask)
    for a, b in zip(refbn[:2 if two else 1], bns):
        assert torch.allclose(a.running_mean, b.running_mean, rtol=1e-4, atol=1e-5)
        assert torch.allclose(a.running_var, b.running_var, rtol=1e-4, atol=1e-5)
        assert int(b.num_batches_tracked) == 1
    d = gout.permute(0, 2, 3, 1).reshape(B * P, C).contiguous()
    if two:
        dyA = torch.empty(B * P + 2 * B, C, dtype=torch.bfloat16, device=dev)
        dyB = torch.empty(B * P + 2 * B, C, dtype=torch.bfloat16, device=dev)
        bsA, bsB = torch.empty(4 * C, device=dev), torch.empty(4 * C, device=dev)
        ws = torch.empty(L.gin_bn_pair_ws_bytes(C), dtype=torch.uint8, device=dev)
        _lib.check(L.gin_bn_act_bwd_pair(d.data_ptr(), C, out_w.data_ptr(), ybuf.data_ptr() + 4 * col0, ld, stat1.data_ptr(), bsA.data_ptr(),
                                         dyA.data_ptr(), C, y2.data_ptr(), C, stat2.data_ptr(), bsB.data_ptr(), dyB.data_ptr(), C, 0, ws.data_ptr(),
                                         B, level, C, from_y, torch.cuda.current_stream().cuda_stream))
        pairs = [(bsA, dyA, y1r, refbn[0]), (bsB, dyB, y2r, refbn[1])]
    else:
        dyA = torch.empty(B * P + 2 * B, C, dtype=torch.bfloat16, device=dev)
        bsA, dyf = fused._bn_bwd(d, out_w, ybuf, col0, ld, stat1, B, level, C, dy_b=dyA, dy_b_col=0, ldo=C, want_f=True, mask_from_y=from_y)
        assert torch.allclose(dyf, y1r.grad, rtol=1e-3, atol=1e-5)
        pairs = [(bsA, dyA, y1r, refbn[0])]
    torch.cuda.synchronize()
    for bs, dyb, yr, rb in pairs:
        scale = yr.grad.abs().max().item()
        assert torch.allclose(bs[:C], rb.bias.grad, rtol=1e-3, atol=1e-3)             # dbeta
        assert torch.allclose(bs[C:2 * C], rb.weight.grad, rtol=1e-3, atol=1e-3)      # dgamma
        assert (dyb[:B * P].float() - yr.grad).abs().max().item() <= 1e-2 * scale
        gp = torch.stack([yr.grad.view(B, P, C)[:, ring[p]].mean(1) for p in (0, 1)], 1).reshape(2 * B, C)
        assert (dyb[B * P:].float() - gp).abs().max().item() <= 1e-2 * scale


@pytest.mark.parametrize('env', [{'GIN_SEAM': 'gather'}, {'GIN_SEAM': 'patch'}, {'GIN_TC_MODE': 'patch1'}, {'GIN_TC_MODE': 'gather'}, {'GIN_FWD_FP16': '0'}])
def test_alternative_kernel_paths_match_oracle(env):
    """The A/B kernel selections (read once per process from the environment): signature-sorted gather seam pass (the default is the regular-form pass through the patch kernel),
    first-generation patch kernels, first-generation gather kernels -- each in a fresh process, conv fwd / dgrad / wgrad of a
    stride-1 and a stride-2 layer against the oracle fed bf16-rounded operands (tests/diag/diag_conv.py)."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for spec in (['64', '128', '1', '3', '3'], ['64', '64', '2', '3', '2']):
        out = subprocess.run([sys.executable, os.path.join(root, 'tests', 'diag', 'diag_conv.py')] + spec, env=dict(os.environ, **env),
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        rels = dict(re.findall(r'^(fwd|dgrad|wgrad) max err \S+ rel (\S+)', out.stdout, flags=re.M))
        assert set(rels) == {'fwd', 'dgrad', 'wgrad'}, out.stdout
        for k, v in rels.items():
            assert float(v) < 2e-3, (env, spec, k, v)


@pytest.mark.parametrize('level,B', [(1, 3), (2, 1), (5, 4)])
def test_decoder_head_matches_float64(level, B):
    """Conv2d(64,3,1) -> Tanh (models.py:151-154) through gin_head_fwd / gin_head_bwd against the same maths in float64; level 1
    has P = 40, so the 32-pixel blocks straddle samples and the last one is ragged."""
    from geniconet_b200 import fused
    n = 2 ** level
    g = torch.Generator().manual_seed(level)
    head = torch.nn.Sequential(torch.nn.Conv2d(64, 3, kernel_size=(1, 1)), torch.nn.Tanh()).cuda()
    x = torch.randn(B, 64, 5 * n, 2 * n, generator=g).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gy = torch.randn(B, 3, 5 * n, 2 * n, generator=g).cuda()
    assert fused.head_supported(head, x)
    y = fused.run_head(head, x)
    assert y.shape == (B, 3, 5 * n, 2 * n) and y.is_contiguous()
    y.backward(gy)
    w64, b64, x64 = head[0].weight.detach().double().reshape(3, 64), head[0].bias.detach().double(), x.detach().double().requires_grad_(True)
    w64.requires_grad_(True); b64.requires_grad_(True)
    y64 = torch.tanh(torch.einsum('oc,bchw->bohw', w64, x64) + b64[None, :, None, None])
    y64.backward(gy.double())
    assert (y.double() - y64).abs().max().item() <= 2e-6
    assert (x.grad.double() - x64.grad).abs().max().item() <= 1e-5 * x64.grad.abs().max().item()
    assert (head[0].weight.grad.double().reshape(3, 64) - w64.grad).abs().max().item() <= 1e-4 * w64.grad.abs().max().item()
    assert (head[0].bias.grad.double() - b64.grad).abs().max().item() <= 1e-4 * max(1.0, b64.grad.abs().max().item())
    # a hooked or differently shaped head keeps the stock modules
    hooked = torch.nn.Sequential(torch.nn.Conv2d(64, 3, kernel_size=(1, 1)), torch.nn.Tanh()).cuda()
    hooked[1].register_forward_hook(lambda m, i, o: None)
    assert not fused.head_supported(hooked, x)
    assert not fused.head_supported(torch.nn.Sequential(torch.nn.Conv2d(32, 3, kernel_size=(1, 1)), torch.nn.Tanh()).cuda(), x)


@pytest.mark.gpu
@pytest.mark.parametrize('level,B', [(2, 3), (4, 2)])
def test_mesh_shim_ops_are_differentiable_and_match_oracle(level, B):
    """mesh.utils as the reference's unmodified losses.py calls it (losses.py:39,54,57,71-80): forward values and the gradient
    that flows back through compute_vertex_normals / compute_laplacian_batch, against the CPU oracle's autograd."""
    import mesh.utils as mu                               # the product's shim package
    from oracle import mesh_ref, ico_geometry_ref as geo
    faces = torch.from_numpy(geo.get_ico_faces(level))
    V = int(faces.max()) + 1
    g = torch.Generator().manual_seed(17 + level)
    base = torch.from_numpy(geo.get_icosahedral_grid(level)[0]).float()
    v0 = (base[None] * (0.5 + 0.1 * torch.rand(B, V, 1, generator=g)) + 0.02 * torch.randn(B, V, 3, generator=g))
    tgt = torch.randn(B, V, 9, generator=g) * 0.5
    adj_ref = mesh_ref.compute_adjacency_matrix_sparse(V, faces)
    adj = mu.compute_adjacency_matrix_sparse(V, faces)
    assert adj.is_sparse and torch.equal(adj.indices(), adj_ref.indices())
    crit = torch.nn.Module()
    crit.register_buffer('adj_mat', adj)                  # losses.py:40
    crit = crit.cuda()

    def loss_of(v, t, nrm_fn, lap_fn, a, f):
        nrm, lap = nrm_fn(v, f), lap_fn(v, a)
        l_nor = torch.mean(1 - torch.nn.functional.cosine_similarity(nrm, t[:, :, 3:6], dim=2))
        l_lap = torch.nn.functional.mse_loss(lap, t[:, :, 6:9])
        return 0.6 * torch.nn.functional.mse_loss(v, t[:, :, :3]) + 0.2 * l_nor + 0.2 * l_lap, nrm, lap
    vr = v0.clone().requires_grad_(True)
    lr, nr, lapr = loss_of(vr, tgt, mesh_ref.compute_vertex_normals, mesh_ref.compute_laplacian_batch, adj_ref, faces)
    lr.backward()
    vc = v0.clone().cuda().requires_grad_(True)
    lc, nc, lapc = loss_of(vc, tgt.cuda(), mu.compute_vertex_normals, mu.compute_laplacian_batch, crit.adj_mat, faces.cuda())
    lc.backward()
    assert torch.allclose(nc.detach().cpu(), nr.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(lapc.detach().cpu(), lapr.detach(), rtol=1e-4, atol=1e-6)
    assert abs(lc.item() - lr.item()) <= 1e-5 * abs(lr.item())
    gc, gr = vc.grad.cpu(), vr.grad
    assert (gc - gr).norm() <= 1e-4 * gr.norm(), ((gc - gr).norm() / gr.norm()).item()
    with pytest.raises(ValueError):
        mu.compute_vertex_normals(vc, faces[:-1])


@pytest.mark.gpu
def test_upsample_sharing_survives_in_place_ops():
    """upsample00 / upsample10 of an Up block read the same tensor (models.py:59-60) and share one kernel launch; an in-place
    operation on the first result, or on the input, between the two calls must not leak into the second result."""
    from geniconet_b200.ico_conv import IcoUpsampleS2S, clear_caches
    clear_caches()
    up0, up1 = IcoUpsampleS2S(8, 2, 'average'), IcoUpsampleS2S(8, 2, 'average')
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 8, 20, 8, generator=g).cuda()
    want = up0(x.clone()).clone()
    clear_caches()
    a = up0(x)
    b = up1(x)
    assert b is a                                       # shared: same input, nothing changed in between
    assert up1(x) is not a                              # the entry is handed out once
    clear_caches()
    a = up0(x)
    a.mul_(2.0)                                         # consumer modifies its map in place
    b = up1(x)
    assert b is not a and torch.equal(b, want)
    clear_caches()
    a = up0(x)
    x.add_(1.0)                                         # the input changes in place
    b = up1(x)
    assert b is not a and torch.equal(b, up0(x.clone()))
    clear_caches()
    xg = x.clone().requires_grad_(True)
    a, b = up0(xg), up1(xg)
    (a.sum() + 2 * b.sum()).backward()                  # autograd adds both consumers' gradients before the single backward
    xr = x.clone().requires_grad_(True)
    clear_caches()
    (3 * up0(xr).sum()).backward()
    assert torch.allclose(xg.grad, xr.grad)
    clear_caches()


@pytest.mark.gpu
@pytest.mark.parametrize('y16', [0, 1])
def test_stem_forward_with_epilogue_statistics(y16):
    """gin_hexconv_fwd_narrow_stats (the xyz layer inside a fused chain): output against the oracle conv, BatchNorm column sums
    from the epilogue against sums over the fp32 oracle output, fp16 or fp32 output."""
    import ctypes
    from geniconet_b200 import _lib
    from geniconet_b200.ico_conv import get_plan
    L = _lib.lib
    level, B, C = 4, 3, 64
    n, P = 2 ** level, 10 * 4 ** level
    torch.manual_seed(4)
    ref = icocnn_ref.IcoConvS2S(3, C, 1, True, level, 'average')
    x = torch.randn(B, 3, 5 * n, 2 * n)
    with torch.no_grad():
        yr = ref(x).permute(0, 2, 3, 1).reshape(B * P, C)
    plan = get_plan(_lib.PLAN_HEXCONV, level, 1, 'average', 'cuda')
    st = torch.cuda.current_stream().cuda_stream
    w = ref.weight.detach().cuda().contiguous()
    packed = torch.empty(L.gin_hexconv_packed_bytes(3, C), dtype=torch.uint8, device='cuda')
    _lib.check(L.gin_hexconv_pack_weights(w.data_ptr(), packed.data_ptr(), 3, C, st))
    xc = x.cuda()
    y = torch.empty(B * P, C, dtype=torch.float16 if y16 else torch.float32, device='cuda')
    parts = torch.empty(L.gin_hexconv_narrow_stats_ws_bytes(C) // 4, dtype=torch.float32, device='cuda')
    npart = ctypes.c_int(0)
    _lib.check(L.gin_hexconv_fwd_narrow_stats(plan.host_ptr, plan.dev_ptr, xc.data_ptr(), xc.stride(0), xc.stride(3), xc.stride(1), packed.data_ptr(),
                                              ref.bias.detach().cuda().data_ptr(), y.data_ptr(), y16, B, 3, C, parts.data_ptr(), ctypes.addressof(npart), st))
    torch.cuda.synchronize()
    assert npart.value > 0
    tol = 2e-3 if y16 else 1e-5
    assert torch.allclose(y.float().cpu(), yr, rtol=tol, atol=tol)
    sums = parts[:npart.value * 2 * C].view(npart.value, 2, C).double().sum(0).cpu()
    assert torch.allclose(sums[0], yr.double().sum(0), rtol=1e-4, atol=1e-3)
    assert torch.allclose(sums[1], (yr.double() ** 2).sum(0), rtol=1e-4, atol=1e-3)
