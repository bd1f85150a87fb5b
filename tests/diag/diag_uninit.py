"""Looks for reads of uninitialised memory / run-to-run nondeterminism in a training step: the caching allocator's free blocks
are poisoned with a byte pattern (0xFF = NaN in fp32 and bf16, 0x00 = zeros) before each run of the same seeded step; gradients
that turn NaN or differ between patterns depend on memory the step never wrote."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import oracle_models as om                                                   # noqa: E402
from geniconet_b200 import models as gm, losses, data, reparam               # noqa: E402
from geniconet_b200.ico_conv import set_impl                                 # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'ico2ico_vae'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
level = int(sys.argv[3]) if len(sys.argv) > 3 else 5
params = gm.default_params(name, level)
x, tgt = data.synthetic_batch(level, 0, B)
x, tgt = x.cuda(), tgt.cuda()
f = (params['ico']['factor_pos'], params['ico']['factor_nor'], params['ico']['factor_lap'])
if os.environ.get('DIAG_FACTORS'):                          # e.g. 1,0,0: position term only (well conditioned at random init)
    f = tuple(float(v) for v in os.environ['DIAG_FACTORS'].split(','))


def poison(byte, gb=12):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    blocks = [torch.full((1 << 30,), byte, dtype=torch.uint8, device='cuda') for _ in range(gb)]
    small = [torch.full((n,), byte, dtype=torch.uint8, device='cuda') for n in (512, 4096, 65536, 1 << 20) for _ in range(64)]
    torch.cuda.synchronize()
    del blocks, small


def step(fused, head, impl):
    gm.set_fused(fused, head)
    torch.manual_seed(3)
    mod = set_impl(om.fill_params_deterministic(getattr(gm, name)(params)).cuda().train(), impl)
    crit = losses.P2PKLD_Loss(level, *f, 1.0) if name == 'ico2ico_vae' else losses.P2P_Loss(level, *f)
    reparam.manual_seed(11)
    loss = crit(mod(x), tgt)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {k: p.grad.detach().clone() for k, p in mod.named_parameters()}


for tag, fused, head, impl in (('tc modulewise', False, True, 'auto'), ('fused + gin head', True, True, 'auto'), ('fused + stock head', True, False, 'auto'),
                               ('simt modulewise', False, True, 'simt')):
    step(fused, head, impl)                                                   # warm-up: plans, caches
    runs = []
    for byte in (0xFF, 0x00, 0xFF):
        poison(byte)
        runs.append(step(fused, head, impl))
    (l0, g0), (l1, g1), (l2, g2) = runs
    nan = [k for k in g0 if not torch.isfinite(g0[k]).all()]
    diff_pattern = [k for k in g0 if not torch.equal(g0[k], g1[k])]
    diff_repeat = [k for k in g0 if not torch.equal(g0[k], g2[k])]
    print('%-20s loss %.8f %.8f %.8f | non-finite grads: %d | differ 0xFF vs 0x00: %d | differ 0xFF vs 0xFF again: %d of %d' % (
        tag, l0, l1, l2, len(nan), len(diff_pattern), len(diff_repeat), len(g0)))
    for k in (nan or diff_pattern or diff_repeat)[:6]:
        d = (g0[k].double() - g1[k].double()).norm() / (g0[k].double().norm() + 1e-300)
        print('    %-28s rel diff between patterns %.3e' % (k, d.item()))
