"""Bisects run-to-run nondeterminism of the fused chain: prefixes of the encoder+decoder module list are run twice from the same
seeded state; reports for each prefix whether the output and the gradients are bit-identical, and how far apart they are if not."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import oracle_models as om                                                   # noqa: E402
from geniconet_b200 import models as gm, data, fused                          # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'ico2ico'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
level = 5
params = gm.default_params(name, level)
x, _ = data.synthetic_batch(level, 0, B)
x = x.cuda()
torch.manual_seed(3)
model = om.fill_params_deterministic(getattr(gm, name)(params)).cuda().train()
mods = list(model.encoder) + (list(model.decoder) if name == 'ico2ico' else [])


def run(k):
    for p in model.parameters():
        p.grad = None
    xi = x.clone()
    y = fused.run_chain(xi, mods[:k])
    g = torch.Generator(device='cuda').manual_seed(5)
    w = torch.randn(y.shape, device='cuda', generator=g)
    (y * w).sum().backward()
    torch.cuda.synchronize()
    return y.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}, torch.zeros(1)


def rel(a, b):
    return ((a.double() - b.double()).norm() / (a.double().norm() + 1e-300)).item()


for k in range(3, len(mods) + 1):
    if not fused.chain_supported(mods[:k]):
        continue
    run(k)
    y0, g0, dx0 = run(k)
    junk = torch.full((1 << 28,), 0xFF, dtype=torch.uint8, device='cuda'); del junk
    y1, g1, dx1 = run(k)
    bad = sorted(((rel(g0[n], g1[n]), n) for n in g0 if not torch.equal(g0[n], g1[n])), reverse=True)
    print('prefix %2d (%s): output %s (rel %.2e) | dx %s (rel %.2e) | grads differing %d of %d%s' % (
        k, type(mods[k - 1]).__name__, 'same' if torch.equal(y0, y1) else 'DIFFERS', rel(y0, y1), 'same' if torch.equal(dx0, dx1) else 'DIFFERS',
        rel(dx0, dx1), len(bad), len(g0), ''.join('\n      %-30s rel %.2e' % (n, r) for r, n in bad[:4])))
