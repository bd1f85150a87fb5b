"""Per-parameter gradient cosines of the tcgen05 module-wise path and of the fused chain against the fp32 CUDA-core path."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import oracle_models as om
from geniconet_b200 import models as gm, losses, data, reparam
from geniconet_b200.ico_conv import set_impl
name = sys.argv[1] if len(sys.argv) > 1 else 'ico2ico_vae'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
level = 5
params = gm.default_params(name, level)
x, tgt = data.synthetic_batch(level, 0, B)
x, tgt = x.cuda(), tgt.cuda()
f = (params['ico']['factor_pos'], params['ico']['factor_nor'], params['ico']['factor_lap'])
if os.environ.get('DIAG_FACTORS'):                          # e.g. 1,0,0: position term only (well conditioned at random init)
    f = tuple(float(v) for v in os.environ['DIAG_FACTORS'].split(','))
if os.environ.get('DIAG_TF32', '1') == '0':                 # make the stock 1x1 head (cuDNN) compute in fp32 as well
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for tag, fused, impl, head in (('fp32', False, 'simt', True), ('tc', False, 'auto', True), ('fused', True, 'auto', True), ('fused_stockhead', True, 'auto', False)):
    gm.set_fused(fused, head)
    torch.manual_seed(3)
    mod = set_impl(om.fill_params_deterministic(getattr(gm, name)(params)).cuda().train(), impl)
    crit = losses.P2PKLD_Loss(level, *f, 1.0) if name == 'ico2ico_vae' else losses.P2P_Loss(level, *f)
    reparam.manual_seed(11)
    loss = crit(mod(x), tgt)
    loss.backward()
    torch.cuda.synchronize()
    res[tag] = (loss.item(), {k: p.grad.detach().double().flatten() for k, p in mod.named_parameters()})
print('loss', {k: v[0] for k, v in res.items()})
def cos(a, b):
    return (a @ b / (a.norm() * b.norm() + 1e-300)).item()
for k in res['fp32'][1]:
    r = res['fp32'][1][k]
    if r.norm() < 1e-6:
        continue
    print('%-28s |g| %.3e   tc %.5f  fused %.5f  fused(stock head) %.5f  fused-vs-tc %.5f   norm ratio fused/fp32 %.4f' % (
        k, r.norm().item(), cos(res['tc'][1][k], r), cos(res['fused'][1][k], r), cos(res['fused_stockhead'][1][k], r),
        cos(res['fused'][1][k], res['tc'][1][k]), (res['fused'][1][k].norm() / r.norm()).item()))
