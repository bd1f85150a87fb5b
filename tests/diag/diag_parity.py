import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import oracle_models as om
from geniconet_b200 import models as gm, losses, data
from geniconet_b200.ico_conv import set_impl
level = 5
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
params = gm.default_params('ico2ico', level)
ref = om.fill_params_deterministic(om.build_oracle_model('ico2ico', params))
x, tgt = data.synthetic_batch(level, 0, B)
acts_r = {}
def hook(store):
    def mk(name):
        def h(m, i, o):
            if torch.is_tensor(o): store[name] = o.detach().float().cpu()
        return h
    return mk
for n, m in ref.named_modules():
    m.register_forward_hook(hook(acts_r)(n))
out_r = ref(x); loss_r, _ = om.ref_p2p_loss(level, out_r, tgt, 1., 0., 0.); loss_r.backward()
gr = {k: p.grad.clone() for k, p in ref.named_parameters()}
for impl in ('simt', 'auto'):
    mod = gm.ico2ico(params); mod.load_state_dict(ref.state_dict()); mod = set_impl(mod.cuda(), impl)
    acts = {}
    for n, m in mod.named_modules():
        if len(list(m.children())) == 0: m.register_forward_hook(hook(acts)(n))
    crit = losses.P2P_Loss(level, 1., 0., 0.)
    out = mod(x.cuda()); loss = crit(out, tgt.cuda()); loss.backward(); torch.cuda.synchronize()
    print('==', impl, 'B', B, 'loss', loss.item(), loss_r.item())
    for n in acts:
        if n not in acts_r: continue
        a, b = acts[n], acts_r[n]
        print('  act %-28s rel %.2e' % (n, ((a - b).norm() / b.norm().clamp_min(1e-20)).item()))
    for k, p in mod.named_parameters():
        a, b = p.grad.cpu().flatten().double(), gr[k].flatten().double()
        print('  grad %-32s cos %.6f rel %.2e' % (k, (a @ b / (a.norm() * b.norm())).item(), ((a - b).norm() / b.norm()).item()))
