"""Does the bf16-operand tcgen05 path train like the fp32 reference?  (VERDICT r01 item 1b / ADVICE: evidence, not tolerance fitting.)

    python tests/diag/precision_study.py [--steps 200] [--cond 100] [--out gpurun_out/precision_study.json]

For ico2ico and ico2ico_vae (the reference's loss factors, run.py:689-696) at I5, batch 36:

  1. loss curves: the SAME init, data order and reparameterisation noise trained `--steps` Adam steps twice -- fused tcgen05
     path (bf16 operands, fp32 accumulate) and module-wise exact-fp32 CUDA-core path (impl='simt'); per-step total loss and
     per-term losses are recorded.
  2. gradient parity on a CONDITIONED state: the weights of the bf16 run after `--cond` steps, one held-out batch, gradients
     from (a) the fused tcgen05 path, (b) the module-wise tcgen05 path, (c) the fp32 CUDA-core path, all against (d) the CPU
     oracle (oracle/models_ref.py, fp32, same noise).  The same comparison at the random init for reference.

Test infrastructure (uses oracle/): nothing in the product imports this.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import oracle_models as om                                                   # noqa: E402
from oracle import models_ref, synth_ref                                     # noqa: E402
from geniconet_b200 import models as gm, losses, reparam                     # noqa: E402
from geniconet_b200.ico_conv import set_impl, clear_caches                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=200)
ap.add_argument('--cond', type=int, default=100)
ap.add_argument('--batch', type=int, default=36)
ap.add_argument('--level', type=int, default=5)
ap.add_argument('--lr', type=float, default=3e-4, help='constant; the reference cycles 1e-9..1e-3 (run.py:448-449)')
ap.add_argument('--pool', type=int, default=4, help='distinct batches cycled through')
ap.add_argument('--models', default='ico2ico,ico2ico_vae')
ap.add_argument('--no-oracle', action='store_true')
ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'precision_study.json'))
args = ap.parse_args()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
level, B = args.level, args.batch


def make_model(name, fused, impl, state=None):
    gm.set_fused(fused, fused)
    mod = om.fill_params_deterministic(getattr(gm, name)(gm.default_params(name, level)))
    if state is not None:
        mod.load_state_dict(state)
    return set_impl(mod.cuda().train(), impl)


def criterion(name):
    f = models_ref.LOSS_FACTORS[name]
    return losses.P2PKLD_Loss(level, *f, 1.0) if name == 'ico2ico_vae' else losses.P2P_Loss(level, *f)


def train(name, fused, impl, batches, steps, snap_at):
    mod = make_model(name, fused, impl)
    crit = criterion(name)
    opt = torch.optim.Adam(mod.parameters(), lr=args.lr)
    reparam.manual_seed(1234)
    curve, snap = [], None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        if i == snap_at:
            snap = {k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}
        x, t = batches[i % len(batches)]
        opt.zero_grad(set_to_none=True)
        loss = crit(mod(x), t)
        loss.backward()
        opt.step()
        curve.append([float(v) for v in crit.get_last_losses()])
    torch.cuda.synchronize()
    clear_caches()
    return curve, snap, (time.perf_counter() - t0) / steps


def grads_cuda(name, fused, impl, state, x, t, seed):
    mod = make_model(name, fused, impl, state)
    crit = criterion(name)
    box = {}
    if name == 'ico2ico_vae':
        orig = gm._reparameterize

        def capture(mu, lv):
            z, eps = reparam.reparameterize(mu, lv, seed=seed, offset=1, return_eps=True)
            box['eps'] = eps.detach().cpu()
            return z
        gm._reparameterize = capture
    try:
        loss = crit(mod(x), t)
    finally:
        if name == 'ico2ico_vae':
            gm._reparameterize = orig
    loss.backward()
    torch.cuda.synchronize()
    clear_caches()
    return loss.item(), [float(v) for v in crit.get_last_losses()], {k: p.grad.detach().double().cpu().flatten() for k, p in mod.named_parameters()}, box.get('eps')


def grads_oracle(name, state, x, t, eps):
    ref = models_ref.build(name, level)
    ref.load_state_dict(state)
    ref.train()
    torch.set_num_threads(os.cpu_count() or 1)
    out = ref(x, eps=eps) if name == 'ico2ico_vae' else ref(x)
    loss = models_ref.training_loss(name, level, out, t)
    loss.backward()
    return loss.item(), {k: p.grad.detach().double().flatten() for k, p in ref.named_parameters()}


def compare(ga, gb):
    rows = {}
    for k, b in gb.items():
        if b.norm() < 1e-6:            # conv biases in front of a BatchNorm: the true gradient is zero
            continue
        a = ga[k]
        rows[k] = {'cos': (a @ b / (a.norm() * b.norm() + 1e-300)).item(), 'norm_ratio': (a.norm() / b.norm()).item()}
    cs = [r['cos'] for r in rows.values()]
    return {'min_cos': min(cs), 'mean_cos': sum(cs) / len(cs), 'worst': min(rows, key=lambda k: rows[k]['cos']), 'per_param': rows}


report = {'config': {'level': level, 'batch': B, 'steps': args.steps, 'conditioned_after': args.cond, 'lr': args.lr, 'pool_batches': args.pool,
                     'gpu': torch.cuda.get_device_name(0)}}
for name in args.models.split(','):
    batches = [tuple(v.cuda() for v in synth_ref.synthetic_batch(level, 1000 + b * B, B)) for b in range(args.pool)]
    held = synth_ref.synthetic_batch(level, 50000, B)
    rep = {}
    curve_b, snap, sec_b = train(name, True, 'auto', batches, args.steps, args.cond)
    curve_f, _, sec_f = train(name, False, 'simt', batches, args.steps, -1)
    tot = lambda c: [r[-1] for r in c]
    rel = [abs(a - b) / max(abs(b), 1e-12) for a, b in zip(tot(curve_b), tot(curve_f))]
    rep['curves'] = {'bf16_fused': curve_b, 'fp32_simt': curve_f, 'columns': 'criterion.get_last_losses()', 'sec_per_step': {'bf16_fused': sec_b, 'fp32_simt': sec_f},
                     'max_rel_diff_total': max(rel), 'mean_rel_diff_total': sum(rel) / len(rel),
                     'final_total': {'bf16_fused': tot(curve_b)[-1], 'fp32_simt': tot(curve_f)[-1]},
                     'mean_last10': {'bf16_fused': sum(tot(curve_b)[-10:]) / 10, 'fp32_simt': sum(tot(curve_f)[-10:]) / 10}}
    print('%s: %d steps  bf16-fused %.4f -> %.4f (%.1f ms/step)   fp32 %.4f -> %.4f (%.1f ms/step)   max |rel diff| %.3g' % (
        name, args.steps, tot(curve_b)[0], tot(curve_b)[-1], sec_b * 1e3, tot(curve_f)[0], tot(curve_f)[-1], sec_f * 1e3, max(rel)), flush=True)
    init_state = {k: v.detach().clone() for k, v in om.fill_params_deterministic(getattr(gm, name)(gm.default_params(name, level))).state_dict().items()}
    for tag, state in (('random_init', init_state), ('conditioned', snap)):
        xh, th = held[0].cuda(), held[1].cuda()
        res = {}
        eps = None
        for path, fused, impl in (('bf16_fused', True, 'auto'), ('bf16_modulewise', False, 'auto'), ('fp32_simt', False, 'simt')):
            l, parts, g, e = grads_cuda(name, fused, impl, state, xh, th, seed=77)
            res[path] = (l, parts, g)
            eps = e if e is not None else eps
        ent = {'loss': {p: res[p][0] for p in res}, 'last_losses': {p: res[p][1] for p in res}}
        if not args.no_oracle:
            lo, go = grads_oracle(name, state, held[0], held[1], eps)
            ent['loss']['oracle_cpu_fp32'] = lo
            ent['vs_oracle'] = {p: compare(res[p][2], go) for p in res}
        ent['vs_fp32_simt'] = {p: compare(res[p][2], res['fp32_simt'][2]) for p in ('bf16_fused', 'bf16_modulewise')}
        rep[tag] = ent
        base = 'vs_oracle' if not args.no_oracle else 'vs_fp32_simt'
        print('%s / %s: loss %s' % (name, tag, {k: round(v, 6) for k, v in ent['loss'].items()}))
        for p, c in ent[base].items():
            print('    %-16s gradient cosine vs %s: min %.5f (%s)  mean %.5f' % (p, base[3:], c['min_cos'], c['worst'], c['mean_cos']), flush=True)
    report[name] = rep
    del batches
    torch.cuda.empty_cache()
gm.set_fused(True, True)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
with open(args.out, 'w') as fh:
    json.dump(report, fh)
print('wrote', args.out)
