"""GPU timeline of the replayed training-step graph (torch.profiler / CUPTI): per-kernel device time INSIDE the graph, idle gaps
between kernels, and what overlaps the NCCL kernels.

    python tools/trace_step.py [--model ico2ico] [--level 5] [--batch 36] [--replays 3] [--out gpurun_out/trace_step.json]
    python -m torch.distributed.run --nproc-per-node 2 ... tools/trace_step.py          # data parallel: rank 0 reports
"""
import argparse
import collections
import json
import os
import re
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument('--model', default='ico2ico')
ap.add_argument('--level', type=int, default=5)
ap.add_argument('--batch', type=int, default=None)
ap.add_argument('--replays', type=int, default=3)
ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'trace_step.json'))
args = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
if world > 1:
    spare = int(os.environ.get('GIN_DP_SPARE_SMS', '0'))
    if spare > 0:
        os.environ.setdefault('GIN_SMS', str(148 - spare))
        os.environ.setdefault('NCCL_MAX_CTAS', str(spare))
from geniconet_b200 import models as gm, losses, data                        # noqa: E402
from geniconet_b200.dp import GradBuckets, shard_sample_ids, broadcast_parameters   # noqa: E402
from geniconet_b200.graph import GraphedStep                                  # noqa: E402

torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
B = args.batch or (36 if args.level == 5 else 16)
params = gm.default_params(args.model, args.level)
torch.manual_seed(0)
model = getattr(gm, args.model)(params).cuda()
if world > 1:
    broadcast_parameters(model)
f = (params['ico']['factor_pos'], params['ico']['factor_nor'], params['ico']['factor_lap'])
crit = losses.P2PKLD_Loss(args.level, *f, 1.0) if args.model == 'ico2ico_vae' else losses.P2P_Loss(args.level, *f)
from geniconet_b200 import fused as _fused                                   # noqa: E402
buckets = GradBuckets(model.parameters(), world, adjacent=_fused.weight_pairs(model))
if os.environ.get('GIN_BENCH_OPTIMIZER', 'gin') == 'gin':
    from geniconet_b200.optim import Adam as _Adam                           # noqa: E402
    make_opt = lambda ps: _Adam(ps, lr=1e-4)                                 # noqa: E731
else:
    make_opt = lambda ps: torch.optim.Adam(ps, lr=1e-4, fused=True, capturable=True)      # noqa: E731
opts = [make_opt(ps) for ps in (buckets.bucket_params() if world > 1 else [list(model.parameters())])]
ids = shard_sample_ids(0, rank, world, min(B, 4))
xs, ts = zip(*(data.synthetic_mesh(args.level, i) for i in ids))
x = torch.stack(xs).repeat((B + 3) // 4, 1, 1, 1)[:B].cuda()
t = torch.stack(ts).repeat((B + 3) // 4, 1, 1)[:B].cuda()


def step(xb, tb):
    buckets.reset()
    loss = crit(model(xb), tb)
    loss.backward()
    for bi, o in enumerate(opts):
        buckets.finish_bucket(bi)
        o.step()
    buckets.finish()
    return loss


graphed = GraphedStep(step, (x, t), warmup=3)
for _ in range(3):
    graphed()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(args.replays):
        graphed()
    torch.cuda.synchronize()
if rank == 0:
    tmp = args.out + '.chrome.json'
    prof.export_chrome_trace(tmp)
    ev = [e for e in json.load(open(tmp))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
    os.remove(tmp)
    ev.sort(key=lambda e: e['ts'])
    # split into replays at the largest gaps
    t0 = ev[0]['ts']
    span = ev[-1]['ts'] + ev[-1]['dur'] - t0
    per = span / args.replays
    last = [e for e in ev if e['ts'] >= t0 + (args.replays - 1) * per - 1]        # the last replay
    def short(n):
        n = re.sub(r'\(.*', '', n)
        return re.sub(r'^void ', '', n)[:70]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in last:
        a = agg[short(e['name'])]
        a[0] += 1
        a[1] += e['dur']
    start, end = last[0]['ts'], max(e['ts'] + e['dur'] for e in last)
    busy, cur_end = 0.0, start
    for e in last:                                       # union of the busy intervals over all streams
        s, en = e['ts'], e['ts'] + e['dur']
        if en <= cur_end:
            continue
        busy += en - max(s, cur_end)
        cur_end = en
    nccl = [e for e in last if 'nccl' in e['name'].lower()]
    rep = {'model': args.model, 'level': args.level, 'batch': B, 'world': world, 'replay_us': end - start, 'busy_us': busy, 'idle_us': end - start - busy,
           'kernels': len(last), 'sum_of_kernel_us': sum(e['dur'] for e in last),
           'by_kernel': sorted(([k, v[0], round(v[1], 1)] for k, v in agg.items()), key=lambda r: -r[2]),
           'nccl': [{'start_us': round(e['ts'] - start, 1), 'dur_us': round(e['dur'], 1), 'before_end_us': round(end - (e['ts'] + e['dur']), 1),
                     'overlapped_by': [short(o['name']) for o in last if o is not e and o['ts'] < e['ts'] + e['dur'] and o['ts'] + o['dur'] > e['ts']][:8]} for e in nccl]}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rep, open(args.out, 'w'), indent=1)
    print('replay %.1f us: busy %.1f, idle %.1f over %d kernels (sum of kernel times %.1f us)' % (rep['replay_us'], busy, rep['idle_us'], len(last), rep['sum_of_kernel_us']))
    for k, n, us in rep['by_kernel'][:40]:
        print('%9.1f us %4d x  %5.1f%%  %s' % (us, n, 100 * us / rep['replay_us'], k))
    for r in rep['nccl']:
        print('NCCL at +%.0f us for %.0f us (ends %.0f us before the step ends), concurrent with: %s' % (r['start_us'], r['dur_us'], r['before_end_us'], r['overlapped_by']))
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)
