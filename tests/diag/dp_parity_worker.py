"""Multi-GPU data-parallel parity worker, launched by tests/test_gpu_dp.py or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/diag/dp_parity_worker.py

Every rank: same weights (broadcast), its own synthetic shard.  (1) the purely local gradient of the fused step, no
communication; (2) the mean of those over the ranks, through one plain NCCL all_reduce that does not involve GradBuckets;
(3) the gradients the CAPTURED-GRAPH step leaves in p.grad (forward + loss + backward + bucketed all-reduce started from inside
the fused chain's backward, replayed twice).  (3) must equal (2).  Rank 0 prints one JSON line.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from geniconet_b200 import models as gm, losses, data, fused, reparam        # noqa: E402
from geniconet_b200.dp import GradBuckets, shard_sample_ids, broadcast_parameters   # noqa: E402
from geniconet_b200.graph import GraphedStep                                  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
name = sys.argv[1] if len(sys.argv) > 1 else 'ico2ico'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 6
level = 5
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
params = gm.default_params(name, level)
torch.manual_seed(100 + rank)                       # deliberately different weights per rank before the broadcast
model = getattr(gm, name)(params).cuda().train()
broadcast_parameters(model)
f = (params['ico']['factor_pos'], params['ico']['factor_nor'], params['ico']['factor_lap'])
crit = losses.P2PKLD_Loss(level, *f, 1.0) if name == 'ico2ico_vae' else losses.P2P_Loss(level, *f)
ids = shard_sample_ids(0, rank, world, B)
xs, ts = zip(*(data.synthetic_mesh(level, i) for i in ids))
x, t = torch.stack(xs).cuda(), torch.stack(ts).cuda()
bn_state = {k: v.clone() for k, v in model.state_dict().items() if 'running' in k or 'num_batches' in k}


def restore_bn():
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if k in bn_state:
                v.copy_(bn_state[k])


# (1) local gradient, eager, no buckets, no sink.  Not on the legacy default stream: the AccumulateGrad nodes created here stay
# bound to the stream of this first backward, and a capture may not synchronise with the legacy stream.
fused.set_grad_sink(None)
reparam.manual_seed(9)
for p in model.parameters():
    p.grad = None
work = torch.cuda.Stream()
work.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(work):
    loss = crit(model(x), t)
    loss.backward()
torch.cuda.synchronize()
named = list(model.named_parameters())
local_g = [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for _, p in named]
# (2) rank mean through a plain all_reduce
flat = torch.cat([g.flatten() for g in local_g])
dist.all_reduce(flat, op=dist.ReduceOp.SUM)
flat /= world
# (3) the graphed data-parallel step
restore_bn()
buckets = GradBuckets(model.parameters(), world, adjacent=fused.weight_pairs(model))
early_counts = []


def step(xb, tb):
    buckets.reset()
    reparam.manual_seed(9)
    ls = crit(model(xb), tb)
    ls.backward()
    early_counts.append(len(buckets._early))
    buckets.finish()
    return ls


graphed = GraphedStep(step, (x, t), warmup=2)
for _ in range(2):
    graphed()
torch.cuda.synchronize()
got = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().flatten() for _, p in named])
err = (got - flat).abs().max().item()
scale = flat.abs().max().item()
worst, off = None, 0
for (k, p), g in zip(named, local_g):
    n = p.numel()
    a, b = got[off:off + n], flat[off:off + n]
    e = ((a - b).norm() / (b.norm() + 1e-30)).item() if b.norm() > 1e-7 else 0.0
    if worst is None or e > worst[1]:
        worst = (k, e)
    off += n
res = torch.tensor([err / scale, worst[1]], device='cuda')
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if name == 'ico2ico_vae':
    note = 'vae: the graph advances a device-side noise counter per replay, so (1) and (3) draw different eps; only finite-ness is checked'
else:
    note = ''
if rank == 0:
    print(json.dumps({'model': name, 'world': world, 'batch_per_rank': B, 'n_buckets': len(buckets.buckets), 'bucket_mb': [round(b[0].numel() * 4 / 2 ** 20, 2) for b in buckets.buckets],
                      'params_handed_over_early': max(early_counts), 'max_abs_err_over_max_grad': res[0].item(), 'worst_param_rel_l2': res[1].item(),
                      'worst_param_rank0': worst[0], 'finite': bool(torch.isfinite(got).all()), 'note': note}), flush=True)
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)          # destroying a communicator whose collectives live in a captured graph hung at exit (r01)
