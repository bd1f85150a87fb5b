"""Data-parallel training harness: one process per GPU, gradients averaged with NCCL.

The reference is single-device (run.py:713); SURVEY 8e: the path shards over independent samples,
BatchNorm statistics stay per replica, and the only exchange is one gradient all-reduce per step
(4.63 M fp32 = 18.5 MB for ico2ico).  Gradients are gathered into a few flat bucket buffers; a
bucket's all-reduce is launched from autograd hooks as soon as its last gradient of the step has
arrived, so the exchange overlaps the rest of backward.

Two sources feed the buckets: (1) autograd's post-accumulate hooks (module-by-module path: gradients appear layer by layer);
(2) the fused chains (geniconet_b200/fused.py), whose ONE autograd Function only returns its gradients when the whole chain's
backward has run -- they hand every residual block's gradients to `early_grads()` as soon as the block's wgrad kernels are
enqueued, so a bucket can start its all-reduce (NCCL enqueues on its own stream, also under CUDA-graph capture) while the
remaining blocks are still being differentiated.
"""
import os

import torch
import torch.distributed as dist


class GradBuckets:
    """Flat gradient buckets for the data-parallel all-reduce.

    Autograd is left to ASSIGN each parameter's gradient (p.grad is None at the start of a step), which costs no kernel; with
    pre-allocated bucket views as p.grad every parameter paid one accumulate (add) launch per step -- 78 launches at ico2ico.
    Gradients reach their bucket slot in one of three ways: (1) the fused chains' wgrad kernels write weight gradients straight
    into the slot (grad_dest; sibling weights whose gradient one kernel produces are laid out side by side, `adjacent`) -- the
    copies that used to gather them cost 115 us per step, more than the exchange; (2) the chains hand the remaining gradients
    (BatchNorm weight / bias, conv bias) over from inside their backward (early_grads), copied with one multi-tensor launch when
    the bucket is complete; (3) everything else arrives through autograd's post-accumulate hooks.  When the last gradient of a
    bucket is there its all-reduce starts, overlapping the rest of backward.  finish_bucket(i) waits for bucket i only and points
    p.grad at the reduced views (one optimizer per bucket can then update it while later buckets are still in flight); finish()
    does so for all.  With world_size 1 nothing is copied or exchanged at all.
    """

    def __init__(self, params, world_size, bucket_bytes=None, process_group=None, adjacent=()):
        import os
        if bucket_bytes is None:
            bucket_bytes = int(float(os.environ.get('GIN_DP_BUCKET_MB', '8')) * (1 << 20))
        self.world = int(world_size)
        self.group = process_group
        # reverse registration order ~ the order gradients become ready in backward.  `adjacent`: tuples of parameters whose
        # gradients one kernel produces as ONE tensor (fused.weight_pairs): they are laid out side by side, in the given order, and
        # never split across buckets, so that kernel can write into the bucket directly (grad_dest).
        order = [p for p in params if p.requires_grad][::-1]
        group_of = {}
        for grp in adjacent:
            if all(q.requires_grad for q in grp):
                for q in grp:
                    group_of[id(q)] = tuple(grp)
        self.params, seen, units = [], set(), []
        for p in order:
            if id(p) in seen:
                continue
            unit = list(group_of.get(id(p), (p,)))
            for q in unit:
                seen.add(id(q))
            units.append(unit)
            self.params += unit
        self.buckets = []          # (flat tensor, [params], [views])
        cur, cur_n = [], 0
        for unit in units:
            cur += unit
            cur_n += sum(q.numel() for q in unit) * 4
            if cur_n >= bucket_bytes:
                self._seal(cur)
                cur, cur_n = [], 0
        if cur:
            self._seal(cur)
        self._slot = {}            # id(param) -> (bucket index, element offset)
        for bi, (flat, ps, views) in enumerate(self.buckets):
            off = 0
            for q in ps:
                self._slot[id(q)] = (bi, off)
                off += q.numel()
        self._pending = [0] * len(self.buckets)
        self._handles = []
        self._early = set()          # id(param) of gradients that arrived through early_grads() this step
        self._early_grad = {}        # id(param) -> that gradient, until its bucket is complete
        self._bucket_of = {}
        for bi, (_, ps, _) in enumerate(self.buckets):
            for p in ps:
                self._bucket_of[p] = bi
                if self.world > 1:
                    p.register_post_accumulate_grad_hook(self._on_grad)
        self.reset()

    def _seal(self, ps):
        n = sum(p.numel() for p in ps)
        flat = torch.zeros(n, dtype=ps[0].dtype, device=ps[0].device)
        views, off = [], 0
        for p in ps:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.buckets.append((flat, list(ps), views))

    def reset(self):
        """Drop last step's gradients (autograd will assign, not accumulate) and re-arm the bucket counters; call before backward."""
        for bi, (_, ps, _) in enumerate(self.buckets):
            for p in ps:
                p.grad = None
            self._pending[bi] = len(ps)
        self._handles = []
        self._early = set()
        self._early_grad = {}
        if self.world > 1 and os.environ.get('GIN_DP_EARLY', '1') != '0':
            from . import fused
            fused.set_grad_sink(self.early_grads)
            fused.set_grad_dest(self.grad_dest)

    def _on_grad(self, p):
        if id(p) in self._early:                         # already counted (and copied) when the fused chain produced it
            return
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def grad_dest(self, params):
        """The bucket storage of `params` laid end to end, if that is how the bucket holds them (adjacent, in this order)."""
        if self.world <= 1 or not params or any(id(q) not in self._slot for q in params):
            return None
        bi, off0 = self._slot[id(params[0])]
        off = off0
        for q in params:
            if self._slot[id(q)] != (bi, off):
                return None
            off += q.numel()
        return self.buckets[bi][0][off0:off]

    def early_grads(self, pairs):
        """(parameter, gradient) pairs whose gradient kernels are enqueued on the current stream although autograd has not
        assigned p.grad yet.  Each gradient is copied into its bucket slot now; a bucket whose last gradient arrives this way
        starts its all-reduce immediately.  Parameters of other models (not in any bucket) are ignored."""
        if self.world <= 1:
            return
        touched = set()
        for p, g in pairs:
            bi = self._bucket_of.get(p)
            if bi is None or id(p) in self._early or g is None:
                continue
            self._early.add(id(p))
            slot_bi, slot_off = self._slot[id(p)]
            if g.data_ptr() != self.buckets[slot_bi][0].data_ptr() + 4 * slot_off:      # else: the kernel wrote it in place
                self._early_grad[id(p)] = g.view_as(p)  # copied into the bucket in ONE launch when the bucket is complete
            self._pending[bi] -= 1
            touched.add(bi)
        for bi in touched:
            if self._pending[bi] == 0:
                self._launch(bi)

    def _launch(self, bi):
        flat, ps, views = self.buckets[bi]
        have = [(v, p.grad) for v, p in zip(views, ps) if p.grad is not None and id(p) not in self._early]
        have += [(v, self._early_grad.pop(id(p))) for v, p in zip(views, ps) if id(p) in self._early_grad]
        if len([p for p in ps if p.grad is not None or id(p) in self._early]) < len(ps):
            # parameters that got no gradient this step contribute zeros (their slots may hold last step's averages)
            missing = [v for v, p in zip(views, ps) if p.grad is None and id(p) not in self._early]
            torch._foreach_zero_(missing)
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if os.environ.get('GIN_DP_NOCOMM') == '1':           # diagnosis only: bucket copies without the exchange
            return
        if dist.get_backend(self.group) == 'nccl':
            h = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._handles.append((h, None, bi))
        else:
            h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._handles.append((h, flat, bi))

    def bucket_params(self):
        """The parameters of every bucket, in bucket order (one optimizer per bucket can then update a bucket as soon as ITS
        all-reduce is done, see finish_bucket)."""
        return [list(ps) for _, ps, _ in self.buckets]

    def _flush(self):
        for bi, n in enumerate(self._pending):      # buckets with parameters that got no gradient this step
            if n > 0:
                self._pending[bi] = 0
                self._launch(bi)

    def finish_bucket(self, bi):
        """After backward: wait for bucket `bi` only and expose its averaged gradients as p.grad.  Calling it bucket by bucket
        with an optimizer per bucket hides the last (smallest, latest) all-reduce behind the update of the earlier buckets."""
        if self.world <= 1:
            return
        self._flush()
        for k, (h, flat, hb) in enumerate(self._handles):
            if hb == bi and h is not None:
                h.wait()
                if flat is not None:
                    flat.div_(self.world)
                self._handles[k] = (None, None, hb)
        _, ps, views = self.buckets[bi]
        for p, v in zip(ps, views):
            p.grad = v

    def finish(self):
        """Wait for every bucket and expose the averaged gradients as p.grad; call after backward and before optimizer.step."""
        if self.world <= 1:
            return
        for bi in range(len(self.buckets)):
            self.finish_bucket(bi)
        self._handles = []
        from . import fused
        fused.set_grad_sink(None)
        fused.set_grad_dest(None)

    def total_bytes(self):
        return sum(f.numel() * 4 for f, _, _ in self.buckets)


def shard_sample_ids(step, rank, world, batch):
    """Disjoint synthetic sample ids per rank: step*world*B + rank*B + i."""
    first = step * world * batch + rank * batch
    return list(range(first, first + batch))


def broadcast_parameters(model, src=0, process_group=None):
    """Rank `src`'s parameters and buffers everywhere.  Written in place under no_grad (not through `.data`), so each
    Parameter's version counter moves and weight-derived caches (IcoConvS2S._packed_weights) see the change."""
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t, src=src, group=process_group)
