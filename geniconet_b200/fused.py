"""Fused execution of the residual encoder / decoder bodies of models.py.

The module-by-module path (ico_conv.py + torch BatchNorm2d / ReLU / add, exactly what the unmodified models.py runs) moves every
conv output through five torch kernels before the next conv sees it.  Here a whole chain of

    stem   IcoConvS2S(3 -> C) + BatchNorm2d + ReLU                                   (models.py:104-111)
    block  BasicIcoS2SDownBlock / BasicIcoS2SUpBlock                                 (models.py:22-62)

runs inside ONE autograd Function whose internal activations exist only as the bf16 operand copies the tcgen05 kernels gather
from (plus the fp32 conv outputs BatchNorm needs):

  * the two sibling convolutions of a block (conv00 / conv10 read the same input) are one GEMM with concatenated output
    channels -- forward, wgrad and (concatenated along K, which also adds the two input gradients for free) dgrad;
  * BatchNorm statistics, normalise + ReLU + residual add + bf16 cast are gin_bn_stats / gin_bn_act_fwd, their backward
    gin_bn_act_bwd writes the bf16 gradient copy dgrad and wgrad read (include/geniconet_b200.h);
  * an Up block upsamples once (upsample00 / upsample10 are the same parameter-free map), from the fp32 coarse map straight
    into the bf16 operand copy of the fine level.

Parameters stay where models.py puts them (conv00.weight, icobn00.weight, ...), so state dicts are unchanged; BatchNorm running
statistics are updated as torch does.  Training mode only (batch statistics); in eval mode the modules run one by one.
A conv bias that feeds a BatchNorm has a mathematically zero gradient; the fused path returns exactly zero for it.
"""
import torch

from . import _lib
from .ico_conv import IcoConvS2S, IcoUpsampleS2S, get_plan, _stream, pixel_strides

L = _lib.lib


def _P(level):
    return 10 * 4 ** level


def _is_block(m):
    return hasattr(m, 'conv00') and hasattr(m, 'icobn10')


def chain_supported(mods):
    """True when `mods` is [IcoConvS2S(3->C), BatchNorm2d, ReLU]? + residual blocks with tensor-core channel widths."""
    mods = list(mods)
    i = 0
    if len(mods) >= 3 and isinstance(mods[0], IcoConvS2S) and isinstance(mods[1], torch.nn.BatchNorm2d) and isinstance(mods[2], torch.nn.ReLU):
        if mods[0].in_features != 3 or mods[0].stride != 1 or mods[0].out_features % 64 or mods[0].bias is None:
            return False
        i = 3
    if i == len(mods):
        return False
    if i == 3 and _is_block(mods[3]) and not mods[3]._down:
        return False                                  # an Up block needs the fp32 map of its input, the stem keeps only the bf16 copy
    for m in mods[i:]:
        if not _is_block(m):
            return False
        for c in (m.conv00, m.conv01, m.conv10):
            if c.in_features % 64 or c.out_features % 64 or c.bias is None or 256 % (c.out_features // 8):
                return False
        for b in (m.icobn00, m.icobn01, m.icobn10):
            if not isinstance(b, torch.nn.BatchNorm2d) or not b.affine or not b.track_running_stats or b.momentum is None:
                return False
    return True


def chain_params(mods):
    out = []
    for m in mods:
        if isinstance(m, IcoConvS2S):
            out += [m.weight, m.bias]
        elif isinstance(m, torch.nn.BatchNorm2d):
            out += [m.weight, m.bias]
        elif _is_block(m):
            for c, b in ((m.conv00, m.icobn00), (m.conv01, m.icobn01), (m.conv10, m.icobn10)):
                out += [c.weight, c.bias, b.weight, b.bias]
    return out


# ------------------------------------------------------------------------------------------------ thin wrappers over the C ABI
def _empty(n, dtype, dev):
    return torch.empty(n, dtype=dtype, device=dev)


def _pack(w, cin, cout):
    packed = _empty(L.gin_hexconv_packed_bytes(cin, cout), torch.uint8, w.device)
    _lib.check(L.gin_hexconv_pack_weights(w.data_ptr(), packed.data_ptr(), cin, cout, _stream()), 'gin_hexconv_pack_weights')
    return packed


def _pack_bf16(w0, w1, cin):
    """bf16 tile images of [w0; w1] (concatenated output channels) without materialising the concatenation."""
    c0, c1 = w0.shape[0], (w1.shape[0] if w1 is not None else 0)
    packed = _empty(L.gin_hexconv_packed_bytes(cin, c0 + c1), torch.uint8, w0.device)
    _lib.check(L.gin_hexconv_pack_weights_bf16(w0.data_ptr(), c0, w1.data_ptr() if w1 is not None else None, c1, packed.data_ptr(), cin, _stream()),
               'gin_hexconv_pack_weights_bf16')
    return packed


def _pack_many(jobs):
    """jobs: list of (w0, w1 or None, cin) -> list of packed blobs (16-bit tile images only), ONE launch for all of them."""
    import ctypes
    n = len(jobs)
    if n == 0:
        return []
    if n > 24:
        return _pack_many(jobs[:24]) + _pack_many(jobs[24:])
    dev = jobs[0][0].device
    outs = [_empty(L.gin_hexconv_packed_bytes(cin, w0.shape[0] + (w1.shape[0] if w1 is not None else 0)), torch.uint8, dev) for w0, w1, cin in jobs]
    P, I = ctypes.c_void_p * n, ctypes.c_int * n
    w0s = P(*[w0.data_ptr() for w0, _, _ in jobs])
    w1s = P(*[(w1.data_ptr() if w1 is not None else None) for _, w1, _ in jobs])
    c0 = I(*[w0.shape[0] for w0, _, _ in jobs])
    c1 = I(*[(w1.shape[0] if w1 is not None else 0) for _, w1, _ in jobs])
    pk = P(*[o.data_ptr() for o in outs])
    ci = I(*[cin for _, _, cin in jobs])
    _lib.check(L.gin_hexconv_pack_weights_bf16_multi(n, w0s, c0, w1s, c1, pk, ci, _stream()), 'gin_hexconv_pack_weights_bf16_multi')
    return outs


def _conv_fwd(plan, xb, packed, bias, B, cin, cout, p_out, bias1=None):
    """Returns (y fp32 [B*p_out][cout], per-CTA BatchNorm partial sums [nparts][2][cout] or None).  With `bias1` the output
    channels are two sibling convolutions side by side: `bias` belongs to the first half, `bias1` to the second."""
    import ctypes
    parts = _empty(L.gin_hexconv_stats_ws_bytes(cout) // 4, torch.float32, xb.device)
    n = ctypes.c_int(0)
    # The output is only ever read by the BatchNorm kernels: written as fp16 (GIN_Y_FP16=0: fp32) it costs half the bytes of its
    # one write and three reads; the statistics come from the fp32 accumulators either way.
    y = _empty((B * p_out, cout), torch.float16 if _Y16 else torch.float32, xb.device)
    rc = L.gin_hexconv_fwd_bf16_stats2(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), packed.data_ptr(), bias.data_ptr(),
                                       bias1.data_ptr() if bias1 is not None else None, cout // 2 if bias1 is not None else 0,
                                       y.data_ptr(), 1 if _Y16 else 0, B, cin, cout, parts.data_ptr(), ctypes.addressof(n), _stream())
    if rc == 0:
        return y, ((parts, n.value) if n.value > 0 else None)
    if rc != _lib.ERR_UNSUPPORTED:
        _lib.check(rc, 'gin_hexconv_fwd_bf16_stats2')
    # a kernel generation without the second bias pointer / the fp16 epilogue
    y = _empty((B * p_out, cout), torch.float32, xb.device)
    if bias1 is not None:
        bias = torch.cat((bias, bias1), 0)
    _lib.check(L.gin_hexconv_fwd_bf16_stats(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), packed.data_ptr(), bias.data_ptr(), y.data_ptr(), B, cin, cout,
                                            parts.data_ptr(), ctypes.addressof(n), _stream()), 'gin_hexconv_fwd_bf16_stats')
    return y, ((parts, n.value) if n.value > 0 else None)


def _conv_dgrad(plan, dyb, packed, B, cin, cout, p_in):
    dx = _empty((B * p_in, cin), torch.float32, dyb.device)
    _lib.check(L.gin_hexconv_dgrad_bf16(plan.host_ptr, plan.dev_ptr, dyb.data_ptr(), packed.data_ptr(), dx.data_ptr(), B, cin, cout, _stream()),
               'gin_hexconv_dgrad_bf16')
    return dx


def _conv_wgrad(plan, xb, dyb, B, cin, cout, side=None, out=None):
    """dW [cout][cin][7] (written into `out` when given: a data-parallel bucket slot).  With `side` (a CUDA stream) the wgrad is
    issued there: nothing on the rest of the backward pass depends on it, so the tensor-bound wgrad overlaps the memory-bound
    BatchNorm backward kernels of the main stream."""
    if side is None:
        dW = out if out is not None else _empty((cout, cin, 7), torch.float32, xb.device)
        ws = _empty(L.gin_hexconv_wgrad_ws_bytes(cin, cout), torch.uint8, xb.device)
        _lib.check(L.gin_hexconv_wgrad_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), dyb.data_ptr(), None, dW.data_ptr(), None, ws.data_ptr(),
                                            B, cin, cout, _stream()), 'gin_hexconv_wgrad_bf16')
        return dW
    main = torch.cuda.current_stream()
    side.wait_stream(main)                               # dyb (and xb) are complete on the main stream
    with torch.cuda.stream(side):
        dW = _empty((cout, cin, 7), torch.float32, xb.device)
        ws = _empty(L.gin_hexconv_wgrad_ws_bytes(cin, cout), torch.uint8, xb.device)
        _lib.check(L.gin_hexconv_wgrad_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), dyb.data_ptr(), None, dW.data_ptr(), None, ws.data_ptr(),
                                            B, cin, cout, side.cuda_stream), 'gin_hexconv_wgrad_bf16')
    xb.record_stream(side)                               # their memory must not be recycled on the main stream while the side stream reads it
    dyb.record_stream(side)
    return dW


import os as _os
_Y16 = _os.environ.get('GIN_Y_FP16', '1') != '0'
_side_streams = {}
_grad_sink = None
_grad_dest = None


def set_grad_sink(fn):
    """fn(list of (parameter, gradient)) is called from inside a chain's backward as soon as a block's gradients are enqueued
    (data-parallel training: geniconet_b200.dp.GradBuckets starts the all-reduce of a complete bucket right there); None
    removes it.  The chain still returns the same gradients to autograd afterwards."""
    global _grad_sink
    _grad_sink = fn


def set_grad_dest(fn):
    """fn(tuple of parameters) -> a contiguous fp32 tensor that IS the gradient storage of those parameters laid end to end (a
    slot of a data-parallel bucket), or None.  The chain's wgrad kernels then write straight into it (no bucket copy)."""
    global _grad_dest
    _grad_dest = fn


def weight_pairs(model):
    """The (conv00.weight, conv10.weight) pairs of every residual block: their gradients come out of ONE wgrad GEMM as one
    [2*Cout, Cin, 7] tensor, so a data-parallel bucket that keeps each pair adjacent can receive it in place."""
    return [(m.conv00.weight, m.conv10.weight) for m in model.modules() if _is_block(m)]


def _dest(params, shape):
    if _grad_dest is None:
        return None
    t = _grad_dest(params)
    return t.view(shape) if t is not None and t.numel() == int(torch.Size(shape).numel()) else None


def _wgrad_stream(dev):
    import os
    # measured at I5 / B=36: no gain (4.30 vs 4.26 ms per step) -- the persistent conv kernels leave no room for co-resident CTAs --
    # so the side stream is opt-in (GIN_WGRAD_STREAM=1)
    if os.environ.get('GIN_WGRAD_STREAM', '0') != '1':
        return None
    key = (dev.type, dev.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def _bn_stats(y, col0, ld, rows, C, bn, parts=None):
    stat = _empty(4 * C, torch.float32, y.device)
    if parts is not None:         # the conv epilogue already summed y and y^2 per CTA
        pbuf, n = parts
        _lib.check(L.gin_bn_stats_from_parts(pbuf.data_ptr() + 4 * col0, n, ld, rows, C, bn.weight.data_ptr(), bn.bias.data_ptr(), float(bn.eps),
                                             float(bn.momentum), bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                             bn.num_batches_tracked.data_ptr(), stat.data_ptr(), _stream()), 'gin_bn_stats_from_parts')
        return stat
    assert y.dtype == torch.float32, 'gin_bn_stats reads fp32 maps (fp16 conv outputs always come with epilogue statistics)'
    ws = _empty(L.gin_bn_ws_bytes(C), torch.uint8, y.device)
    _lib.check(L.gin_bn_stats(y.data_ptr() + 4 * col0, ld, rows, C, bn.weight.data_ptr(), bn.bias.data_ptr(), float(bn.eps), float(bn.momentum),
                              bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), stat.data_ptr(), ws.data_ptr(),
                              _stream()), 'gin_bn_stats')
    return stat


_DUAL = bool(L.gin_forward_operand_is_fp16())        # forward operands fp16: every activation copy also exists in bf16 for wgrad


def _bn_act(y1, col1, ld1, stat1, y2, col2, ld2, stat2, B, level, C, want_b=True, want_f=False, want_w=True):
    """Returns (out_b: the next conv's forward operand copy, out_f: fp32 map, out_w: its bf16 twin = wgrad operand and ReLU mask;
    out_w is out_b itself when the forward format is bf16).  16-bit copies are allocated as torch.bfloat16 containers."""
    dev = y1.device
    out_b = _empty(((B * _P(level) + 2 * B), C), torch.bfloat16, dev) if want_b else None
    out_f = _empty((B * _P(level), C), torch.float32, dev) if want_f else None
    out_w = _empty(((B * _P(level) + 2 * B), C), torch.bfloat16, dev) if (want_w and (_DUAL or not want_b)) else None
    assert y2 is None or y2.dtype == y1.dtype
    _lib.check(L.gin_bn_act_fwd(y1.data_ptr() + y1.element_size() * col1, ld1, stat1.data_ptr(),
                                (y2.data_ptr() + y2.element_size() * col2) if y2 is not None else None, ld2, stat2.data_ptr() if stat2 is not None else None,
                                1 if y1.dtype == torch.float16 else 0, 1,
                                out_b.data_ptr() if want_b else None, out_f.data_ptr() if want_f else None,
                                out_w.data_ptr() if out_w is not None else None, B, level, C, _stream()), 'gin_bn_act_fwd')
    if want_w and out_w is None:
        out_w = out_b
    return out_b, out_f, out_w


# The BatchNorm backward kernels re-evaluate the ReLU mask from y and the BatchNorm constants instead of reading it from the
# activation copy (include/geniconet_b200.h: relu_from_y): 2 B / element less in each pass, step 3.40 -> 3.27 ms on the same box
# (profiles/r02_ab_mask_*.json).  GIN_BN_MASK_FROM_Y=0 reads the mask.
_MASK_FROM_Y = 0 if _os.environ.get('GIN_BN_MASK_FROM_Y', '1') == '0' else 1


def _bn_bwd(dout, mask_b, y, col0, ld, stat, B, level, C, dy_b=None, dy_b_col=0, ldo=0, want_f=False, mask_from_y=None):
    """Returns (bstat [4C]: dbeta, dgamma, ..., dy_f or None); writes the bf16 gradient copy into dy_b[:, dy_b_col:dy_b_col+C].
    mask_b must be the activation copy _bn_act produced from (y, stat) with ReLU (it may then be re-evaluated instead of read)."""
    dev = dout.device
    bstat = _empty(4 * C, torch.float32, dev)
    ws = _empty(L.gin_bn_ws_bytes(C), torch.uint8, dev)
    dy_f = _empty((B * _P(level), C), torch.float32, dev) if want_f else None
    _lib.check(L.gin_bn_act_bwd(dout.data_ptr(), C, mask_b.data_ptr() if mask_b is not None else None, y.data_ptr() + y.element_size() * col0, ld,
                                1 if y.dtype == torch.float16 else 0, stat.data_ptr(), bstat.data_ptr(), (dy_b.data_ptr() + 2 * dy_b_col) if dy_b is not None else None, ldo,
                                dy_f.data_ptr() if want_f else None, C, ws.data_ptr(), B, level, C,
                                _MASK_FROM_Y if mask_from_y is None else int(mask_from_y), _stream()), 'gin_bn_act_bwd')
    return bstat, dy_f


# ------------------------------------------------------------------------------------------------ the chain
class _Chain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mods, *params):
        dev = x.device
        B = x.shape[0]
        saved = []                # per stage: dict of what backward needs
        i = 0
        act_b, level, C = None, None, None
        ctx.x_needs_grad = x.requires_grad
        if isinstance(mods[0], IcoConvS2S):                                  # ---- stem: xyz conv (fp32, warp-level kernels) + BN + ReLU
            conv, bn = mods[0], mods[1]
            level, C = conv.subdivisions, conv.out_features
            plan = get_plan(_lib.PLAN_HEXCONV, level, 1, conv.corner_mode, dev)
            packed = _pack(conv.weight.detach().contiguous(), 3, C)
            xs, sb, sp, sc = pixel_strides(x)
            import ctypes
            y = _empty((B * _P(level), C), torch.float16 if _Y16 else torch.float32, dev)
            sparts = _empty(L.gin_hexconv_narrow_stats_ws_bytes(C) // 4, torch.float32, dev)
            npart = ctypes.c_int(0)
            rc = L.gin_hexconv_fwd_narrow_stats(plan.host_ptr, plan.dev_ptr, xs.data_ptr(), sb, sp, sc, packed.data_ptr(), conv.bias.data_ptr(),
                                                y.data_ptr(), 1 if _Y16 else 0, B, 3, C, sparts.data_ptr(), ctypes.addressof(npart), _stream())
            if rc == 0 and npart.value > 0:          # output statistics came out of the conv epilogue
                stat = _bn_stats(y, 0, C, B * _P(level), C, bn, (sparts, npart.value))
            else:
                if rc != 0 and rc != _lib.ERR_UNSUPPORTED:
                    _lib.check(rc, 'gin_hexconv_fwd_narrow_stats')
                y = _empty((B * _P(level), C), torch.float32, dev)
                _lib.check(L.gin_hexconv_fwd(plan.host_ptr, plan.dev_ptr, xs.data_ptr(), sb, sp, sc, packed.data_ptr(), conv.bias.data_ptr(),
                                             y.data_ptr(), B, 3, C, _lib.IMPL_AUTO, _stream()), 'gin_hexconv_fwd')
                stat = _bn_stats(y, 0, C, B * _P(level), C, bn)
            act_b, _, act_w = _bn_act(y, 0, C, stat, None, 0, 0, None, B, level, C)
            saved.append(dict(kind='stem', plan=plan, xs=xs, strides=(sb, sp, sc), y=y, stat=stat, out_b=act_w, C=C, level=level, packed=packed))
            i = 3
        else:                                                                # ---- chain starts on an fp32 map: make the operand copy
            blk = mods[0]
            level, C = blk.conv00.subdivisions - (0 if blk._down else 1), blk.conv00.in_features
            plan = get_plan(_lib.PLAN_HEXCONV, level, 1, blk.conv00.corner_mode, dev)
            from .ico_conv import as_channels_last, cast_bf16
            act_b = cast_bf16(as_channels_last(x), plan, 0, level)
            act_w = cast_bf16(as_channels_last(x), plan, 2, level) if _DUAL else act_b
            saved.append(dict(kind='input', level=level, C=C))
        ctx.in_shape = tuple(x.shape)
        act_f = None
        if saved[0]['kind'] == 'input':
            from .ico_conv import as_channels_last as _acl
            act_f = _acl(x).permute(0, 2, 3, 1).reshape(-1, C)
        last = len(mods) - 1
        out_f = None
        # every weight of the chain's tensor-core convolutions is packed by ONE launch
        jobs = []
        for blk in mods[i:]:
            jobs.append((blk.conv00.weight.detach().contiguous(), blk.conv10.weight.detach().contiguous(), blk.conv00.in_features))
            jobs.append((blk.conv01.weight.detach().contiguous(), None, blk.conv01.in_features))
        packs = _pack_many(jobs)
        for j in range(i, len(mods)):
            blk = mods[j]
            cm = blk.conv00.corner_mode
            cin, cout = blk.conv00.in_features, blk.conv00.out_features
            st = dict(kind='down' if blk._down else 'up', cin=cin, cout=cout, in_level=level, blk=blk)
            if blk._down:
                lvl = level - 1
                plan_a = get_plan(_lib.PLAN_HEXCONV, level, 2, cm, dev)
                a_b, a_w = act_b, act_w
            else:
                lvl = level + 1
                up_plan = get_plan(_lib.PLAN_UPSAMPLE, level, 1, blk.upsample00.corner_mode, dev)
                plan_a = get_plan(_lib.PLAN_HEXCONV, lvl, 1, cm, dev)
                # upsample straight into the operand copy of the fine level, from the fp32 coarse map (one rounding)
                a_b = _empty((B * _P(lvl) + 2 * B, cin), torch.bfloat16, dev)
                a_w = _empty((B * _P(lvl) + 2 * B, cin), torch.bfloat16, dev) if _DUAL else a_b
                _lib.check(L.gin_upsample_bf16(up_plan.host_ptr, up_plan.dev_ptr, act_f.data_ptr(), 1, a_b.data_ptr(), a_w.data_ptr() if _DUAL else None,
                                               B, cin, _stream()), 'gin_upsample_bf16')
                st['up_plan'] = up_plan
            plan_b = get_plan(_lib.PLAN_HEXCONV, lvl, 1, cm, dev)
            rows = B * _P(lvl)
            pk_cat, pk01 = packs[2 * (j - i)], packs[2 * (j - i) + 1]
            ycat, pcat = _conv_fwd(plan_a, a_b, pk_cat, blk.conv00.bias.detach(), B, cin, 2 * cout, _P(lvl), bias1=blk.conv10.bias.detach())   # [rows][conv00 | conv10]
            if pcat is not None:                 # both sibling BatchNorms from the same epilogue sums: one launch
                stat00, stat10 = _empty(4 * cout, torch.float32, dev), _empty(4 * cout, torch.float32, dev)
                bA, bB = blk.icobn00, blk.icobn10
                _lib.check(L.gin_bn_stats_from_parts2(pcat[0].data_ptr(), pcat[1], 2 * cout, rows, cout,
                                                      0, bA.weight.data_ptr(), bA.bias.data_ptr(), float(bA.eps), float(bA.momentum), bA.running_mean.data_ptr(),
                                                      bA.running_var.data_ptr(), bA.num_batches_tracked.data_ptr(), stat00.data_ptr(),
                                                      cout, bB.weight.data_ptr(), bB.bias.data_ptr(), float(bB.eps), float(bB.momentum), bB.running_mean.data_ptr(),
                                                      bB.running_var.data_ptr(), bB.num_batches_tracked.data_ptr(), stat10.data_ptr(), _stream()),
                           'gin_bn_stats_from_parts2')
            else:
                stat00 = _bn_stats(ycat, 0, 2 * cout, rows, cout, blk.icobn00, pcat)
                stat10 = _bn_stats(ycat, cout, 2 * cout, rows, cout, blk.icobn10, pcat)
            h_b, _, h_w = _bn_act(ycat, 0, 2 * cout, stat00, None, 0, 0, None, B, lvl, cout)
            y01, p01 = _conv_fwd(plan_b, h_b, pk01, blk.conv01.bias.detach(), B, cout, cout, _P(lvl))
            stat01 = _bn_stats(y01, 0, cout, rows, cout, blk.icobn01, p01)
            is_last = j == last
            keep_f = is_last or not mods[j + 1]._down         # an Up block upsamples from the fp32 map (one rounding instead of two)
            # the bf16 copy of the block output is the ReLU mask of the backward, so it is always produced
            # the forward copy of the last block's output has no consumer; its bf16 twin is still the ReLU mask of the backward
            out_b, out_f, out_w = _bn_act(y01, 0, cout, stat01, ycat, cout, 2 * cout, stat10, B, lvl, cout, want_b=not is_last or not _DUAL, want_f=keep_f)
            act_f = out_f
            # saved for backward: the bf16 twins (wgrad operands / ReLU masks); the forward-format copies die with the forward pass
            st.update(plan_a=plan_a, plan_b=plan_b, a_b=a_w, pk_cat=pk_cat, pk01=pk01, ycat=ycat, y01=y01, h_b=h_w, out_b=out_w,
                      stat00=stat00, stat01=stat01, stat10=stat10, level=lvl)
            saved.append(st)
            act_b, act_w, level, C = out_b, out_w, lvl, cout
        ctx.saved, ctx.B, ctx.nparams = saved, B, len(params)
        ctx.mods = mods
        n = 2 ** level
        return out_f.view(B, 5 * n, 2 * n, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, d_out):
        B = ctx.B
        saved = ctx.saved
        from .ico_conv import as_channels_last
        d = as_channels_last(d_out).permute(0, 2, 3, 1).reshape(-1, d_out.shape[1])          # fp32 [rows][C], no copy when channels-last
        if not d.is_contiguous():
            d = d.contiguous()
        grads = []                # filled back to front, reversed at the end
        dx = None
        side = _wgrad_stream(d.device)
        plist = chain_params(ctx.mods)          # the Parameter objects, in the order of the gradients this function returns
        zmax = max([st['cout'] for st in saved if st['kind'] in ('down', 'up')] + [1])
        zeros = torch.zeros(zmax, dtype=torch.float32, device=d.device)     # the (exactly zero) gradients of conv biases that feed a BatchNorm
        for st in reversed(saved):
            n_before = len(grads)
            if st['kind'] in ('down', 'up'):
                cin, cout, lvl = st['cin'], st['cout'], st['level']
                dev = d.device
                dycat_b = torch.empty((B * _P(lvl) + 2 * B, 2 * cout), dtype=torch.bfloat16, device=dev)
                dy01_b = torch.empty((B * _P(lvl) + 2 * B, cout), dtype=torch.bfloat16, device=dev)
                # bn01 and bn10 see the same g = d * (out > 0): one pass pair for both
                bs01 = torch.empty(4 * cout, dtype=torch.float32, device=dev)
                bs10 = torch.empty(4 * cout, dtype=torch.float32, device=dev)
                ws = torch.empty(L.gin_bn_pair_ws_bytes(cout), dtype=torch.uint8, device=dev)
                assert st['y01'].dtype == st['ycat'].dtype
                _lib.check(L.gin_bn_act_bwd_pair(d.data_ptr(), cout, st['out_b'].data_ptr(), st['y01'].data_ptr(), cout, st['stat01'].data_ptr(),
                                                 bs01.data_ptr(), dy01_b.data_ptr(), cout, st['ycat'].data_ptr() + st['ycat'].element_size() * cout, 2 * cout,
                                                 st['stat10'].data_ptr(), bs10.data_ptr(), dycat_b.data_ptr() + 2 * cout, 2 * cout,
                                                 1 if st['ycat'].dtype == torch.float16 else 0, ws.data_ptr(), B, lvl, cout, _MASK_FROM_Y, _stream()), 'gin_bn_act_bwd_pair')
                blk = st['blk']
                dW01 = _conv_wgrad(st['plan_b'], st['h_b'], dy01_b, B, cout, cout, side,
                                   out=_dest((blk.conv01.weight,), (cout, cout, 7)) if side is None else None)
                d_h = _conv_dgrad(st['plan_b'], dy01_b, st['pk01'], B, cout, cout, _P(lvl))
                bs00, _ = _bn_bwd(d_h, st['h_b'], st['ycat'], 0, 2 * cout, st['stat00'], B, lvl, cout, dycat_b, 0, 2 * cout)
                dWcat = _conv_wgrad(st['plan_a'], st['a_b'], dycat_b, B, cin, 2 * cout, side,
                                    out=_dest((blk.conv00.weight, blk.conv10.weight), (2 * cout, cin, 7)) if side is None else None)
                p_in = _P(lvl + 1) if st['kind'] == 'down' else _P(lvl)
                d_in = _conv_dgrad(st['plan_a'], dycat_b, st['pk_cat'], B, cin, 2 * cout, p_in)
                if st['kind'] == 'up':
                    upp = st['up_plan']
                    d_prev = torch.empty((B * _P(st['in_level']), cin), dtype=torch.float32, device=dev)
                    _lib.check(L.gin_upsample_bwd(upp.host_ptr, upp.dev_ptr, d_in.data_ptr(), d_prev.data_ptr(), B, cin, _stream()), 'gin_upsample_bwd')
                    d_in = d_prev
                zero_b = zeros[:cout]
                # parameter order of chain_params: conv00 (w, b, gamma, beta), conv01 (...), conv10 (...); appended reversed
                grads += [bs10[:cout], bs10[cout:2 * cout], zero_b, dWcat[cout:],
                          bs01[:cout], bs01[cout:2 * cout], zero_b, dW01,
                          bs00[:cout], bs00[cout:2 * cout], zero_b, dWcat[:cout]]
                d = d_in
            elif st['kind'] == 'stem':
                C, lvl = st['C'], st['level']
                bs, dy_f = _bn_bwd(d, st['out_b'], st['y'], 0, C, st['stat'], B, lvl, C, want_f=True)
                plan = st['plan']
                dW = torch.empty((C, 3, 7), dtype=torch.float32, device=d.device)
                db = torch.empty((C,), dtype=torch.float32, device=d.device)
                ws = torch.empty(L.gin_hexconv_wgrad_ws_bytes(3, C), dtype=torch.uint8, device=d.device)
                sb, sp, sc = st['strides']
                _lib.check(L.gin_hexconv_wgrad(plan.host_ptr, plan.dev_ptr, st['xs'].data_ptr(), sb, sp, sc, dy_f.data_ptr(), dW.data_ptr(), db.data_ptr(),
                                               ws.data_ptr(), B, 3, C, _lib.IMPL_AUTO, _stream()), 'gin_hexconv_wgrad')
                grads += [bs[:C], bs[C:2 * C], db, dW]
                if ctx.x_needs_grad:                     # gradient of the xyz input (exact fp32 gather-GEMM: 3 output channels)
                    n = 2 ** lvl
                    dxs = torch.empty((B * _P(lvl), 3), dtype=torch.float32, device=d.device)
                    _lib.check(L.gin_hexconv_dgrad(plan.host_ptr, plan.dev_ptr, dy_f.data_ptr(), st['packed'].data_ptr(), dxs.data_ptr(), B, 3, C,
                                                   _lib.IMPL_SIMT, _stream()), 'gin_hexconv_dgrad')
                    dx = dxs.view(B, 5 * n, 2 * n, 3).permute(0, 3, 1, 2)
            else:   # 'input': gradient of the fp32 map the chain started from
                n = 2 ** st['level']
                dx = d.view(B, 5 * n, 2 * n, st['C']).permute(0, 3, 1, 2)
            if _grad_sink is not None and side is None and len(grads) > n_before:
                # grads[k] (appended back to front) belongs to plist[nparams - 1 - k]
                _grad_sink([(plist[ctx.nparams - 1 - k], grads[k]) for k in range(n_before, len(grads))])
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)      # the weight gradients are consumed (optimizer, all-reduce) on the main stream
        grads.reverse()
        assert len(grads) == ctx.nparams
        return (dx, None) + tuple(grads)


def run_chain(x, mods):
    """Forward of the module list `mods` (see chain_supported) on the fused path; x fp32 CUDA, returns the fp32 output map."""
    mods = list(mods)
    return _Chain.apply(x, mods, *chain_params(mods))


# ------------------------------------------------------------------ decoder head: Conv2d(64, 3, 1) -> Tanh (models.py:151-154)
def head_supported(head, x):
    """head = nn.Sequential(Conv2d 1x1 64->3 with bias, Tanh) without hooks, x a dense fp32 CUDA map."""
    mods = list(head)
    if len(mods) != 2 or not isinstance(mods[0], torch.nn.Conv2d) or not isinstance(mods[1], torch.nn.Tanh):
        return False
    c = mods[0]
    if (c.kernel_size != (1, 1) or c.stride != (1, 1) or c.padding != (0, 0) or c.groups != 1 or c.bias is None
            or c.in_channels != 64 or c.out_channels != 3 or c.weight.dtype != torch.float32):
        return False
    if any(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks for m in [head] + mods):
        return False
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 64


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        from .ico_conv import as_channels_last
        B, C, H, W = x.shape
        xs = as_channels_last(x)                              # [B*P][64] in memory; no copy after a fused chain
        w = weight.reshape(3, C).contiguous()
        y = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib.gin_head_fwd(xs.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H * W, C, 3, _stream()), 'gin_head_fwd')
        ctx.save_for_backward(xs, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, w, y = ctx.saved_tensors
        B, C, H, W = xs.shape
        dy = dy.contiguous()
        dx = torch.empty_strided(xs.shape, xs.stride(), dtype=torch.float32, device=xs.device)
        dw = torch.empty((3, C, 1, 1), dtype=torch.float32, device=xs.device)
        db = torch.empty(3, dtype=torch.float32, device=xs.device)
        ws = _empty(_lib.lib.gin_head_ws_bytes(), torch.uint8, xs.device)
        _lib.check(_lib.lib.gin_head_bwd(xs.data_ptr(), w.data_ptr(), y.data_ptr(), dy.data_ptr(), dx.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                         ws.data_ptr(), B, H * W, C, 3, _stream()), 'gin_head_bwd')
        return dx, dw, db


def run_head(head, x):
    return _Head.apply(x, head[0].weight, head[0].bias)
