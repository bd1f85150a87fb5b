"""Host-side mirror of ``icocnn.ico_conv`` over the C ABI.

Same constructor signatures and tensor contracts as the reference's call sites
(/root/reference/models.py:13-15, 25-33, 45-55, 104, 165, 269, 279):

    IcoConvS2S(in_features, out_features, stride, bias, subdivisions, corner_mode=...)
        forward: [B, Cin, 5n, 2n] float32 -> [B, Cout, 5n/stride, 2n/stride]
    IcoUpsampleS2S(in_features, subdivisions, corner_mode)
        forward: [B, C, 5n, 2n] -> [B, C, 10n, 4n]

`subdivisions` is the INPUT level.  Outputs are ordinary 4-D float32 tensors with logical NCHW
shape and channels-last strides, so the stock BatchNorm2d/ReLU/add of models.py:37-39 run on
them unchanged and without a transpose.  There is no CPU implementation: tensors that are
not on a CUDA device raise.
"""
import math
import weakref

import numpy as np
import torch

from . import _lib

_PLANS = {}


class Plan:
    """Host blob + device copy of one plan; both stay alive for the life of the process."""

    def __init__(self, kind, level, stride, corner_mode, device):
        self.host = _lib.plan_blob(kind, level, stride, corner_mode)
        self.dev = torch.from_numpy(self.host).to(device)
        self.host_ptr = self.host.ctypes.data
        self.dev_ptr = self.dev.data_ptr()


def get_plan(kind, level, stride, corner_mode, device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('geniconet_b200 has no CPU path: expected a CUDA device, got %s' % device)
    if device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    key = (kind, level, stride, corner_mode, device.index)
    p = _PLANS.get(key)
    if p is None:
        p = _PLANS[key] = Plan(kind, level, stride, corner_mode, device)
    return p


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda_f32(x, what):
    if not x.is_cuda:
        raise RuntimeError('%s: geniconet_b200 has no CPU path (tensor is on %s)' % (what, x.device))
    if x.dtype != torch.float32:
        raise TypeError('%s: expected float32, got %s' % (what, x.dtype))


def pixel_strides(x):
    """(tensor, sb, sp, sc) element strides of a [B,C,H,W] map seen as (sample, pixel, channel)."""
    B, C, H, W = x.shape
    s0, s1, s2, s3 = x.stride()
    if H * W > 1 and s2 != W * s3 and H > 1:
        x = x.contiguous(memory_format=torch.channels_last)
        s0, s1, s2, s3 = x.stride()
    return x, s0, s3, s1


def as_channels_last(x):
    """[B,C,H,W] with channels-last strides (no copy when it already is)."""
    B, C, H, W = x.shape
    if x.stride() == (H * W * C, 1, W * C, C):
        return x
    return x.contiguous(memory_format=torch.channels_last) if C > 1 and H * W > 1 else \
        x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def _new_map(B, C, H, W, device):
    return torch.empty((B, H, W, C), dtype=torch.float32, device=device).permute(0, 3, 1, 2)


class _CastCache:
    """Short-lived sharing of a derived tensor between two consumers of the SAME input tensor: the bf16 operand copy for the
    sibling convolutions of a residual block (conv00 / conv10, models.py:25-33) and the upsampled map for upsample00 /
    upsample10 (models.py:59-60).  An entry is keyed by the identity, version counter AND storage address of the input and by
    the version counter of the cached output, so an in-place change of either between the two uses is a miss; entries whose
    input tensor has died are purged on every access, `take=True` removes an entry on its first hit (the second consumer is the
    last one), and clear_caches() drops everything (e.g. at the end of a step)."""

    def __init__(self, size=3):
        self.size, self.items = size, []

    def _purge(self):
        self.items = [it for it in self.items if it[0]() is not None]

    def get(self, t, key, take=False):
        self._purge()
        for i, (ref, ver, ptr, k, out, out_ver) in enumerate(self.items):
            if ref() is t and ver == t._version and ptr == t.data_ptr() and k == key and out._version == out_ver:
                if take:
                    del self.items[i]
                return out
        return None

    def put(self, t, key, out):
        self._purge()
        self.items.append((weakref.ref(t), t._version, t.data_ptr(), key, out, out._version))
        del self.items[:-self.size]

    def clear(self):
        self.items = []


_cast_cache = _CastCache()
_upsample_cache = _CastCache(size=2)


def clear_caches():
    """Drop the sibling-sharing caches (they hold at most a few tensors of the last forward)."""
    _cast_cache.clear()
    _upsample_cache.clear()


def cast_bf16(x_cl, plan, which, level, use_cache=False, colsum=None):
    """16-bit operand copy [B*P + 2B, C] of a channels-last fp32 map (pixels, then the per-sample pole means), held in a
    torch.bfloat16 container: `which` = 0 forward operand (fp16 by default), 1 gradient dy (bf16), 2 forward tensor in bf16 for
    wgrad (include/geniconet_b200.h: gin_cast_bf16).  `colsum` [C] (optional) receives the per-channel sums over all pixels
    from the same pass (the conv bias gradient)."""
    B, C = x_cl.shape[0], x_cl.shape[1]
    key = (plan.dev_ptr, which)
    if use_cache:
        hit = _cast_cache.get(x_cl, key, take=True)
        if hit is not None:
            return hit
    out = torch.empty((_lib.lib.gin_cast_bf16_bytes(B, level, C) // 2,), dtype=torch.bfloat16, device=x_cl.device)
    if colsum is not None:
        ws = torch.empty((_lib.lib.gin_cast_bf16_colsum_ws_bytes(C),), dtype=torch.uint8, device=x_cl.device)
        _lib.check(_lib.lib.gin_cast_bf16_colsum(plan.host_ptr, plan.dev_ptr, which, x_cl.data_ptr(), out.data_ptr(), colsum.data_ptr(),
                                                 ws.data_ptr(), B, C, _stream()), 'gin_cast_bf16_colsum')
        return out
    _lib.check(_lib.lib.gin_cast_bf16(plan.host_ptr, plan.dev_ptr, which, x_cl.data_ptr(), out.data_ptr(), B, C, _stream()), 'gin_cast_bf16')
    if use_cache:
        _cast_cache.put(x_cl, key, out)
    return out


class _HexConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        _require_cuda_f32(x, 'IcoConvS2S')
        B, C, H, W = x.shape
        n = mod._n
        if C != mod.in_features or H != 5 * n or W != 2 * n:
            raise ValueError('IcoConvS2S(level %d): expected [B,%d,%d,%d], got %s'
                             % (mod.subdivisions, mod.in_features, 5 * n, 2 * n, tuple(x.shape)))
        plan = get_plan(_lib.PLAN_HEXCONV, mod.subdivisions, mod.stride, mod.corner_mode, x.device)
        packed = mod._packed_weights(weight)
        Ho, Wo = H // mod.stride, W // mod.stride
        y = _new_map(B, mod.out_features, Ho, Wo, x.device)
        tc = mod.uses_tensor_cores()
        ctx.mod, ctx.plan, ctx.has_bias, ctx.tc, ctx.in_shape = mod, plan, bias is not None, tc, (B, C, H, W)
        bias_ptr = bias.data_ptr() if bias is not None else None
        if tc:
            xs = as_channels_last(x)
            xb = cast_bf16(xs, plan, 0, mod.subdivisions, use_cache=True) if B > 0 else xs.new_empty(0, dtype=torch.bfloat16)
            if B > 0:
                _lib.check(_lib.lib.gin_hexconv_fwd_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), packed.data_ptr(), bias_ptr, y.data_ptr(),
                                                         B, mod.in_features, mod.out_features, _stream()), 'gin_hexconv_fwd_bf16')
            # wgrad needs x in bf16 (one format per MMA): with fp16 forward operands that is a second 16-bit copy (`which` = 2)
            xw = xb
            if B > 0 and _lib.lib.gin_forward_operand_is_fp16() and (ctx.needs_input_grad[1] or torch.is_grad_enabled()):
                xw = cast_bf16(xs, plan, 2, mod.subdivisions, use_cache=True)
            ctx.save_for_backward(xw, packed)           # a 16-bit copy is all wgrad needs: half the saved-activation bytes
        else:
            xs, sb, sp, sc = pixel_strides(x)
            if B > 0:
                _lib.check(_lib.lib.gin_hexconv_fwd(plan.host_ptr, plan.dev_ptr, xs.data_ptr(), sb, sp, sc, packed.data_ptr(), bias_ptr,
                                                    y.data_ptr(), B, mod.in_features, mod.out_features, mod.impl, _stream()), 'gin_hexconv_fwd')
            ctx.strides = (sb, sp, sc)
            ctx.save_for_backward(xs, packed)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, packed = ctx.saved_tensors
        mod, plan = ctx.mod, ctx.plan
        B, C, H, W = ctx.in_shape
        dev = dy.device
        dy = as_channels_last(dy)
        dx = dW = db = None
        st = _stream()
        need_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        if B == 0:
            return (torch.zeros(ctx.in_shape, device=dev) if ctx.needs_input_grad[0] else None,
                    torch.zeros(mod.out_features, mod.in_features, 7, device=dev),
                    torch.zeros(mod.out_features, device=dev) if ctx.has_bias else None, None)
        if need_w:
            dW = torch.empty((mod.out_features, mod.in_features, 7), dtype=torch.float32, device=dev)
            db = torch.empty((mod.out_features,), dtype=torch.float32, device=dev) if ctx.has_bias else None
            ws = torch.empty((_lib.lib.gin_hexconv_wgrad_ws_bytes(mod.in_features, mod.out_features),), dtype=torch.uint8, device=dev)
        if ctx.needs_input_grad[0]:
            dx = _new_map(B, mod.in_features, H, W, dev)
        if ctx.tc:
            level_out = mod.subdivisions - (1 if mod.stride == 2 else 0)
            fuse_db = db is not None and 256 % (mod.out_features // 8) == 0
            dyb = cast_bf16(dy, plan, 1, level_out, colsum=db if fuse_db else None)   # one cast serves dgrad, wgrad and db
            if fuse_db:
                db_arg = None
            else:
                db_arg = db
            if dx is not None:
                _lib.check(_lib.lib.gin_hexconv_dgrad_bf16(plan.host_ptr, plan.dev_ptr, dyb.data_ptr(), packed.data_ptr(), dx.data_ptr(),
                                                           B, mod.in_features, mod.out_features, st), 'gin_hexconv_dgrad_bf16')
            if need_w:
                _lib.check(_lib.lib.gin_hexconv_wgrad_bf16(plan.host_ptr, plan.dev_ptr, xs.data_ptr(), dyb.data_ptr(), dy.data_ptr(),
                                                           dW.data_ptr(), db_arg.data_ptr() if db_arg is not None else None, ws.data_ptr(),
                                                           B, mod.in_features, mod.out_features, st), 'gin_hexconv_wgrad_bf16')
        else:
            if dx is not None:
                _lib.check(_lib.lib.gin_hexconv_dgrad(plan.host_ptr, plan.dev_ptr, dy.data_ptr(), packed.data_ptr(), dx.data_ptr(),
                                                      B, mod.in_features, mod.out_features, _lib.IMPL_SIMT, st), 'gin_hexconv_dgrad')
            if need_w:
                sb, sp, sc = ctx.strides
                _lib.check(_lib.lib.gin_hexconv_wgrad(plan.host_ptr, plan.dev_ptr, xs.data_ptr(), sb, sp, sc, dy.data_ptr(),
                                                      dW.data_ptr(), db.data_ptr() if db is not None else None, ws.data_ptr(),
                                                      B, mod.in_features, mod.out_features, mod.impl, st), 'gin_hexconv_wgrad')
        return dx, dW, db, None


class IcoConvS2S(torch.nn.Module):
    """Hex-masked 3x3 convolution over the 5 icosahedral charts with the chart padding fused in.

    Parameters: ``weight [Cout, Cin, 7]`` (taps: centre, N, S, W, E, NE, SW in lattice terms, see
    gin_host.cpp kTap) and ``bias [Cout]``.
    """

    def __init__(self, in_features, out_features, stride=1, bias=True, subdivisions=0, corner_mode='zeros', impl='auto'):
        super().__init__()
        if stride not in (1, 2):
            raise ValueError('stride must be 1 or 2')
        if stride == 2 and subdivisions < 1:
            raise ValueError('stride 2 needs subdivisions >= 1')
        if corner_mode not in _lib.CORNER:
            raise ValueError('corner_mode must be zeros or average, got %r' % (corner_mode,))
        self.in_features, self.out_features = int(in_features), int(out_features)
        self.stride, self.subdivisions, self.corner_mode = int(stride), int(subdivisions), corner_mode
        self.impl = {'auto': _lib.IMPL_AUTO, 'simt': _lib.IMPL_SIMT, 'tc': _lib.IMPL_TC}[impl]
        self._n = 2 ** self.subdivisions
        self.weight = torch.nn.Parameter(torch.empty(self.out_features, self.in_features, 7))
        self.bias = torch.nn.Parameter(torch.empty(self.out_features)) if bias else None
        bound = 1.0 / math.sqrt(self.in_features * 7)
        torch.nn.init.uniform_(self.weight, -bound, bound)
        if bias:
            torch.nn.init.uniform_(self.bias, -bound, bound)
        self._packed = None
        self._packed_key = None

    def _packed_weights(self, weight):
        key = (weight.data_ptr(), weight._version, weight.device)
        # under CUDA-graph capture the packing kernel must be part of the graph: the optimizer updates the weights in place
        # on every replay and no Python runs then
        if self._packed is None or self._packed_key != key or torch.cuda.is_current_stream_capturing():
            nbytes = _lib.lib.gin_hexconv_packed_bytes(self.in_features, self.out_features)
            # a fresh buffer per weight version: a pending backward may still hold the previous one
            self._packed = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device)
            w = weight.detach()
            if not w.is_contiguous():
                w = w.contiguous()
            _lib.check(_lib.lib.gin_hexconv_pack_weights(w.data_ptr(), self._packed.data_ptr(), self.in_features,
                                                         self.out_features, _stream()), 'gin_hexconv_pack_weights')
            self._packed_key = key
        return self._packed

    def uses_tensor_cores(self):
        """tcgen05 path (bf16 operands, fp32 accumulate) when the conv is a dense contraction: both channel counts % 64 == 0."""
        wide = self.in_features % 64 == 0 and self.out_features % 64 == 0
        if self.impl == _lib.IMPL_TC and not wide:
            raise ValueError('impl="tc" needs in_features and out_features to be multiples of 64')
        return wide and self.impl != _lib.IMPL_SIMT

    def forward(self, x):
        return _HexConvFn.apply(x, self.weight, self.bias, self)

    def extra_repr(self):
        return '%d -> %d, stride=%d, level=%d, corner_mode=%s' % (self.in_features, self.out_features, self.stride,
                                                                   self.subdivisions, self.corner_mode)


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod):
        _require_cuda_f32(x, 'IcoUpsampleS2S')
        B, C, H, W = x.shape
        n = 2 ** mod.subdivisions
        if H != 5 * n or W != 2 * n:
            raise ValueError('IcoUpsampleS2S(level %d): expected [B,C,%d,%d], got %s' % (mod.subdivisions, 5 * n, 2 * n, tuple(x.shape)))
        if C % 4:
            raise ValueError('IcoUpsampleS2S: channel count must be a multiple of 4, got %d' % C)
        plan = get_plan(_lib.PLAN_UPSAMPLE, mod.subdivisions, 1, mod.corner_mode, x.device)
        xs = as_channels_last(x)
        y = _new_map(B, C, 2 * H, 2 * W, x.device)
        _lib.check(_lib.lib.gin_upsample_fwd(plan.host_ptr, plan.dev_ptr, xs.data_ptr(), y.data_ptr(), B, C, _stream()), 'gin_upsample_fwd')
        ctx.plan, ctx.shape = plan, (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, C, H, W = ctx.shape
        dy = as_channels_last(dy)
        dx = _new_map(B, C, H, W, dy.device)
        _lib.check(_lib.lib.gin_upsample_bwd(ctx.plan.host_ptr, ctx.plan.dev_ptr, dy.data_ptr(), dx.data_ptr(), B, C, _stream()),
                   'gin_upsample_bwd')
        return dx, None


class IcoUpsampleS2S(torch.nn.Module):
    """Level s -> s+1: coarse vertices copy, edge midpoints average their two endpoints (parameter free)."""

    def __init__(self, in_features, subdivisions, corner_mode='zeros'):
        super().__init__()
        if corner_mode not in _lib.CORNER:
            raise ValueError('corner_mode must be zeros or average, got %r' % (corner_mode,))
        self.in_features, self.subdivisions, self.corner_mode = int(in_features), int(subdivisions), corner_mode

    def forward(self, x):
        # The layer has no parameters, so two IcoUpsampleS2S of the same geometry applied to the SAME tensor
        # (upsample00 / upsample10 of a BasicIcoS2SUpBlock, models.py:59-60) produce the same map: the second call returns
        # the first one's output (autograd adds the two incoming gradients before the single backward).
        key = (self.subdivisions, self.corner_mode)
        want_grad = x.requires_grad and torch.is_grad_enabled()
        hit = _upsample_cache.get(x, (key, want_grad), take=True)
        if hit is not None:
            return hit
        y = _UpsampleFn.apply(x, self)
        _upsample_cache.put(x, (key, want_grad), y)
        return y

    def extra_repr(self):
        return '%d ch, level %d -> %d, corner_mode=%s' % (self.in_features, self.subdivisions, self.subdivisions + 1, self.corner_mode)


def set_impl(module, impl):
    """Force 'auto' | 'simt' | 'tc' on every IcoConvS2S below `module` (parity tests use this)."""
    code = {'auto': _lib.IMPL_AUTO, 'simt': _lib.IMPL_SIMT, 'tc': _lib.IMPL_TC}[impl]
    for m in module.modules():
        if isinstance(m, IcoConvS2S):
            m.impl = code
    return module
