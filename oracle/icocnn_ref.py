"""ORACLE (test infrastructure, NOT product code) -- icosahedral layers on CPU.

Pure-PyTorch restatement of ``icocnn.ico_conv.IcoConvS2S`` /
``IcoUpsampleS2S`` as the reference uses them (/root/reference/models.py:5-6,
13-15, 25-33, 45-55, 104, 165, 269, 279).  The real icocnn package is an
un-vendored, un-pinned dependency that is absent from /root/reference, so the
arithmetic follows SURVEY.md sections 8a / 9:

  a1  chart padding = gather through pad_index_map, pole cells filled by
      corner_mode ('zeros' | 'average' = mean of the pole's 5 ring pixels,
      the same rule losses.py:22-31,49-51 uses for the pole vertices)
  a2  hex-masked 3x3 cross-correlation, 7 live taps, weight [Cout,Cin,7]
  a3  stride 2 = the same contraction at fine pixels (2I+1, 2J)
  a4  upsample = coarse copy + edge-midpoint mean (parameter free)

    parity unpinned (see oracle/ico_geometry_ref.py).

Runs in fp32 or fp64 on the CPU (or any torch device); autograd provides the
reference dgrad/wgrad.  Only tests/, smoke() and bench.py's cpu_baseline leg
may import this.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ico_geometry_ref as geo

TAPS = geo.TAPS
# position of tap t inside the 3x3 window (row-major index)
_TAP_TO_3X3 = [(di + 1) * 3 + (dj + 1) for (di, dj) in TAPS]


class _PadIco(torch.nn.Module):
    """[B,C,5n,2n] -> [B,C,5,n+2,2n+2] through the index map (row a1)."""

    def __init__(self, subdivisions, corner_mode):
        super().__init__()
        if corner_mode not in ('zeros', 'average'):
            raise ValueError('corner_mode must be zeros or average, got %r' % (corner_mode,))
        s = subdivisions
        self.s = s
        self.n = 2 ** s
        self.P = geo.n_pixels(s)
        self.corner_mode = corner_mode
        idx = geo.pad_index_map(s).astype(np.int64)
        # extended vector: [pixels(P) | north | south | zero]
        idx = np.where(idx == geo.NEVER, self.P + 2, idx)
        self.register_buffer('idx', torch.from_numpy(idx.reshape(-1)), persistent=False)
        self.register_buffer('rings', torch.from_numpy(geo.pole_rings(s)), persistent=False)

    def forward(self, x):
        B, C, H, W = x.shape
        if H != 5 * self.n or W != 2 * self.n:
            raise ValueError('expected [B,C,%d,%d], got %s' % (5 * self.n, 2 * self.n, tuple(x.shape)))
        flat = x.reshape(B, C, self.P)
        if self.corner_mode == 'average':
            poles = flat[:, :, self.rings].mean(-1)           # [B,C,2]
        else:
            poles = flat.new_zeros(B, C, 2)
        ext = torch.cat((flat, poles, flat.new_zeros(B, C, 1)), dim=2)
        return ext[:, :, self.idx].reshape(B, C, 5, self.n + 2, 2 * self.n + 2)


class IcoConvS2S(torch.nn.Module):
    def __init__(self, in_features, out_features, stride=1, bias=True, subdivisions=0,
                 corner_mode='zeros'):
        super().__init__()
        if stride not in (1, 2):
            raise ValueError('stride must be 1 or 2')
        if stride == 2 and subdivisions < 1:
            raise ValueError('stride 2 needs subdivisions >= 1')
        self.in_features, self.out_features = in_features, out_features
        self.stride, self.subdivisions, self.corner_mode = stride, subdivisions, corner_mode
        self.pad = _PadIco(subdivisions, corner_mode)
        self.weight = torch.nn.Parameter(torch.empty(out_features, in_features, 7))
        self.bias = torch.nn.Parameter(torch.empty(out_features)) if bias else None
        bound = 1.0 / math.sqrt(in_features * 7)
        torch.nn.init.uniform_(self.weight, -bound, bound)
        if bias:
            torch.nn.init.uniform_(self.bias, -bound, bound)
        self.register_buffer('tap_pos', torch.tensor(_TAP_TO_3X3), persistent=False)

    def forward(self, x):
        B = x.shape[0]
        n = self.pad.n
        xp = self.pad(x)                                              # [B,C,5,n+2,2n+2]
        w33 = self.weight.new_zeros(self.out_features, self.in_features, 9)
        w33 = w33.index_copy(2, self.tap_pos, self.weight).reshape(self.out_features, self.in_features, 3, 3)
        xb = xp.permute(0, 2, 1, 3, 4).reshape(B * 5, self.in_features, n + 2, 2 * n + 2)
        if self.stride == 1:
            y = F.conv2d(xb, w33, self.bias)                          # [B*5,Co,n,2n]
            no = n
        else:
            # coarse (I,J) is centred on fine (2I+1, 2J): window rows 2I+1..2I+3, cols 2J..2J+2 of the padded chart
            y = F.conv2d(xb[:, :, 1:, :], w33, self.bias, stride=2)   # [B*5,Co,n/2,n]
            no = n // 2
        y = y.reshape(B, 5, self.out_features, no, 2 * no).permute(0, 2, 1, 3, 4)
        return y.reshape(B, self.out_features, 5 * no, 2 * no)


class IcoUpsampleS2S(torch.nn.Module):
    def __init__(self, in_features, subdivisions, corner_mode='zeros'):
        super().__init__()
        if corner_mode not in ('zeros', 'average'):
            raise ValueError('corner_mode must be zeros or average, got %r' % (corner_mode,))
        self.in_features, self.subdivisions, self.corner_mode = in_features, subdivisions, corner_mode
        s = subdivisions
        self.P = geo.n_pixels(s)
        self.n = 2 ** s
        src = geo.upsample_sources(s)
        self.register_buffer('src', torch.from_numpy(src), persistent=False)
        self.register_buffer('rings', torch.from_numpy(geo.pole_rings(s)), persistent=False)

    def forward(self, x):
        B, C, H, W = x.shape
        if H != 5 * self.n or W != 2 * self.n:
            raise ValueError('expected [B,C,%d,%d], got %s' % (5 * self.n, 2 * self.n, tuple(x.shape)))
        flat = x.reshape(B, C, self.P)
        if self.corner_mode == 'average':
            poles = flat[:, :, self.rings].mean(-1)
        else:
            poles = flat.new_zeros(B, C, 2)
        ext = torch.cat((flat, poles), dim=2)
        y = 0.5 * (ext[:, :, self.src[:, 0]] + ext[:, :, self.src[:, 1]])
        return y.reshape(B, C, 10 * self.n, 4 * self.n)
