"""Host-side mirror of the `mesh.utils` functions the reference's losses.py / generate.py import
(losses.py:7,39,54,57; generate.py:13,152,197).  On CUDA tensors the normals and the Laplacian run
on the loss plan's one-ring table (csrc/gin_loss.cuh); the mesh is identified by its vertex count,
which must be an icosahedral grid (10*4^s + 2)."""
import math

import numpy as np
import torch

from . import _lib
from .ico_conv import get_plan, _stream, _require_cuda_f32


def _level_of(n_vertices):
    s = round(math.log((n_vertices - 2) / 10, 4)) if n_vertices > 2 else -1
    if s < 0 or 10 * 4 ** s + 2 != n_vertices:
        raise ValueError('not an icosahedral grid: %d vertices' % n_vertices)
    return s


class IcoAdjacency:
    """What compute_adjacency_matrix_sparse returns here: a handle on the level's one-ring table."""

    def __init__(self, n_vertices):
        self.n_vertices = int(n_vertices)
        self.level = _level_of(self.n_vertices)

    def to(self, *a, **k):
        return self


def compute_adjacency_matrix_sparse(n_vertices, faces=None):
    return IcoAdjacency(int(n_vertices))


class _RingOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, level, which):
        _require_cuda_f32(v, 'mesh.utils')
        v = v.contiguous()
        plan = get_plan(_lib.PLAN_LOSS, level, 1, 'average', v.device)
        out = torch.empty_like(v)
        fn = _lib.lib.gin_vertex_normals_fwd if which == 'normals' else _lib.lib.gin_laplacian_fwd
        _lib.check(fn(plan.host_ptr, plan.dev_ptr, v.data_ptr(), out.data_ptr(), v.shape[0], _stream()), 'mesh.utils.' + which)
        return out

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError('mesh.utils ops are forward-only here; the differentiable path is geniconet_b200.losses')


def compute_vertex_normals(vertices, faces=None):
    squeeze = vertices.dim() == 2
    v = vertices.unsqueeze(0) if squeeze else vertices
    out = _RingOp.apply(v, _level_of(v.shape[1]), 'normals')
    return out[0] if squeeze else out


def compute_laplacian_batch(vertices, adj=None):
    return _RingOp.apply(vertices, _level_of(vertices.shape[1]), 'laplacian')


def compute_laplacian(vertices, adj=None):
    return compute_laplacian_batch(vertices.unsqueeze(0), adj)[0]
