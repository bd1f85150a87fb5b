"""Host-side mirror of the `mesh.utils` functions the reference's losses.py / generate.py import
(losses.py:7,39,54,57; generate.py:13,152,197), so that the UNMODIFIED losses.py runs over the `mesh` shim package:

    compute_adjacency_matrix_sparse(n_vertices, faces) -> sparse COO float32 [V,V] (a real tensor: losses.py:40 registers it as
                                                          a buffer and `.to(device)` moves it with the module)
    compute_vertex_normals(v, faces)                    -> unit normals [B,V,3]       (losses.py:54)
    compute_laplacian_batch(v, adj)                     -> mean(ring) - v  [B,V,3]    (losses.py:57)

On CUDA tensors the normals and the Laplacian run on the loss plan's one-ring table (csrc/gin_loss.cuh) and are differentiable
(gin_ring_ops_bwd); the mesh is identified by its vertex count, which must be an icosahedral grid (10*4^s + 2), and by the face
count where faces are passed.  There is no CPU path."""
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


def _check_faces(faces, level):
    if faces is not None and hasattr(faces, 'shape') and tuple(faces.shape) != (20 * 4 ** level, 3):
        raise ValueError('mesh.utils: expected the %d faces of the level-%d icosahedral grid, got %s' % (20 * 4 ** level, level, tuple(faces.shape)))


def compute_adjacency_matrix_sparse(n_vertices, faces=None):
    """Symmetric 0/1 vertex adjacency as a coalesced sparse COO tensor.  Built on the host from `faces` when given (any mesh),
    else from the icosahedral grid of that size."""
    V = int(n_vertices)
    if faces is None:
        from .ico_geometry import get_ico_faces
        faces = get_ico_faces(_level_of(V))
    f = torch.as_tensor(faces).detach().cpu().long()
    e = torch.cat((f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]), dim=0)
    key = torch.unique(torch.cat((e[:, 0] * V + e[:, 1], e[:, 1] * V + e[:, 0])))
    idx = torch.stack((torch.div(key, V, rounding_mode='floor'), key % V))
    return torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (V, V)).coalesce()


class _RingOp(torch.autograd.Function):
    """which = 'normals' | 'laplacian'.  forward: one launch over the one-ring table; backward: gin_ring_ops_bwd."""

    @staticmethod
    def forward(ctx, v, level, which):
        _require_cuda_f32(v, 'mesh.utils')
        v = v.contiguous()
        plan = get_plan(_lib.PLAN_LOSS, level, 1, 'average', v.device)
        out = torch.empty_like(v)
        fn = _lib.lib.gin_vertex_normals_fwd if which == 'normals' else _lib.lib.gin_laplacian_fwd
        _lib.check(fn(plan.host_ptr, plan.dev_ptr, v.data_ptr(), out.data_ptr(), v.shape[0], _stream()), 'mesh.utils.' + which)
        ctx.save_for_backward(v)
        ctx.plan, ctx.level, ctx.which = plan, level, which
        return out

    @staticmethod
    def backward(ctx, g):
        v, = ctx.saved_tensors
        g = g.contiguous()
        B = v.shape[0]
        dv = torch.empty_like(v)
        ws = torch.empty(max(1, _lib.lib.gin_ring_ops_ws_bytes(B, ctx.level)), dtype=torch.uint8, device=v.device)
        gn = g.data_ptr() if ctx.which == 'normals' else None
        gl = g.data_ptr() if ctx.which == 'laplacian' else None
        _lib.check(_lib.lib.gin_ring_ops_bwd(ctx.plan.host_ptr, ctx.plan.dev_ptr, v.data_ptr(), gn, gl, dv.data_ptr(), ws.data_ptr(), B, _stream()),
                   'gin_ring_ops_bwd')
        return dv, None, None


def compute_vertex_normals(vertices, faces=None):
    squeeze = vertices.dim() == 2
    v = vertices.unsqueeze(0) if squeeze else vertices
    level = _level_of(v.shape[1])
    _check_faces(faces, level)
    out = _RingOp.apply(v, level, 'normals')
    return out[0] if squeeze else out


def compute_laplacian_batch(vertices, adj=None):
    if adj is not None and hasattr(adj, 'shape') and adj.shape[0] != vertices.shape[1]:
        raise ValueError('mesh.utils: adjacency is %s for %d vertices' % (tuple(adj.shape), vertices.shape[1]))
    return _RingOp.apply(vertices, _level_of(vertices.shape[1]), 'laplacian')


def compute_laplacian(vertices, adj=None):
    return compute_laplacian_batch(vertices.unsqueeze(0), adj)[0]
