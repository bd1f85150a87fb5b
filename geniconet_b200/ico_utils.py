"""Evaluation helpers with the reference's names (ico_utils.py:10-103), SURVEY 8f rank 4.

output2vertices     grid -> vertex list with averaged poles, on the GPU (gin_pole_vertices_fwd; ico_utils.py:10-24)
computeDistance     mode 'point2mesh': mean squared point-to-surface distance of the output vertices to the reference mesh
                    (ico_utils.py:26-44 over kaolin 0.9.1 point_to_mesh_distance), on the GPU (gin_point_mesh_distance)
point_to_mesh_distance   the batched primitive underneath, same return convention as kaolin's (distance, face index)

The reference's logging / file helpers (saveDistance, getEpochNumber, save_to_file, get_*_shape; ico_utils.py:46-103) are host
utilities outside the hot path and are not provided.

No CPU fallback: CPU tensors raise RuntimeError.
"""
import torch

from . import _lib
from .ico_conv import _stream
from .losses import output2vertices  # noqa: F401  (ico_utils.py:10-24; the kernel-backed implementation lives with the losses)


def point_to_mesh_distance(pointclouds, vertices, faces):
    """pointclouds [B,N,3], vertices [B,V,3] float32 CUDA, faces [F,3] integer -> (squared distance [B,N] float32,
    face index [B,N] int64).  kaolin 0.9.1 also returns a region code; the reference discards it (ico_utils.py:40)."""
    for name, t in (('pointclouds', pointclouds), ('vertices', vertices)):
        if not isinstance(t, torch.Tensor) or t.dtype != torch.float32 or t.dim() != 3 or t.shape[-1] != 3:
            raise ValueError('point_to_mesh_distance: %s must be a float32 [B,n,3] tensor' % name)
        if not t.is_cuda:
            raise RuntimeError('point_to_mesh_distance: %s is on the CPU; geniconet_b200 has no CPU path' % name)
    if pointclouds.shape[0] != vertices.shape[0]:
        raise ValueError('point_to_mesh_distance: batch sizes differ (%d vs %d)' % (pointclouds.shape[0], vertices.shape[0]))
    faces = torch.as_tensor(faces)
    if faces.dim() != 2 or faces.shape[1] != 3 or faces.is_floating_point():
        raise ValueError('point_to_mesh_distance: faces must be an integer [F,3] tensor')
    B, N, _ = pointclouds.shape
    V, F = vertices.shape[1], faces.shape[0]
    if F and (int(faces.min()) < 0 or int(faces.max()) >= V):
        raise ValueError('point_to_mesh_distance: face index out of range [0,%d)' % V)
    dev = pointclouds.device
    p, v = pointclouds.contiguous(), vertices.contiguous()
    f = faces.to(device=dev, dtype=torch.int32).contiguous()
    dist = torch.empty((B, N), dtype=torch.float32, device=dev)
    fidx = torch.empty((B, N), dtype=torch.int32, device=dev)
    ws = torch.empty(max(1, _lib.lib.gin_point_mesh_ws_bytes(B, N)), dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib.gin_point_mesh_distance(p.data_ptr(), v.data_ptr(), f.data_ptr(), dist.data_ptr(), fidx.data_ptr(), ws.data_ptr(),
                                                B, N, V, F, _stream()), 'gin_point_mesh_distance')
    return dist, fidx.long()


def writeOffMesh(path, vertices, faces):
    """python_utils.writeOffMesh (PythonFunctions, absent): plain OFF text."""
    v = torch.as_tensor(vertices).detach().cpu().numpy().reshape(-1, 3)
    f = torch.as_tensor(faces).detach().cpu().numpy().reshape(-1, 3)
    with open(path if str(path).endswith('.off') else str(path) + '.off', 'w') as fh:
        fh.write('OFF\n%d %d 0\n' % (v.shape[0], f.shape[0]))
        for row in v:
            fh.write('%.8f %.8f %.8f\n' % tuple(row))
        for row in f:
            fh.write('3 %d %d %d\n' % tuple(row))


def computeDistance(outvertices, refvertices, reffaces, f, mode='point2point', write_mesh=False, outfaces=None):
    """ico_utils.py:26-44.  outvertices [N,3], refvertices [V,3], reffaces [F,3]; 'point2mesh' -> numpy scalar, the mean
    squared distance; any other mode -> None (as the reference)."""
    if write_mesh:
        writeOffMesh(f, outvertices, reffaces if outfaces is None else outfaces)
    if mode == 'point2mesh':
        dist, _ = point_to_mesh_distance(outvertices[None, :, :], refvertices[None, :, :], reffaces)
        return torch.mean(dist).cpu().numpy()
    return None
