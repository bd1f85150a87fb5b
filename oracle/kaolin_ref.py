"""ORACLE (test infrastructure, NOT product code) -- point-to-mesh distance on CPU, float64.

The reference's evaluation metric (/root/reference/ico_utils.py:26-44, mode 'point2mesh') calls
``kaolin.metrics.trianglemesh.point_to_mesh_distance`` of kaolin 0.9.1 (pinned by Dockerfile:50-52; the package is
absent here) and averages its first return value.  Published behaviour of that function: for every point the SQUARED
Euclidean distance to the closest triangle of the mesh, the index of that triangle, and a region code (unused by the
reference, not restated).

Restated from the definition, deliberately NOT with the region-test algorithm the CUDA kernel uses: the distance to a
triangle is the distance to its plane when the projection falls inside it, otherwise the smallest distance to its three
edge segments.

    parity unpinned (no golden vectors for this function exist in the reference).
"""
import numpy as np


def _segment_sq(p, a, b):
    """p [N,1,3], a/b [1,F,3] -> squared distance to segment ab, [N,F]."""
    ab = b - a
    t = ((p - a) * ab).sum(-1) / np.maximum((ab * ab).sum(-1), 1e-300)
    t = np.clip(t, 0.0, 1.0)[..., None]
    d = p - (a + t * ab)
    return (d * d).sum(-1)


def point_to_mesh_distance(points, vertices, faces, chunk=256):
    """points [N,3], vertices [V,3], faces [F,3] int -> (squared distance [N] float64, face index [N] int64)."""
    p_all = np.asarray(points, dtype=np.float64)
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    a, b, c = v[f[:, 0]][None], v[f[:, 1]][None], v[f[:, 2]][None]
    nrm = np.cross(b - a, c - a)
    nn = (nrm * nrm).sum(-1)
    out_d = np.empty(p_all.shape[0])
    out_f = np.empty(p_all.shape[0], dtype=np.int64)
    for s in range(0, p_all.shape[0], chunk):
        p = p_all[s:s + chunk, None, :]
        best = np.minimum(np.minimum(_segment_sq(p, a, b), _segment_sq(p, b, c)), _segment_sq(p, c, a))
        # inside test: the projection is on the inner side of all three edges
        ap, bp, cp = p - a, p - b, p - c
        s0 = (np.cross(b - a, ap) * nrm).sum(-1)
        s1 = (np.cross(c - b, bp) * nrm).sum(-1)
        s2 = (np.cross(a - c, cp) * nrm).sum(-1)
        inside = (s0 >= 0) & (s1 >= 0) & (s2 >= 0) & (nn > 0)
        plane = ((ap * nrm).sum(-1) ** 2) / np.maximum(nn, 1e-300)
        d = np.where(inside, plane, best)
        out_f[s:s + chunk] = d.argmin(1)
        out_d[s:s + chunk] = d.min(1)
    return out_d, out_f
