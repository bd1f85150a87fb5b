"""Synthetic level-s meshes in the reference's on-disk tensor contract (data.py:64-69,
generate.py:200-203): input [3,5n,2n] = xyz of the P grid vertices, target [9,P+2] = xyz | normals |
Laplacian.  SURVEY.md 8d recipe: an icosphere radially displaced by a smooth seeded field, |xyz| < 1.

Targets are made on the host with numpy (the generate.py:20-43 normals recipe and the uniform
Laplacian) -- dataset preparation, not the hot path.
"""
import numpy as np
import torch

from .ico_geometry import get_icosahedral_grid

_cache = {}


def _topology(s):
    if s not in _cache:
        v, f = get_icosahedral_grid(s)
        V = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e = np.concatenate([e, e[:, ::-1]])
        e = np.unique(e, axis=0)
        deg = np.bincount(e[:, 0], minlength=V).astype(np.float64)
        _cache[s] = (v.astype(np.float64), f, e, deg)
    return _cache[s]


def vertex_normals(v, f, eps=1e-10):
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    vn = np.zeros_like(v)
    for c in range(3):
        np.add.at(vn, f[:, c], fn)
    return vn / np.clip(np.sqrt((vn ** 2).sum(1)), eps, None)[:, None]


def laplacian(v, e, deg):
    acc = np.zeros_like(v)
    np.add.at(acc, e[:, 0], v[e[:, 1]])
    return acc / deg[:, None] - v


def synthetic_mesh(s, sample_idx):
    """(input float32 [3,5n,2n], target float32 [9,P+2]) for one seeded sample."""
    base, f, e, deg = _topology(s)
    g = torch.Generator().manual_seed(1234 + int(sample_idx))
    a = torch.rand(4, generator=g).numpy() * 2 - 1
    w = torch.randn(4, 3, generator=g).numpy() * 2.0
    ph = torch.rand(4, generator=g).numpy() * 2 * np.pi
    field = sum(a[j] * np.sin(base @ w[j] + ph[j]) for j in range(4))
    r = 0.5 * (1.0 + 0.25 * field)
    v = base * r[:, None]
    tgt = np.concatenate([v, vertex_normals(v, f), laplacian(v, e, deg)], axis=1).T.astype(np.float32)   # [9,P+2]
    n = 2 ** s
    x = np.ascontiguousarray(tgt[:3, :-2].reshape(3, 5 * n, 2 * n))
    return torch.from_numpy(x), torch.from_numpy(np.ascontiguousarray(tgt))


def synthetic_batch(s, first_idx, batch):
    xs, ts = zip(*(synthetic_mesh(s, first_idx + i) for i in range(batch)))
    return torch.stack(xs), torch.stack(ts)
