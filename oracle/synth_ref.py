"""ORACLE (test infrastructure, NOT product code) -- synthetic level-s meshes in the reference's tensor contract.

SURVEY.md 8d recipe, restated over the oracle geometry only (no product import): a unit icosphere radially displaced by a
smooth seeded field, |xyz| < 1 (the network ends in tanh, models.py:154).

    input  float32 [3, 5n, 2n]   xyz of the P grid vertices                      data.py:64-69
    target float32 [9, P+2]      xyz | vertex normals | Laplacian                generate.py:200-203

Normals follow the generate.py:20-43 recipe (area weighted, eps clip) with true accumulation; the Laplacian is the uniform
graph Laplacian of oracle/mesh_ref.py.  Same seeds and formulas as the product's data.synthetic_mesh, which
tests/test_datapath.py compares against this file.
"""
import numpy as np
import torch

from . import ico_geometry_ref as geo

_topo = {}


def _topology(s):
    if s not in _topo:
        v = geo._positions_only(s).astype(np.float32).astype(np.float64)     # the product reads float32 vertices from its C ABI
        f = geo.get_ico_faces(s)
        half = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e = np.unique(np.concatenate([half, half[:, ::-1]]), axis=0)
        _topo[s] = (v, f, e, np.bincount(e[:, 0], minlength=v.shape[0]).astype(np.float64))
    return _topo[s]


def synthetic_mesh(s, sample_idx):
    base, f, e, deg = _topology(s)
    g = torch.Generator().manual_seed(1234 + int(sample_idx))
    amp = torch.rand(4, generator=g).numpy() * 2 - 1
    omega = torch.randn(4, 3, generator=g).numpy() * 2.0
    phase = torch.rand(4, generator=g).numpy() * 2 * np.pi
    radius = 0.5 * (1.0 + 0.25 * sum(amp[j] * np.sin(base @ omega[j] + phase[j]) for j in range(4)))
    v = base * radius[:, None]
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    vn = np.zeros_like(v)
    for c in range(3):
        np.add.at(vn, f[:, c], fn)
    vn /= np.clip(np.sqrt((vn ** 2).sum(1)), 1e-10, None)[:, None]
    nb_sum = np.zeros_like(v)
    np.add.at(nb_sum, e[:, 0], v[e[:, 1]])
    tgt = np.concatenate([v, vn, nb_sum / deg[:, None] - v], axis=1).T.astype(np.float32)
    n = 2 ** s
    return torch.from_numpy(np.ascontiguousarray(tgt[:3, :-2].reshape(3, 5 * n, 2 * n))), torch.from_numpy(np.ascontiguousarray(tgt))


def synthetic_batch(s, first_idx, batch):
    xs, ts = zip(*(synthetic_mesh(s, first_idx + i) for i in range(batch)))
    return torch.stack(xs), torch.stack(ts)
