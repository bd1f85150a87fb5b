"""Known-answer tests that do not depend on icocnn: they pin the index maps of BOTH the
oracle (oracle/ico_geometry_ref.py) and the product's C++ generator (gin_host.cpp through the C ABI) against the
mesh itself, and against each other bit-exactly."""
import numpy as np
import pytest
import torch

from oracle import ico_geometry_ref as geo
from oracle import icocnn_ref
from geniconet_b200 import _lib

LEVELS = [0, 1, 2, 3, 4, 5]


@pytest.mark.parametrize('s', LEVELS + [6, 7])
def test_index_map_bit_exact_product_vs_oracle(s):
    got = _lib.index_map(s)
    assert np.array_equal(got, geo.pad_index_map(s))          # every level the layer sweep uses (I3-I7) and below
    n, P = 2 ** s, 10 * 4 ** s
    assert got.shape == (5, n + 2, 2 * n + 2) and got.dtype == np.int32
    # interior cells are the identity
    own = got[:, 1:-1, 1:-1].reshape(-1)
    assert np.array_equal(own, np.arange(P, dtype=np.int32))
    assert (got == -1).sum() == 10 and (got == P).sum() == 5 and (got == P + 1).sum() == 5


@pytest.mark.parametrize('s', LEVELS)
def test_topology_invariants(s):
    for faces in (geo.get_ico_faces(s), _lib.ico_faces(s).astype(np.int64)):
        P = 10 * 4 ** s
        V = P + 2
        assert faces.shape == (20 * 4 ** s, 3) and faces.max() + 1 == V          # losses.py:38
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
        directed = set(map(tuple, e.tolist()))
        assert len(directed) == len(e)                                            # consistent winding: every directed edge once
        und = {tuple(sorted(x)) for x in directed}
        assert len(und) == 30 * 4 ** s and V - len(und) + len(faces) == 2         # Euler characteristic
        val = np.bincount(np.array(list(und)).reshape(-1), minlength=V)
        assert (val == 5).sum() == 12 and (val == 6).sum() == V - 12
        assert all((b, a) in directed for (a, b) in directed)                     # closed manifold


@pytest.mark.parametrize('s', LEVELS)
def test_faces_product_equals_oracle_and_point_outward(s):
    def canon(F):
        return sorted(tuple(np.roll(t, -int(np.argmin(t)))) for t in F.tolist())
    fp, fo = _lib.ico_faces(s).astype(np.int64), geo.get_ico_faces(s)
    assert canon(fp) == canon(fo)
    v = _lib.ico_vertices(s).astype(np.float64)
    assert np.allclose(v, geo.get_icosahedral_grid(s)[0], atol=1e-6)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-6)
    a, b, c = v[fp[:, 0]], v[fp[:, 1]], v[fp[:, 2]]
    assert (np.einsum('ij,ij->i', np.cross(b - a, c - a), a + b + c) > 0).all()


@pytest.mark.parametrize('s', [1, 2, 3, 5])
def test_pole_rings_match_reference_index_buffers(s):
    """losses.py:23-29: corner_src_y = [[0,n,..,4n],[n-1,..,5n-1]], corner_src_x = [[0],[-1]]."""
    n = 2 ** s
    top_y, bot_y = np.arange(5) * n, np.arange(1, 6) * n - 1
    want = np.stack([top_y * 2 * n + 0, bot_y * 2 * n + (2 * n - 1)])
    assert np.array_equal(geo.pole_rings(s), want)
    faces = _lib.ico_faces(s)
    P = 10 * 4 ** s
    for pole in (0, 1):
        ring = set(faces[(faces == P + pole).any(1)].reshape(-1).tolist()) - {P + pole}
        assert ring == set(want[pole].tolist())


@pytest.mark.parametrize('s', [1, 2, 3, 4])
def test_padding_is_mesh_adjacency(s):
    """Every pixel's 6 live taps through the padded chart hit exactly its one-ring in the face mesh
    (the 10 non-pole icosahedron corners hit one neighbour twice: the 60-degree wedge)."""
    idx = _lib.index_map(s)
    faces = _lib.ico_faces(s)
    n, P = 2 ** s, 10 * 4 ** s
    ring = [set() for _ in range(P + 2)]
    for a, b, c in faces.tolist():
        ring[a] |= {b, c}; ring[b] |= {a, c}; ring[c] |= {a, b}
    dup = 0
    for k in range(5):
        for i in range(n):
            for j in range(2 * n):
                hits = [int(idx[k, i + 1 + di, j + 1 + dj]) for (di, dj) in geo.TAPS[1:]]
                p = k * n * 2 * n + i * 2 * n + j
                assert set(hits) == ring[p], (k, i, j)
                dup += len(hits) - len(set(hits))
    assert dup == 10


@pytest.mark.parametrize('s', [2, 3])
def test_laplacian_identity(s):
    """Hex-conv with centre -1 and ring 1/6 applied to xyz equals the uniform graph Laplacian at valence-6 vertices."""
    from oracle import mesh_ref
    v, f = geo.get_icosahedral_grid(s)
    n, P = 2 ** s, 10 * 4 ** s
    conv = icocnn_ref.IcoConvS2S(3, 3, 1, False, s, 'average').double()
    with torch.no_grad():
        conv.weight.zero_()
        for c in range(3):
            conv.weight[c, c, 0] = -1.0
            conv.weight[c, c, 1:] = 1.0 / 6.0
    x = torch.from_numpy(v[:P].T.reshape(1, 3, 5 * n, 2 * n).copy())
    y = conv(x).reshape(3, P).T
    ft = torch.from_numpy(f)
    lap = mesh_ref.compute_laplacian(torch.from_numpy(v), mesh_ref.compute_adjacency_matrix_sparse(P + 2, ft).double())
    deg = np.bincount(f.reshape(-1), minlength=P + 2)
    # pixels next to a pole see the pole through corner_mode='average' = mean of the ring, not the pole vertex itself
    near_pole = set(geo.pole_rings(s).reshape(-1).tolist())
    ok = [p for p in range(P) if deg[p] == 6 and p not in near_pole]
    assert torch.allclose(y[ok], lap[ok], atol=1e-12)


@pytest.mark.parametrize('s', [1, 2, 3])
def test_isotropic_kernel_commutes_with_icosahedral_rotations(s):
    """An isotropic hex kernel must commute with the mesh symmetries; any stitching error breaks this."""
    v, f = geo.get_icosahedral_grid(s)
    n, P = 2 ** s, 10 * 4 ** s
    conv = icocnn_ref.IcoConvS2S(1, 1, 1, False, s, 'average').double()
    with torch.no_grad():
        conv.weight[0, 0, 0] = 0.3
        conv.weight[0, 0, 1:] = 0.11
    # rotation by 72 degrees about the pole axis maps chart k -> k+1: a pure roll of the chart axis
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 1, 5 * n, 2 * n, generator=g, dtype=torch.double)
    assert torch.allclose(conv(torch.roll(x, n, dims=2)), torch.roll(conv(x), n, dims=2), atol=1e-12)
    # a generic icosahedral rotation: build the vertex permutation from the geometry
    N, S, U, L = geo._corner_positions()
    # rotation taking N -> U[0] that is a symmetry of the icosahedron: 120-degree turn about the centre of face (N, U0, U4)
    axis = (N + U[0] + U[4]); axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(2 * np.pi / 3) * K + (1 - np.cos(2 * np.pi / 3)) * K @ K
    rv = v @ R.T
    d = ((rv[:, None, :] - v[None, :, :]) ** 2).sum(-1)
    perm = d.argmin(1)
    assert d[np.arange(P + 2), perm].max() < 1e-18 and len(set(perm.tolist())) == P + 2
    # field on vertices incl. poles consistent with corner_mode='average': poles are not free values, so use a field
    # that is the restriction of a smooth function and compare only vertices whose one-ring avoids the poles' rings
    fld = np.sin(3 * v[:, 0]) + v[:, 1] * v[:, 2]
    def run(field):
        xx = torch.from_numpy(field[:P].reshape(1, 1, 5 * n, 2 * n).copy())
        return conv(xx).detach().reshape(P).numpy()
    y = run(fld)
    fld_rot = np.empty_like(fld); fld_rot[perm] = fld          # value travels with the vertex
    y_rot = run(fld_rot)
    ring = [set() for _ in range(P + 2)]
    for a, b, c in f.tolist():
        ring[a] |= {b, c}; ring[b] |= {a, c}; ring[c] |= {a, b}
    deg = np.array([len(r) for r in ring])
    ok = [p for p in range(P) if perm[p] < P and deg[p] == 6 and deg[perm[p]] == 6
          and not (ring[p] & {P, P + 1}) and not (ring[perm[p]] & {P, P + 1})]
    assert len(ok) > 0 or s == 1
    assert np.allclose(y_rot[perm[ok]], y[ok], atol=1e-12)


@pytest.mark.parametrize('s', [1, 2, 3, 4])
def test_stride2_and_upsample_lattices(s):
    """Coarse vertices are fine pixels (2I+1, 2J); every other fine vertex is the midpoint of exactly one coarse edge."""
    cf = geo.coarse_to_fine(s)
    up = geo.upsample_sources(s - 1)
    nb = geo.neighbours(s)
    P, Pc = 10 * 4 ** s, 10 * 4 ** (s - 1)
    fine_of = lambda c: int(cf[c]) if c < Pc else (P if c == Pc else P + 1)
    copies = 0
    for fv in range(P):
        a, b = int(up[fv, 0]), int(up[fv, 1])
        if a == b:
            assert fine_of(a) == fv
            copies += 1
        else:
            assert fine_of(a) in nb[fv] and fine_of(b) in nb[fv]
    assert copies == Pc
    coarse_faces = geo.get_ico_faces(s - 1)
    edges = {tuple(sorted(x)) for x in np.concatenate([coarse_faces[:, [0, 1]], coarse_faces[:, [1, 2]], coarse_faces[:, [2, 0]]]).tolist()}
    mids = {tuple(sorted((int(a), int(b)))) for a, b in up.tolist() if a != b}
    assert mids == edges
