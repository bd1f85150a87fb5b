"""ORACLE (test infrastructure, NOT product code) -- icosahedral chart geometry.

CPU/numpy restatement of what the reference imports as
``icocnn.utils.ico_geometry`` (call sites: /root/reference/losses.py:5,34,
generate.py:11,151, run.py:144,529).  The dependency itself
(github.com/hrdkjain/IcosahedralCNN, cloned at HEAD, never version-pinned --
README.md:18-24) is NOT in /root/reference and not on this machine, so this
file restates the geometry from the facts the reference *does* pin:

  * tensor layout [B,C,5n,2n], chart k = rows [k*n,(k+1)*n)   data.py:64-69
  * vertex id of pixel (r,c) = r*2n+c, poles are P and P+1     ico_utils.py:20-23
  * pole 0 ring = pixels (k*n,0); pole 1 ring = ((k+1)*n-1,-1) losses.py:23-29
  * get_ico_faces(s).max()+1 == P+2                            losses.py:38

and from the derived stitching of SURVEY.md section 9.2 ("option A").

    parity unpinned: the reference ships no tests / golden vectors for this
    path and its icocnn dependency is absent; this oracle is pinned only by
    the mesh-topology known-answer tests in tests/test_geometry.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
import numpy as np

# 7 live taps (di, dj) of the hex-masked 3x3 stencil; masked cells are
# (-1,-1) and (+1,+1).  SURVEY.md section 9.2.  Order == weight[..., t] order.
TAPS = ((0, 0), (-1, 0), (1, 0), (0, -1), (0, 1), (-1, 1), (1, -1))

NEVER = -1  # padded cells (-1,-1) and (n,2n): never read by a live tap


def n_pixels(s):
    return 10 * 4 ** s


def vid(s, k, i, j):
    n = 2 ** s
    return (k % 5) * n * 2 * n + i * 2 * n + j


def source(s, k, i, j):
    """Vertex id feeding padded cell (i,j), i in [-1,n], j in [-1,2n], of chart k.

    Direct transcription of the rule table in SURVEY.md section 9.2.
    Returns P for the north pole, P+1 for the south pole, NEVER for the two
    cells no live tap ever reads.
    """
    n = 2 ** s
    P = n_pixels(s)
    if 0 <= i < n and 0 <= j < 2 * n:
        return vid(s, k, i, j)
    if (i, j) == (-1, 0):
        return P
    if (i, j) == (n - 1, 2 * n):
        return P + 1
    if (i, j) in ((-1, -1), (n, 2 * n)):
        return NEVER
    if i == -1:
        return vid(s, k - 1, j - 1, 0) if 1 <= j <= n else vid(s, k - 1, n - 1, j - n)
    if j == 2 * n:
        return vid(s, k - 1, n - 1, n + i + 1)
    if i == n:
        return vid(s, k + 1, 0, j + n) if -1 <= j <= n - 1 else vid(s, k + 1, j - n, 2 * n - 1)
    if j == -1:
        return vid(s, k + 1, 0, i)
    raise AssertionError((s, k, i, j))


def pad_index_map(s):
    """int32 [5, n+2, 2n+2]; entry [k, i+1, j+1] = source(s,k,i,j)."""
    n = 2 ** s
    out = np.empty((5, n + 2, 2 * n + 2), dtype=np.int32)
    for k in range(5):
        for i in range(-1, n + 1):
            for j in range(-1, 2 * n + 1):
                out[k, i + 1, j + 1] = source(s, k, i, j)
    return out


def pole_rings(s):
    """[2,5] pixel ids averaged into the two poles (losses.py:23-29)."""
    n = 2 ** s
    north = [k * n * 2 * n + 0 for k in range(5)]                      # (k*n, 0)
    south = [((k + 1) * n - 1) * 2 * n + (2 * n - 1) for k in range(5)]  # ((k+1)n-1, 2n-1)
    return np.array([north, south], dtype=np.int64)


def neighbours(s):
    """list over P+2 vertices of the ordered 6 tap targets (pixels) / ring (poles)."""
    n = 2 ** s
    P = n_pixels(s)
    nb = [None] * (P + 2)
    for k in range(5):
        for i in range(n):
            for j in range(2 * n):
                nb[vid(s, k, i, j)] = [source(s, k, i + di, j + dj) for (di, dj) in TAPS[1:]]
    rings = pole_rings(s)
    nb[P] = [int(x) for x in rings[0]]
    nb[P + 1] = [int(x) for x in rings[1]]
    return nb


def _corner_positions():
    """12 icosahedron corners: N, S, upper ring U_k, lower ring L_k (k=0..4)."""
    zc = 1.0 / np.sqrt(5.0)
    rc = 2.0 / np.sqrt(5.0)
    N = np.array([0.0, 0.0, 1.0])
    S = np.array([0.0, 0.0, -1.0])
    U = [np.array([rc * np.cos(2 * np.pi * k / 5), rc * np.sin(2 * np.pi * k / 5), zc]) for k in range(5)]
    # L_k is adjacent to U_k and U_{k-1}  ->  longitude half-way between them
    L = [np.array([rc * np.cos(2 * np.pi * (k - 0.5) / 5), rc * np.sin(2 * np.pi * (k - 0.5) / 5), -zc])
         for k in range(5)]
    return N, S, U, L


def get_icosahedral_grid(s):
    """(ico_v [P+2,3] float64 on the unit sphere, ico_f [20*4^s,3] int64).

    Restates icocnn.utils.ico_geometry.get_icosahedral_grid (generate.py:151).
    Vertices: flat barycentric interpolation inside each of the chart's four
    icosahedron faces, then radial projection.  (Whether real icocnn subdivides
    recursively instead is UNPINNED; only the topology is used by the hot path.)
    """
    n = 2 ** s
    P = n_pixels(s)
    N, S, U, L = _corner_positions()
    v = np.zeros((P + 2, 3))
    for k in range(5):
        Uk, Ukm, Lk, Lkm = U[k], U[(k - 1) % 5], L[k], L[(k - 1) % 5]
        for i in range(n):
            for j in range(2 * n):
                a, b = (i + 1) / n, j / n       # lattice coords: N=(0,0) U_k=(1,0) U_{k-1}=(0,1) L_k=(1,1) L_{k-1}=(0,2) S=(1,2)
                if b <= 1.0:
                    if a + b <= 1.0:
                        p = N + a * (Uk - N) + b * (Ukm - N)
                    else:
                        p = (1 - b) * Uk + (a + b - 1) * Lk + (1 - a) * Ukm
                else:
                    bb = b - 1.0
                    if a + bb <= 1.0:
                        p = Ukm + a * (Lk - Ukm) + bb * (Lkm - Ukm)
                    else:
                        p = (1 - bb) * Lk + (a + bb - 1) * S + (1 - a) * Lkm
                v[vid(s, k, i, j)] = p / np.linalg.norm(p)
    v[P] = N
    v[P + 1] = S
    return v, get_ico_faces(s)


def get_ico_faces(s):
    """int64 [20*4^s, 3], vertex ids in grid order + 2 poles last (losses.py:34-38).

    Each lattice unit cell (i,j),(i+1,j),(i,j+1),(i+1,j+1) is split along the
    live diagonal (i+1,j)-(i,j+1).  Cells are enumerated over every chart's
    padded lattice, mapped through source(), de-duplicated, and wound so that
    the normal points away from the origin (counter-clockwise seen from outside).
    """
    n = 2 ** s
    seen = {}
    for k in range(5):
        for i in range(-1, n):
            for j in range(-1, 2 * n):
                quad = [(i, j), (i + 1, j), (i, j + 1), (i + 1, j + 1)]
                ids = [source(s, k, a, b) for (a, b) in quad]
                for tri in ((ids[0], ids[1], ids[2]), (ids[1], ids[3], ids[2])):
                    if NEVER in tri or len(set(tri)) < 3:
                        continue
                    key = tuple(sorted(tri))
                    seen.setdefault(key, tri)
    faces = np.array(list(seen.values()), dtype=np.int64)
    # orientation: outward
    vv = _positions_only(s)
    a, b, c = vv[faces[:, 0]], vv[faces[:, 1]], vv[faces[:, 2]]
    flip = np.einsum('ij,ij->i', np.cross(b - a, c - a), a + b + c) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    return faces[order]


_pos_cache = {}


def _positions_only(s):
    if s not in _pos_cache:
        n = 2 ** s
        P = n_pixels(s)
        N, S, U, L = _corner_positions()
        v = np.zeros((P + 2, 3))
        kk, ii, jj = np.meshgrid(np.arange(5), np.arange(n), np.arange(2 * n), indexing='ij')
        a = (ii + 1) / n
        b = jj / n
        Ua = np.array(U)
        La = np.array(L)
        Uk, Ukm, Lk, Lkm = Ua[kk], Ua[(kk - 1) % 5], La[kk], La[(kk - 1) % 5]
        a3, b3 = a[..., None], b[..., None]
        bb3 = b3 - 1.0
        t1 = N + a3 * (Uk - N) + b3 * (Ukm - N)
        t2 = (1 - b3) * Uk + (a3 + b3 - 1) * Lk + (1 - a3) * Ukm
        t3 = Ukm + a3 * (Lk - Ukm) + bb3 * (Lkm - Ukm)
        t4 = (1 - bb3) * Lk + (a3 + bb3 - 1) * S + (1 - a3) * Lkm
        p = np.where(b3 <= 1.0, np.where(a3 + b3 <= 1.0, t1, t2), np.where(a3 + bb3 <= 1.0, t3, t4))
        p = p / np.linalg.norm(p, axis=-1, keepdims=True)
        v[:P] = p.reshape(P, 3)
        v[P] = N
        v[P + 1] = S
        _pos_cache[s] = v
    return _pos_cache[s]


def coarse_to_fine(s_fine):
    """int64 [P(s_fine-1)]: fine pixel id of every coarse pixel ((I,J) -> (2I+1, 2J))."""
    nf = 2 ** s_fine
    nc = nf // 2
    out = np.empty(n_pixels(s_fine - 1), dtype=np.int64)
    for k in range(5):
        for I in range(nc):
            for J in range(2 * nc):
                out[vid(s_fine - 1, k, I, J)] = vid(s_fine, k, 2 * I + 1, 2 * J)
    return out


def upsample_sources(s_coarse):
    """For every fine pixel of level s_coarse+1: (src0, src1) coarse vertex ids.

    SURVEY.md section 9.4 (ASSUMPTION): coarse vertices copy (src0 == src1),
    every other fine vertex is the midpoint of exactly one coarse edge.  Ids
    are coarse pixel ids, P_c / P_c+1 for the poles.
    """
    sc, sf = s_coarse, s_coarse + 1
    nf = 2 ** sf
    out = np.empty((n_pixels(sf), 2), dtype=np.int64)
    for k in range(5):
        for i in range(nf):
            for j in range(2 * nf):
                if i % 2 == 1 and j % 2 == 0:
                    a = b = ((i - 1) // 2, j // 2)
                elif i % 2 == 0 and j % 2 == 0:      # vertical coarse edge
                    a, b = (i // 2 - 1, j // 2), (i // 2, j // 2)
                elif i % 2 == 1 and j % 2 == 1:      # horizontal coarse edge
                    a, b = ((i - 1) // 2, (j - 1) // 2), ((i - 1) // 2, (j + 1) // 2)
                else:                                 # live-diagonal coarse edge
                    a, b = (i // 2 - 1, (j + 1) // 2), (i // 2, (j - 1) // 2)
                out[vid(sf, k, i, j)] = (source(sc, k, *a), source(sc, k, *b))
    return out
