"""GPU parity of the evaluation / data-path rows: point-to-mesh distance against the float64
oracle (oracle/kaolin_ref.py), output2vertices -> computeDistance, and the pinned-memory device prefetcher."""
import numpy as np
import pytest
import torch

from geniconet_b200 import data as gd
from geniconet_b200 import ico_utils as iu
from geniconet_b200.ico_geometry import get_ico_faces
from oracle.kaolin_ref import point_to_mesh_distance as p2m_ref

pytestmark = pytest.mark.gpu


def _mesh(level, idx):
    _, t = gd.synthetic_mesh(level, idx)
    return t[:3].T.contiguous()                             # [P+2, 3]


def _check_against_oracle(pts, verts, faces, d, f):
    """squared distances within fp32 rounding of the oracle's; the chosen face attains the minimum (ties may differ)."""
    d_ref, _ = p2m_ref(pts, verts, faces)
    d = d.double().cpu().numpy()
    assert np.all(np.abs(d - d_ref) <= 1e-5 * d_ref + 1e-9), np.abs(d - d_ref).max()
    f = f.cpu().numpy()
    for i in np.linspace(0, len(pts) - 1, 24).astype(int):
        d_face, _ = p2m_ref(pts[i:i + 1], verts, faces[f[i]:f[i] + 1])
        assert abs(d_face[0] - d_ref[i]) <= 1e-5 * d_ref[i] + 1e-9


def test_point_to_mesh_known_answers():
    V = torch.tensor([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1]]], dtype=torch.float32, device='cuda')
    F = torch.tensor([[0, 1, 2], [1, 3, 2]])
    pts = torch.tensor([[[0.25, 0.25, 0.5], [-1, -1, 0], [0.5, -2, 0], [0.25, 0.25, 0], [1, 1, 1]]], dtype=torch.float32, device='cuda')
    d, f = iu.point_to_mesh_distance(pts, V, F)
    assert f.dtype == torch.int64 and d.shape == (1, 5)
    assert torch.allclose(d.cpu(), torch.tensor([[0.25, 2.0, 4.0, 0.0, 0.0]]), atol=1e-7)
    assert f.cpu().tolist() == [[0, 0, 0, 0, 1]]
    # degenerate (collinear) triangle behaves like its longest edge; empty point set is a no-op
    Vd = torch.tensor([[[0, 0, 0], [1, 0, 0], [2, 0, 0]]], dtype=torch.float32, device='cuda')
    d, _ = iu.point_to_mesh_distance(torch.tensor([[[0.5, 1.0, 0.0]]], device='cuda'), Vd, torch.tensor([[0, 1, 2]]))
    assert torch.allclose(d.cpu(), torch.tensor([[1.0]]))
    d, f = iu.point_to_mesh_distance(torch.zeros(1, 0, 3, device='cuda'), Vd, torch.tensor([[0, 1, 2]]))
    assert d.shape == (1, 0) and f.shape == (1, 0)
    with pytest.raises(ValueError):
        iu.point_to_mesh_distance(pts, V, torch.tensor([[0, 1, 4]]))


@pytest.mark.parametrize('level', [1, 3])
def test_point_to_mesh_matches_oracle(level):
    faces = get_ico_faces(level)
    verts = torch.stack([_mesh(level, 0), _mesh(level, 3)])
    g = torch.Generator().manual_seed(level)
    pts = torch.stack([_mesh(level, 7) * 1.02 + 0.003, _mesh(level, 3) + 0.01 * torch.randn(verts.shape[1], 3, generator=g)])
    pts = torch.cat([pts, torch.randn(2, 37, 3, generator=g)], 1).contiguous()          # ragged sizes, far-away points too
    d, f = iu.point_to_mesh_distance(pts.cuda(), verts.cuda(), torch.from_numpy(faces))
    for b in range(2):
        _check_against_oracle(pts[b].numpy(), verts[b].numpy(), faces, d[b], f[b])


def test_point_to_mesh_full_size_properties():
    """Level 5 (10242 points x 20480 faces, the reference's evaluation size), batch of 2: a mesh's own vertices are at distance
    exactly 0; moving every point by h along any direction gives distances <= h^2; a sample of points agrees with the oracle."""
    level = 5
    faces = get_ico_faces(level)
    verts = torch.stack([_mesh(level, 1), _mesh(level, 2)]).cuda()
    ft = torch.from_numpy(faces)
    d, f = iu.point_to_mesh_distance(verts, verts, ft)
    assert float(d.max()) == 0.0
    owner = torch.from_numpy(faces.astype(np.int64)).cuda()[f[0]]                       # the reported face contains the vertex
    assert bool((owner == torch.arange(verts.shape[1], device='cuda')[:, None]).any(1).all())
    h = 0.01
    g = torch.Generator().manual_seed(0)
    dirs = torch.nn.functional.normalize(torch.randn(2, verts.shape[1], 3, generator=g), dim=-1).cuda()
    d2, _ = iu.point_to_mesh_distance(verts + h * dirs, verts, ft)
    assert float(d2.max()) <= h * h * (1 + 1e-4) and float(d2.min()) >= 0.0
    sel = torch.linspace(0, verts.shape[1] - 1, 48).long()
    pts = (verts[1] + h * dirs[1])[sel.cuda()].cpu()
    d_ref, _ = p2m_ref(pts.numpy(), verts[1].cpu().numpy(), faces)
    assert np.all(np.abs(d2[1][sel.cuda()].double().cpu().numpy() - d_ref) <= 1e-5 * d_ref + 1e-10)


def test_compute_distance_of_grid_output():
    """ico_utils.py:10-44 end to end: a [1,3,5n,2n] grid -> vertex list with averaged poles -> mean squared distance to the mesh."""
    level = 3
    x, t = gd.synthetic_mesh(level, 4)
    faces = torch.from_numpy(get_ico_faces(level))
    ref_v = t[:3].T.contiguous().cuda()
    out_v = iu.output2vertices(level, x[None].cuda())[0]
    assert out_v.shape == ref_v.shape
    assert torch.equal(out_v[:-2], ref_v[:-2])              # grid vertices pass through, poles are the 5-pixel means
    dist = iu.computeDistance(out_v, ref_v, faces, None, mode='point2mesh')
    d_ref, _ = p2m_ref(out_v.cpu().numpy(), ref_v.cpu().numpy(), faces.numpy())
    assert isinstance(dist, np.ndarray) and abs(float(dist) - d_ref.mean()) <= 1e-5 * d_ref.mean() + 1e-10
    assert iu.computeDistance(out_v, ref_v, faces, None) is None                        # default mode computes nothing (ico_utils.py:26,42)


def test_device_prefetcher_delivers_every_batch_in_order():
    level, B = 2, 3
    batches = [gd.synthetic_batch(level, 10 * i, B) for i in range(5)]
    seen = []
    for x, t in gd.DevicePrefetcher(iter(batches)):
        assert x.is_cuda and t.is_cuda
        seen.append((x * 2.0, t.clone()))                   # consume on the current stream while the next copy is in flight
    torch.cuda.synchronize()
    assert len(seen) == 5
    for (x2, t), (xh, th) in zip(seen, batches):
        assert torch.equal(x2.cpu(), xh * 2.0) and torch.equal(t.cpu(), th)
