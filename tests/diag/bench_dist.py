"""Times gin_point_mesh_distance at the reference's evaluation size (level 5: 10242 points x 20480 faces per mesh) with CUDA
events and prints one JSON line: meshes/s, point-triangle tests/s and the oracle's (numpy float64, one core) rate beside it."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from geniconet_b200 import data as gd, ico_utils as iu          # noqa: E402
from geniconet_b200.ico_geometry import get_ico_faces            # noqa: E402
from oracle.kaolin_ref import point_to_mesh_distance as p2m_ref  # noqa: E402

level, B = 5, 16
faces = get_ico_faces(level)
verts = torch.stack([gd.synthetic_mesh(level, i)[1][:3].T for i in range(B)]).contiguous().cuda()
pts = (verts.roll(1, 0) * 1.01).contiguous()
ft = torch.from_numpy(faces).cuda()
for _ in range(3):
    iu.point_to_mesh_distance(pts, verts, ft)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
K = 10
for _ in range(K):
    iu.point_to_mesh_distance(pts, verts, ft)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / K
tests = B * verts.shape[1] * faces.shape[0]
t0 = time.perf_counter()
p2m_ref(pts[0, :256].cpu().numpy(), verts[0].cpu().numpy(), faces)
cpu_s = time.perf_counter() - t0
print(json.dumps({'op': 'point_to_mesh_distance', 'level': level, 'batch': B, 'ms': ms, 'meshes_per_s': B / ms * 1e3,
                  'tests_per_s': tests / ms * 1e3, 'oracle_tests_per_s': 256 * faces.shape[0] / cpu_s}))
