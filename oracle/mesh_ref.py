"""ORACLE (test infrastructure, NOT product code) -- mesh helpers on CPU.

Restates what the reference imports as ``mesh.utils`` (/root/reference/
losses.py:7,39,54,57; generate.py:13,152,197).  The module lives in the
un-vendored github.com/hrdkjain/PythonFunctions (README.md:22), absent here.

  compute_vertex_normals   follows the only in-tree statement of the maths,
                           generate.py:20-43 (area weighted, eps=1e-10 clip) --
                           the training targets were produced by it (:194).
  compute_laplacian(_batch) uniform graph Laplacian  L(v)_i = mean_{j in N(i)} v_j - v_i
                           (sign / normalisation UNPINNED)

    parity unpinned.
"""
import torch


def compute_vertex_normals(vertices, faces, eps=1e-10):
    """vertices [B,V,3] (or [V,3]), faces [F,3] int64 -> unit normals, same shape."""
    squeeze = vertices.dim() == 2
    v = vertices.unsqueeze(0) if squeeze else vertices
    faces = faces.long()
    v0, v1, v2 = v[:, faces[:, 0]], v[:, faces[:, 1]], v[:, faces[:, 2]]
    fn = torch.cross(v1 - v0, v2 - v0, dim=2)                 # area-weighted face normals
    vn = torch.zeros_like(v)
    vn = vn.index_add(1, faces[:, 0], fn)
    vn = vn.index_add(1, faces[:, 1], fn)
    vn = vn.index_add(1, faces[:, 2], fn)
    mag = vn.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(eps)
    vn = vn / mag
    return vn[0] if squeeze else vn


def compute_adjacency_matrix_sparse(n_vertices, faces):
    """Symmetric 0/1 vertex adjacency as a torch sparse COO float32 [V,V]."""
    faces = faces.long()
    e = torch.cat((faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]), dim=0)
    e = torch.cat((e, e.flip(1)), dim=0)
    key = torch.unique(e[:, 0] * int(n_vertices) + e[:, 1])
    idx = torch.stack((key // int(n_vertices), key % int(n_vertices)))
    return torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (int(n_vertices), int(n_vertices)), check_invariants=False).coalesce()


def compute_laplacian(vertices, adj):
    """vertices [V,3] -> [V,3]."""
    a = adj.to(vertices.dtype)
    deg = torch.sparse.sum(a, dim=1).to_dense().unsqueeze(-1)
    return torch.sparse.mm(a, vertices) / deg - vertices


def compute_laplacian_batch(vertices, adj):
    """vertices [B,V,3] -> [B,V,3]."""
    B, V, D = vertices.shape
    a = adj.to(vertices.dtype)
    deg = torch.sparse.sum(a, dim=1).to_dense().view(1, V, 1)
    flat = vertices.permute(1, 0, 2).reshape(V, B * D)
    out = torch.sparse.mm(a, flat).reshape(V, B, D).permute(1, 0, 2)
    return out / deg - vertices
