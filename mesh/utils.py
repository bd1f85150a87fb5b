from geniconet_b200.mesh_utils import (compute_vertex_normals, compute_laplacian_batch, compute_laplacian,  # noqa: F401
                                       compute_adjacency_matrix_sparse)
