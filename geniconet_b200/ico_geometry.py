"""Host-side mirror of ``icocnn.utils.ico_geometry`` (reference call sites losses.py:5,34;
generate.py:11,151; run.py:144,529), backed by the C ABI's host geometry."""
import numpy as np

from . import _lib


def get_ico_faces(subdivisions):
    """int64 [20*4^s, 3]; vertex ids in grid order with the two poles last (max()+1 == P+2)."""
    return _lib.ico_faces(int(subdivisions)).astype(np.int64)


def get_icosahedral_grid(subdivisions):
    """(vertices float32 [P+2,3] on the unit sphere, faces int64 [20*4^s,3])."""
    return _lib.ico_vertices(int(subdivisions)), get_ico_faces(subdivisions)


def pad_index_map(subdivisions):
    return _lib.index_map(int(subdivisions))
