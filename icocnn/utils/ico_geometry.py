from geniconet_b200.ico_geometry import get_ico_faces, get_icosahedral_grid, pad_index_map  # noqa: F401
