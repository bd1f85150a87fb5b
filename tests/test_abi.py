"""The C-ABI library loads on a machine without a GPU and exports every symbol include/geniconet_b200.h declares."""
import ctypes
import os
import re

from geniconet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'geniconet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gin_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_host_entry_points_without_gpu():
    assert _lib.lib.gin_version() >= 100
    assert _lib.lib.gin_index_map_len(5) == 5 * 34 * 66
    assert _lib.lib.gin_ico_faces_len(5) == 20 * 4 ** 5 * 3
    assert _lib.lib.gin_hexconv_packed_bytes(128, 64) == 84 * 128 * 64
    assert _lib.lib.gin_index_map(5, None) != 0 and b'bad argument' in _lib.lib.gin_last_error()
    assert _lib.lib.gin_plan_bytes(1, 5, 1, 7) == 0
    assert _lib.lib.gin_p2p_ws_bytes(36, 5) > 36 * 10242 * 48


def test_library_is_sm100a_native():
    """The shipped cubin must be sm_100a and contain the tcgen05 / TMEM instructions (UTC*MMA, LDTM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    assert re.search(r'UTC[A-Z]*MMA', sass) and 'LDTM' in sass
