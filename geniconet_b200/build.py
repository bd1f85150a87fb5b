"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no torch involved).

    python -m geniconet_b200.build            # -> geniconet_b200/libgeniconet_b200.so
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgeniconet_b200.so')
STAMP = os.path.join(HERE, 'csrc', '.build_stamp')
SOURCES = ['gin_api.cu', 'gin_host.cpp']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC,-O3,-Wall', '-shared',
              '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found: the geniconet_b200 CUDA library cannot be built')


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in sorted(os.listdir(root)):
            if f.endswith(('.cu', '.cuh', '.cpp', '.h')):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    """True when the built library carries the digest of the sources and flags it would be built from now."""
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == _digest()
    except OSError:
        return False


def build(force=False, verbose=False):
    import fcntl
    with open(os.path.join(HERE, '.build_lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)                 # one builder at a time (ranks of one job, parallel test workers)
        dig = _digest()
        if not force and is_current():
            return LIB
        tmp = LIB + '.tmp.%d' % os.getpid()
        cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', tmp]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(os.path.join(CSRC, '.build_log.txt'), 'w') as f:
            f.write(' '.join(cmd) + '\n' + log)
        if res.returncode != 0:
            sys.stderr.write(log)
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError('nvcc failed building libgeniconet_b200.so')
        os.replace(tmp, LIB)                             # readers never see a half-written library
        if verbose:
            print(log)
        with open(STAMP, 'w') as f:
            f.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
