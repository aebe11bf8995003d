"""Builds libcmr_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library is plain CUDA C++ behind ``extern "C"`` entry points (include/cmr_b200.h);
it has no torch dependency and is loaded with ctypes by ``_lib.py``.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libcmr_b200.so')
STAMP_PATH = os.path.join(HERE, 'libcmr_b200.so.stamp')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared',
]
NVCC_FLAGS += os.environ.get('CMR_EXTRA_NVCC_FLAGS', '').split()   # (-D... for A/B builds)
if os.environ.get('CMR_CONV_INSTRUMENT') == '1':     # measurement build (tools/conv_shape_bench.py)
    NVCC_FLAGS.append('-DCMR_CONV_INSTRUMENT=1')


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return 'nvcc'


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + \
        [os.path.join(HERE, '..', 'include', 'cmr_b200.h')]
    for p in files:
        with open(p, 'rb') as f:
            h.update(os.path.basename(p).encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    """The library exists, its stamp holds the digest of the present sources, and the stamp
    was written together with the library (a restored or copied stamp next to a library
    built from other sources must not count)."""
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    if abs(os.path.getmtime(STAMP_PATH) - os.path.getmtime(LIB_PATH)) > 120:
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns its path."""
    if not force and is_current():
        return LIB_PATH
    objs = []
    procs = []
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != '-shared']
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        cmd = [_nvcc()] + compile_flags + (['-Xptxas', '-v'] if verbose else []) + \
            ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + out)
        if verbose:
            print(out)
    cmd = [_nvcc(), '-shared', '-o', LIB_PATH] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + ' '.join(cmd) + '\n' + res.stdout)
    with open(STAMP_PATH, 'w') as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
