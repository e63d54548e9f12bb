"""Build libvarsep_sm100a.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m spatiotemporal_variable_separation_b200.csrc.build [--force]
"""
import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, 'libvarsep_sm100a.so')
OBJ_DIR = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(HERE, '*.cu')))
    deps = srcs + glob.glob(os.path.join(HERE, '*.cuh')) + [os.path.join(PKG, '..', 'include', 'varsep.h')]
    stamp = os.path.join(OBJ_DIR, 'stamp')
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)

    def cc(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        r = subprocess.run([NVCC] + FLAGS + ['-c', src, '-o', obj], capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f'nvcc failed on {src}')
        with open(obj + '.ptxas.log', 'w') as f:
            f.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, srcs))
    r = subprocess.run([NVCC, '-shared', '-o', LIB] + objs + ['-lcudart'], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    open(stamp, 'w').write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
