"""Build the native library in-tree: ``neural-flow-style_b200/lnst/liblnst_b200.so``.

    python neural-flow-style_b200/build.py            # sm_100a, -O3 -lineinfo
    python neural-flow-style_b200/build.py --ptxas-v  # also print registers/spills per kernel

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'lnst', 'liblnst_b200.so')
STAMP = OUT + '.stamp'

SOURCES = ['splat.cu', 'field.cu', 'render.cu', 'lossnet.cu', 'optim.cu', 'gather.cu', 'reg.cu', 'graphnet.cu', 'conv_tc.cu', 'tiles_tma.cu']


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ['../../include/lnst_b200.h']
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    return h.hexdigest()


def build(force=False, ptxas_v=False, verbose=True):
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '--shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '-cudart', 'static',
           '-o', OUT] + _sources()
    if ptxas_v:
        cmd[1:1] = ['-Xptxas', '-v']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(dig)
    return OUT


if __name__ == '__main__':
    build(force=True, ptxas_v='--ptxas-v' in sys.argv)
