"""Compile the SIMT kernel sources with g++ against cpu_emu.h -> tools/cpu_emu/liblnst_emu.so.

TEST TOOLING ONLY (see cpu_emu.h).  tcgen05/TMA sources (conv_tc.cu) are not emulated.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'neural-flow-style_b200', 'csrc')
OUT = os.path.join(HERE, 'liblnst_emu.so')
SOURCES = ['splat.cu', 'field.cu', 'render.cu', 'lossnet.cu', 'optim.cu', 'gather.cu', 'reg.cu', 'graphnet.cu']


def build(force=False):
    if os.environ.get('LNST_EMU_LIB'):          # e.g. an AddressSanitizer build of the same sources (tools/cpu_emu/README)
        return os.environ['LNST_EMU_LIB']
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(HERE, 'cpu_emu.h'), os.path.join(CSRC, 'common.cuh'),
                   os.path.join(ROOT, 'include', 'lnst_b200.h')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(d) for d in deps):
        return OUT
    objs = []
    for s in srcs:
        o = os.path.join(HERE, os.path.basename(s) + '.emu.o')
        subprocess.check_call(['g++', '-x', 'c++', '-std=c++17', '-O2', '-fPIC', '-DLNST_CPU_EMU',
                               '-Wno-unknown-pragmas', '-I', HERE, '-c', s, '-o', o])
        objs.append(o)
    subprocess.check_call(['g++', '-shared', '-o', OUT] + objs)
    return OUT


if __name__ == '__main__':
    print(build(force=True))
