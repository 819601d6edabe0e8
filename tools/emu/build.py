#!/usr/bin/env python
"""Build the CPU emulation of libicqLaplaceMatricesB200 (tests only):
tools/emu/_build/libicqb_emu.so from the translated sources of icqsol_b200/csrc."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'icqsol_b200', 'csrc')
OUT = os.path.join(HERE, '_build')
LIB = os.path.join(OUT, 'libicqb_emu.so')

sys.path.insert(0, HERE)
from translate import translate  # noqa: E402


def build(force=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    deps = srcs + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '*.cpp')) + \
        glob.glob(os.path.join(HERE, 'include', '*.h')) + [os.path.join(ROOT, 'include', 'icqb.h'),
                                                           os.path.join(HERE, 'translate.py')]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    gen = os.path.join(OUT, 'src', 'icqsol_b200', 'csrc')
    os.makedirs(gen, exist_ok=True)
    os.makedirs(os.path.join(OUT, 'src', 'include'), exist_ok=True)
    # keep the relative layout: the sources include "../../include/icqb.h"
    with open(os.path.join(ROOT, 'include', 'icqb.h')) as f, \
            open(os.path.join(OUT, 'src', 'include', 'icqb.h'), 'w') as g:
        g.write(f.read())
    cpps = []
    for path in srcs + glob.glob(os.path.join(CSRC, '*.cuh')):
        name = os.path.basename(path)
        dst = os.path.join(gen, name if name.endswith('.cuh') else name[:-3] + '.cpp')
        with open(path) as f, open(dst, 'w') as g:
            g.write(translate(f.read()))
        if dst.endswith('.cpp'):
            cpps.append(dst)
    cmd = ['g++', '-std=c++17', '-O1', '-g', '-fPIC', '-shared', '-ffp-contract=off', '-DICQB_EMU', '-D__CUDACC__',
           '-D_GNU_SOURCE', '-U_FORTIFY_SOURCE', '-D_FORTIFY_SOURCE=0', '-Wno-unused-function', '-Wno-unused-variable',
           '-I', os.path.join(HERE, 'include'), '-o', LIB] + cpps + \
        [os.path.join(HERE, 'emu_runtime.cpp'), '-ldl']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
