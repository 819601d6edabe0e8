#!/usr/bin/env python
"""Instruction mix of every loop (backward branch) of the kernels whose demangled name
contains a pattern:  tools/sass_loops.py <lib.so|obj.o> <pattern>
Used to count the FP64-pipe instructions per kernel evaluation without a GPU
(cuobjdump -sass)."""
import collections
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    cur, res = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = res.setdefault(name, [])
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    return res


def mix(instrs):
    c = collections.Counter()
    for _, text in instrs:
        op = re.sub(r'^@!?U?P\d+\s+', '', text).split()[0]
        c[op.split('.')[0]] += 1
    return c


FP64 = ('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX')


def main():
    path, pat = sys.argv[1], sys.argv[2]
    for name, ins in functions(path).items():
        if pat not in name:
            continue
        print('==', name[:140], '-- {0} instructions'.format(len(ins)))
        total = mix(ins)
        print('   whole kernel:', ', '.join('{0} {1}'.format(k, v) for k, v in total.most_common(14)))
        for k, (addr, text) in enumerate(ins):
            m = re.search(r'\bBRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)', text)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt >= addr:
                continue
            body = [x for x in ins if tgt <= x[0] <= addr]
            c = mix(body)
            fp64 = sum(c[o] for o in FP64)
            print('   loop 0x{0:04x}..0x{1:04x}: {2} instr, FP64 pipe {3} (DFMA {4} DADD {5} DMUL {6}), MUFU {7}, '
                  'LDS {8}, LDG {9}, LDC {10}'.format(tgt, addr, len(body), fp64, c['DFMA'], c['DADD'], c['DMUL'],
                                                     c['MUFU'], c['LDS'], c['LDG'], c['LDC'] + c['ULDC']))


if __name__ == '__main__':
    main()
