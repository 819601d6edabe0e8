#!/usr/bin/env python
"""SASS evidence files of profiles/ (no GPU needed):
    tools/sass_evidence.py <obj.o> <title> <pattern> [<pattern> ...] > profiles/sass_<kernel>.txt
For every kernel whose demangled name contains a pattern: architecture flags, static instruction
counts, the bulk-copy-engine / mbarrier / MUFU.RSQ64H / barrier instructions, and the loops that
hold FP64-pipe instructions (tools/sass_loops.py)."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import sass_loops  # noqa: E402

MARK = re.compile(r'UBLKCP|SYNCS|MUFU\.RSQ64H|BAR\.SYNC|UTMA|UTC|LDL|STL')


def main():
    path, title, pats = sys.argv[1], sys.argv[2], sys.argv[3:]
    print('# {0}: SASS evidence'.format(title))
    print('# cuobjdump -sass {0} (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a), filtered by '
          'tools/sass_evidence.py'.format(os.path.relpath(path, os.path.dirname(HERE))))
    sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r'arch = (\S+)', sass)) | set(re.findall(r'EF_CUDA_SM\d+', sass)))
    funcs = sass_loops.functions(path)
    for name, ins in funcs.items():
        if not any(p in name for p in pats):
            continue
        print('\n## ' + name)
        print('   arch flags of the object: ' + ' '.join(arch))
        c = sass_loops.mix(ins)
        print('   instruction counts (static, {0} in all): '.format(len(ins)) +
              ', '.join('{0} {1}'.format(k, v) for k, v in c.most_common(16)))
        marked = [(a, t) for a, t in ins if MARK.search(t)]
        hist = collections.Counter(MARK.search(t).group(0) for _, t in marked)
        print('   bulk-copy engine / mbarrier / MUFU.RSQ64H / barrier / local-memory instructions: ' +
              ', '.join('{0} {1}'.format(k, v) for k, v in sorted(hist.items())))
        for a, t in marked[:20]:
            print('    /*{0:04x}*/  {1}'.format(a, t))
    for p in pats:
        print("\n## loops with FP64-pipe instructions (tools/sass_loops.py {0} '{1}')".format(
            os.path.relpath(path, os.path.dirname(HERE)), p))
        out = subprocess.run([sys.executable, os.path.join(HERE, 'sass_loops.py'), path, p],
                             capture_output=True, text=True).stdout
        for line in out.splitlines():
            if line.startswith('==') or 'whole kernel' in line or re.search(r'FP64 pipe [1-9]', line):
                print(line)


if __name__ == '__main__':
    main()
