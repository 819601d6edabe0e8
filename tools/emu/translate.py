#!/usr/bin/env python
"""Rewrite CUDA-only syntax so that g++ can compile a translation unit against
tools/emu/include/cuda_runtime.h (test infrastructure, see tools/emu/README.md):

    kernel<<<grid, block, smem, stream>>>(args);   ->  EMU_LAUNCH((kernel), grid, block, smem, stream, args);
    extern __shared__ [__align__(n)] T name[];     ->  T* name = (T*)emu::dyn_smem();
"""
import re
import sys


def split_top(text):
    """Split at top-level commas (outside (), [], {})."""
    parts, depth, cur = [], 0, ''
    for ch in text:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur.strip())
            cur = ''
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def matching_paren(text, start):
    """Index of the parenthesis closing the one at text[start]."""
    depth = 0
    for k in range(start, len(text)):
        if text[k] == '(':
            depth += 1
        elif text[k] == ')':
            depth -= 1
            if depth == 0:
                return k
    raise ValueError('unbalanced parentheses')


LAUNCH = re.compile(r'([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<')


def translate(src):
    out, pos = '', 0
    while True:
        m = LAUNCH.search(src, pos)
        if not m:
            out += src[pos:]
            break
        close = src.index('>>>', m.end())
        cfg = split_top(src[m.end():close])
        while len(cfg) < 4:
            cfg.append('0')
        lp = src.index('(', close + 3)
        if src[close + 3:lp].strip():
            raise ValueError('unexpected text between >>> and (')
        rp = matching_paren(src, lp)
        args = src[lp + 1:rp].strip()
        out += src[pos:m.start()]
        out += 'EMU_LAUNCH(({0}), {1}, {2}, {3}, {4}{5})'.format(
            m.group(1), cfg[0], cfg[1], cfg[2], cfg[3], (', ' + args) if args else '')
        pos = rp + 1
    out = re.sub(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];',
                 lambda g: '{0}* {1} = ({0}*)emu::dyn_smem();'.format(g.group(1).strip(), g.group(2)), out)
    return out


if __name__ == '__main__':
    sys.stdout.write(translate(open(sys.argv[1]).read()))
