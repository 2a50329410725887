#!/usr/bin/env python
'''Opcode histogram of the hot path of forward_pipe_kernel<SEM,OUT,EXTRA> in an object file:
the main time loop (backward branch spanning the most MUFU), minus the inline rare blocks --
ranges skipped by a forward branch over more than 64 instructions.
usage: hotpath.py <mangled-substring> <obj> [--dump]'''
import re, subprocess, sys, collections
txt = subprocess.run(['cuobjdump', '-sass', sys.argv[2]], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    if sys.argv[1] not in name:
        continue
    ins = []
    for line in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    idx = dict((a, k) for k, (a, t) in enumerate(ins))
    loops = []
    for a, t in ins:
        m = re.search(r'BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            loops.append((sum(1 for b, u in ins if lo <= b <= a and 'MUFU' in u), -(a - lo), lo, a))
    # the innermost loop holding a full pair of steps (38 MUFU for the static sum)
    mufu = max(l[0] for l in loops)
    cands = [l for l in loops if l[0] >= 30]
    _, _, lo, hi = max(cands, key=lambda l: l[1])
    path, k = [], idx[lo]
    while ins[k][0] <= hi:
        a, t = ins[k]
        path.append((a, t))
        m = re.search(r'\bBRA\S*\s+(?:(\S+),\s+)?0x([0-9a-f]+)', t)
        if m:
            tgt = int(m.group(2), 16)
            if tgt > a and tgt <= hi and (tgt - a) // 16 > 64 and tgt in idx:
                k = idx[tgt]        # a long forward skip: the skipped block is rare code
                continue
        k += 1
    cnt = collections.Counter()
    for a, t in path:
        w = t.split()
        op = w[1] if w[0].startswith('@') else w[0]
        cnt[op.split('.')[0]] += 1
    print(name)
    print('loop %x..%x: hot path %d instructions per pair of steps (%.1f warp-instructions per 32 units)' % (
        lo, hi, len(path), len(path) / 2.0))
    print(', '.join('%s %d' % kv for kv in cnt.most_common()))
    if '--dump' in sys.argv:
        for a, t in path:
            print('%05x  %s' % (a, t))
