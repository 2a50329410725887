#!/usr/bin/env python
'''Opcode histogram of the largest loop (backward branch span) of a kernel in an object file.
usage: loopstat.py <mangled-substring> <obj> [--dump file]'''
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
    best = None
    for a, t in ins:
        m = re.search(r'BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            if best is None or a - lo > best[1] - best[0]:
                best = (lo, a)
    body = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
    c = collections.Counter()
    for a, t in body:
        p = t.split()
        c[(p[1] if p[0].startswith('@') else p[0]).split('.')[0]] += 1
    print(name, 'total', len(ins), 'loop', len(body), hex(best[0]), hex(best[1]))
    print(' '.join('%s:%d' % kv for kv in c.most_common(40)))
    if '--dump' in sys.argv:
        open(sys.argv[sys.argv.index('--dump') + 1], 'w').write(
            '\n'.join('%05x %s' % (a, t) for a, t in body))
