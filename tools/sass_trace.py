#!/usr/bin/env python
'''Follows the fall-through path of a kernel's SASS from a start address: at every conditional
forward branch it stays on the fall-through side (taken only for unconditional BRA), until the
backward branch that closes the loop.  Prints the opcode histogram of that path -- the hot
path when rare code is reached through taken branches.
usage: sass_trace.py <mangled-substring> <obj-or-lib> [start-hex] [--dump]'''
import re, subprocess, sys, collections
txt = subprocess.run(['cuobjdump', '-sass', sys.argv[2]], capture_output=True, text=True).stdout
args = [a for a in sys.argv[3:] if not a.startswith('--')]
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    if sys.argv[1] not in name:
        continue
    ins = []
    for line in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr2i = dict((a, k) for k, (a, t) in enumerate(ins))
    # the hot loop: backward branch whose range holds the most MUFU, smallest such
    loops = []
    for a, t in ins:
        m = re.search(r'BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            loops.append((lo, a, sum(1 for b, u in ins if lo <= b <= a and 'MUFU' in u)))
    mx = max(l[2] for l in loops)
    lo, hi, _ = min((l for l in loops if l[2] >= 0.3 * mx and l[2] > 0), key=lambda l: l[1] - l[0])
    start = int(args[0], 16) if args else lo
    k = addr2i[start]
    path = []
    seen = set()
    while True:
        a, t = ins[k]
        if a in seen:
            break
        seen.add(a)
        path.append((a, t))
        m = re.search(r'\bBRA\S*\s+(?:(\S+),\s+)?0x([0-9a-f]+)', t)
        w = t.split()
        pred = w[0].startswith('@')
        if m:
            tgt = int(m.group(2), 16)
            uncond = not pred and m.group(1) is None
            if tgt <= a and tgt == start:
                break
            if uncond and tgt in addr2i:
                k = addr2i[tgt]
                continue
        if re.match(r'(@\S+\s+)?EXIT', t) and not pred:
            break
        k += 1
        if k >= len(ins):
            break
    cnt = collections.Counter()
    for a, t in path:
        w = t.split()
        op = w[1] if w[0].startswith('@') else w[0]
        cnt[op.split('.')[0]] += 1
    print(name)
    print('loop %x..%x, path from %x: %d instructions' % (lo, hi, start, len(path)))
    print(', '.join('%s %d' % kv for kv in cnt.most_common()))
    if '--dump' in sys.argv:
        for a, t in path:
            print('%05x  %s' % (a, t))
