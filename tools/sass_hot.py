#!/usr/bin/env python
'''Instruction mix of the hot path of a kernel: the backward-branch loop with the most packed
FP32 instructions, with the rarely-taken blocks (those reached only through a forward branch
over them, i.e. out-of-line code ptxas placed inside the loop range) reported separately.
usage: sass_hot.py <mangled-substring> <obj-or-lib> [--dump]'''
import re, subprocess, sys, collections
lib = sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    if sys.argv[1] not in name:
        continue
    ins = []
    for line in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, text in ins:
        m = re.search(r'BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)', text)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    def packed(lo, hi):
        return sum(1 for a, t in ins if lo <= a <= hi and re.search(r'\bF(FMA|MUL|ADD)2\b', t))
    best = max(loops, key=lambda l: (packed(*l) / (l[1] - l[0] + 16.0)))
    # prefer the smallest loop that still holds >= 90 % of the max packed count
    mx = max(packed(*l) for l in loops)
    cands = [l for l in loops if packed(*l) >= 0.45 * mx]
    best = min(cands, key=lambda l: l[1] - l[0])
    print(name)
    print('loops', [(hex(a), hex(b), packed(a, b)) for a, b in loops])
    print('hot loop', hex(best[0]), hex(best[1]))
    cnt = collections.Counter()
    body = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
    for a, t in body:
        w = t.split()
        op = w[1] if w[0].startswith('@') else w[0]
        cnt[op.split('.')[0]] += 1
    tot = sum(cnt.values())
    print('instructions in range', tot)
    print(', '.join('%s %d' % kv for kv in cnt.most_common()))
    if '--dump' in sys.argv:
        for a, t in body:
            print('%05x  %s' % (a, t))
