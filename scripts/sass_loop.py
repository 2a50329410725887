#!/usr/bin/env python
'''Static instruction mix of the hottest loop of a kernel in the built library.
usage: sass_loop.py <mangled-substring> [lib]'''
import re, subprocess, sys, collections
lib = sys.argv[2] if len(sys.argv) > 2 else 'mod16_b200/lib/libmod16_b200.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
for f in funcs[1:]:
    name = f.split('\n', 1)[0]
    if sys.argv[1] not in name:
        continue
    ins = []
    for line in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    # backward branches = loops; take the longest
    best = None
    for addr, text in ins:
        m = re.search(r'BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)', text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    print(name, 'instructions', len(ins), 'loop', [hex(b) for b in best] if best else None)
    if not best:
        continue
    cnt = collections.Counter()
    for addr, text in ins:
        if best[0] <= addr <= best[1]:
            t = text.split()
            op = t[1] if t[0].startswith('@') else t[0]
            cnt[op.split('.')[0]] += 1
    tot = sum(cnt.values())
    print('loop body instructions', tot, '(per pixel-timestep: %.1f)' % (tot / 4.0))
    print(', '.join('%s %d' % kv for kv in cnt.most_common()))
