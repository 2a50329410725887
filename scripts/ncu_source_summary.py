#!/usr/bin/env python
'''Summarises the SASS source page of an ncu report: instruction mix by opcode, executed
counts and stall samples.  usage: ncu_source_summary.py <report.ncu-rep> [kernel-index]'''
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name"')
args = [x for x in sys.argv[2:] if not x.startswith('--')]
which = int(args[0]) if args else len(blocks) - 1
blk = '"Kernel Name"' + blocks[which]
rows = list(csv.reader(io.StringIO(blk)))
print(rows[0][1])
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
op_ex = collections.Counter(); op_smp = collections.Counter(); stalls = collections.Counter()
tot = 0; tots = 0
body = []
for r in rows[2:]:
    if len(r) <= iex: continue
    try: ex = int(r[iex]); smp = int(r[ismp])
    except ValueError: continue
    op = r[isrc].split()
    op = [o for o in op if not o.startswith('@')][0].rstrip(';')
    op_ex[op.split('.')[0]] += ex; op_smp[op.split('.')[0]] += smp
    tot += ex; tots += smp
    for i in stall_cols:
        try: stalls[hdr[i]] += int(r[i])
        except ValueError: pass
    body.append((r[ia], r[isrc].strip(), ex, smp))
print('total warp instructions executed', tot, 'samples', tots)
for op, n in op_ex.most_common(40):
    print('%-12s %12d %6.2f%%  samples %6.2f%%' % (op, n, 100.0 * n / tot, 100.0 * op_smp[op] / max(tots, 1)))
print('stalls:', [(k, round(100.0 * v / max(tots, 1), 1)) for k, v in stalls.most_common(10)])
if '--listing' in sys.argv:
    for a, s, ex, smp in body:
        print('%s %10d %6d  %s' % (a[-5:], ex, smp, s))
