'''Per-kernel register / stack / spill table from mod16_b200/lib/build.log (ptxas -v).'''
import re, subprocess, sys, os
log = open(os.path.join(os.path.dirname(__file__), '..', 'mod16_b200', 'lib', 'build.log')).read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", re.S)
names = [m.group(1) for m in pat.finditer(log)]
dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.split('\n')
flt = sys.argv[1] if len(sys.argv) > 1 else ''
for m, d in zip(pat.finditer(log), dem):
    d = re.sub(r'\(.*', '', d).replace('void mod16::', '')
    if flt in d:
        print('%-60s regs %3s stack %4s spill st %4s ld %4s' % (d, m.group(5), m.group(2), m.group(3), m.group(4)))
