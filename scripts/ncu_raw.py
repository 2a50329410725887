#!/usr/bin/env python
'''Key raw metrics of the last kernel in an ncu report.  usage: ncu_raw.py <report> [regex]'''
import csv, io, re, subprocess, sys
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r'gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__throughput.avg.pct|registers_per_thread|warps_active.avg.pct|smsp__inst_executed.sum$|issue_active.avg.pct|pipe_fma_cycles_active.avg.pct|inst_executed_pipe_xu|pipe_alu_cycles_active.avg.pct|sm__cycles_elapsed.avg$|launch__grid_size|occupancy_limit|launch__waves|smsp__warp_issue_stalled.*_per_warp_active|smsp__average_warp.*per_issue|shared_mem_per_block|inst_executed_pipe_(fma|alu|fp32|xu|lsu|uniform|cbu|adu)')
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:100])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print('  %-86s %-12s %s' % (h, units[i], r[i]))
