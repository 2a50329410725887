#!/usr/bin/env python
'''Writes profiles/<tag>_launches_summary.txt and profiles/<tag>_<kernel>_full.txt from the ncu
outputs of scripts/gpu_check.sh / gpu_ncu.sh under gpurun_out/.
usage: make_profile_summary.py <tag> <launches.csv|-> <report.ncu-rep> <units per launch> "<command>" [bytes per unit]
(bytes per unit: 48 for the float32 sum, the default; 96 for the float64 entry on float64 drivers)'''
import collections, csv, io, subprocess, sys
tag, launches, rep, units, cmd = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
bpu = float(sys.argv[6]) if len(sys.argv) > 6 else 48.0
if launches != '-':
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); ig = hdr.index('Grid Size')
    agg = collections.OrderedDict(); ours = []
    for r in rows[1:]:
        name = r[ik].split('(')[0][:90]; t = float(r[iv])
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
        if 'mod16::' in r[ik]:
            ours.append((r[0], r[ik][:70], r[ig], t))
    tot = sum(v[1] for v in agg.values())
    tot_ours = sum(o[3] for o in ours)
    with open('profiles/%s_launches_summary.txt' % tag, 'w') as f:
        f.write('ncu --metrics gpu__time_duration.sum --clock-control none -c 400: %s\n' % cmd)
        f.write('(first 400 launches; at:: kernels are the synthetic-driver generator, outside the timed region)\n\n')
        f.write('%-92s %6s %14s %7s\n' % ('kernel', 'count', 'total ns', 'share'))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-92s %6d %14.0f %6.2f%%\n' % (k, c, t, 100 * t / tot))
        f.write('\nlaunches of this repo\'s kernels (id, name, grid, ns, share of this repo\'s kernel time):\n')
        for o in ours:
            f.write('%s %s %s %.0f %.2f%%\n' % (o + (100 * o[3] / tot_ours,)))
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, un, r = rows[0], rows[1], rows[-1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct']
kname = r[hdr.index('Kernel Name')]
short = kname.split('<')[0].split('(')[0].split('::')[-1].replace('void ', '')
with open('profiles/%s_%s_full.txt' % (tag, short), 'w') as f:
    f.write('ncu --set full --clock-control none --import-source on: %s\n' % cmd)
    f.write('kernel: %s\n%.4g pixel-timesteps per launch (algorithmic bytes %.4g GB at %g B each)\n\n' % (kname, units, units * bpu / 1e9, bpu))
    vals = {}
    for k in keys:
        if k in hdr:
            i = hdr.index(k); vals[k] = r[i]
            f.write('%-70s %-16s %s\n' % (k, un[i], r[i]))
    try:
        ms = float(vals['gpu__time_duration.sum']); 
        if un[hdr.index('gpu__time_duration.sum')] == 'us': ms /= 1e3
        f.write('\n=> %.1f G units/s under the profiler; %.1f warp-instructions per 32 units\n' % (
            units / ms / 1e6, float(vals['smsp__inst_executed.sum']) / (units / 32)))
    except Exception as e:
        f.write('(derived figures unavailable: %s)\n' % e)
    f.write('\n')
    f.write(subprocess.run(['python', 'scripts/ncu_source_summary.py', rep], capture_output=True, text=True).stdout)
print(open('profiles/%s_%s_full.txt' % (tag, short)).read()[:2500])
