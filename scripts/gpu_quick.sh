#!/bin/bash
# GPU parity tests + short bench + dynamic instruction count of the forward kernel.
# usage: gpu_quick.sh <tag> [T]
tag=${1:-q}; T=${2:-128}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -4 $out/pytest_$tag.log
B="python bench.py --steps 10 --warmup 3 --T $T --no-cpu --no-e2e --no-secondary"
timeout 600 $B > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('$out/bench_$tag.json').read().strip().splitlines()[-1])
    print('%.1f G/s' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks'])
except Exception as e:
    print('no json', e); print(open('$out/bench_$tag.err').read()[-2000:])
PY
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:forward_pipe -s 4 -c 1 python bench.py --steps 2 --warmup 3 --T $T --no-cpu --no-e2e --no-secondary 2>&1 | grep -E "inst_executed|duration|issue_active" 
