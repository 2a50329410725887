#!/bin/bash
# A/B of the f32 kernel variants on a reduced resident block + GPU tests.  usage: gpu_ab.sh <tag>
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_$tag.log
for k in pipe; do
  timeout 600 python bench.py --steps 10 --warmup 3 --T 64 --no-cpu --no-e2e --no-secondary --kernel $k > $out/bench_${tag}_$k.json 2> $out/bench_${tag}_$k.err; echo "bench $k exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_${tag}_$k.json').read().strip().splitlines()[-1])
    print('$k', '%.1f G/s' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks'])
except Exception as e:
    print('no json', e); print(open('$out/bench_${tag}_$k.err').read()[-2000:])
PY
done
