#!/bin/bash
# A/B of library variants on a reduced resident block.  usage: gpu_variants.sh <tag> <T> <variant...>
tag=$1; T=$2; shift; shift
out=gpurun_out; mkdir -p $out
for v in "$@"; do
  MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so timeout 600 python bench.py --steps 10 --warmup 3 --T $T --no-cpu --no-e2e --no-secondary > $out/bench_${tag}_$v.json 2> $out/bench_${tag}_$v.err; echo "bench $v exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_${tag}_$v.json').read().strip().splitlines()[-1])
    print('$v', '%.1f G/s' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks'])
except Exception as e:
    print('no json', e); print(open('$out/bench_${tag}_$v.err').read()[-2000:])
PY
done
