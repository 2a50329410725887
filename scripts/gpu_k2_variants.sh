#!/bin/bash
# A/B of library variants on the K2 calibration config.  usage: gpu_k2_variants.sh <tag> <draws> <variant...>
tag=$1; D=$2; shift; shift
out=gpurun_out; mkdir -p $out
for v in "$@"; do
  MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so timeout 600 python bench.py --workload calibration --draws $D --steps 3 --cpu-draws 1 > $out/k2_${tag}_$v.json 2> $out/k2_${tag}_$v.err; echo "k2 $v exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('$out/k2_${tag}_$v.json').read().strip().splitlines()[-1])
    print('$v', '%.1f G/s' % (d['value']/1e9), 'ms', d['ms_per_step'], 'rmsd diff', d['cpu_baseline'] and d['cpu_baseline']['rmsd_max_abs_diff'])
except Exception as e:
    print('no json', e); print(open('$out/k2_${tag}_$v.err').read()[-2000:])
PY
done
