#!/bin/bash
# Interleaved A/B of two library variants on the tile-year (several rounds, same box).
# usage: gpu_ab_repeat.sh <rounds> <variantA> <variantB> [T]
R=$1; A=$2; B=$3; T=${4:-365}
for i in $(seq 1 $R); do
  for v in $A $B; do
    MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so python bench.py --steps 10 --warmup 3 --T $T --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
