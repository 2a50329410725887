#!/bin/bash
# A/B of library variants on the resident tile in a storage format. usage: gpu_ab_format.sh <rounds> <format> <T> <variant...>
R=$1; F=$2; T=$3; shift; shift; shift
for i in $(seq 1 $R); do
  for v in "$@"; do
    MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so python bench.py --format $F --steps 10 --warmup 3 --T $T --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v $F', round(d['value']/1e9,2), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
