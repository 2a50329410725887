#!/bin/bash
# masked-tile and clean-tile throughput of library variants.  usage: gpu_fill_variants.sh <variant...>
for v in "$@"; do
  export MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so
  for f in 0 0.01 0.3; do echo -n "$v "; timeout 300 python scripts/bench_fill.py $f 2>&1 | tail -1; done
  echo -n "$v bench T=128: "; python bench.py --T 128 --steps 10 --no-cpu --no-e2e --no-secondary | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['roofline']['frac'], d['clocks'])"
done
