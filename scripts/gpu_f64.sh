#!/bin/bash
# float64 entry: parity tests, then throughput of each library variant.  usage: gpu_f64.sh <tag> [variant...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q -k "float64 or f64 or promotion or mirror or call_forms or golden or parity" > $out/pytest_$tag.log 2>&1; tail -5 $out/pytest_$tag.log
python scripts/bench_f64.py 32 2>&1 | tee $out/f64_${tag}_main.jsonl
for v in "$@"; do
  echo "--- variant $v"
  MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so python scripts/bench_f64.py 32 2>&1 | tee $out/f64_${tag}_$v.jsonl
done
