#!/bin/bash
# Same-box A/B of the float64 entry: the main library against one variant, interleaved, twice
# each, after the float64 parity tests on the main library.  usage: gpu_f64_ab.sh <tag> <variant>
tag=$1; v=$2; out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q -k "float64 or f64 or promotion or mirror or golden or parity or layouts" > $out/pytest_$tag.log 2>&1; tail -3 $out/pytest_$tag.log
for rep in 1 2; do
  echo "--- main, run $rep";    python scripts/bench_f64.py 32 2>&1 | grep -v "equal the" | tee -a $out/f64_${tag}_main.jsonl
  echo "--- $v, run $rep"; MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so python scripts/bench_f64.py 32 2>&1 | grep -v "equal the" | tee -a $out/f64_${tag}_$v.jsonl
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader
