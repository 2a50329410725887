#!/bin/bash
# ncu --set full captures of the K1 variants: float64 fast path, PACKED storage format.
# usage: gpu_ncu_variants.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
B64="python bench.py --dtype float64 --T 16 --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary"
timeout 600 $B64 > $out/plain_f64_$tag.log 2>&1 && tail -c 400 $out/plain_f64_$tag.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_fast64 -s 4 -c 1 -f -o $out/prof_f64_$tag $B64 > $out/ncu_f64_$tag.log 2>&1
echo "ncu f64 exit $?"
BP="python bench.py --format packed --T 128 --steps 2 --warmup 3 --no-cpu"
timeout 600 $BP > $out/plain_packed_$tag.log 2>&1 && tail -c 400 $out/plain_packed_$tag.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_pipe -s 4 -c 1 -f -o $out/prof_packed_$tag $BP > $out/ncu_packed_$tag.log 2>&1
echo "ncu packed exit $?"
