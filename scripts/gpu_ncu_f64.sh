#!/bin/bash
# ncu --set full capture of the float64 fast path.  usage: gpu_ncu_f64.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
B64="python bench.py --dtype float64 --T 16 --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary"
timeout 600 $B64 > $out/plain_f64_$tag.log 2>&1 && tail -c 400 $out/plain_f64_$tag.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'forward_(fast|canon)64' -s 4 -c 1 -f -o $out/prof_f64_$tag $B64 > $out/ncu_f64_$tag.log 2>&1
echo "ncu f64 exit $?"
