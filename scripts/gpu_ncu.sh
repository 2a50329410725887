#!/bin/bash
# ncu --set full capture of the f32 kernel variants.  usage: gpu_ncu.sh <tag> <kernel...>
tag=$1; shift
out=gpurun_out
mkdir -p $out
for k in "$@"; do
  case $k in pipe) re=forward_pipe;; direct) re=forward_vec4;; generic) re=forward_strided_fast;; *) re=forward_strided;; esac
  B="python bench.py --steps 2 --warmup 3 --T ${NCU_T:-32} --no-cpu --no-e2e --no-secondary --kernel $k"
  timeout 600 $B > $out/plain_${tag}_$k.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s 4 -c 1 \
      -f -o $out/prof_${tag}_$k $B > $out/ncu_${tag}_$k.log 2>&1
  echo "ncu $k exit $?"
done
