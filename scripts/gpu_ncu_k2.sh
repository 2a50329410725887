#!/bin/bash
# ncu launch list + full capture of the K2 hot loop.  usage: gpu_ncu_k2.sh <tag>
tag=${1:-k2}
out=gpurun_out; mkdir -p $out
B="python bench.py --workload calibration --draws 16384 --steps 1 --warmup 1 --cpu-draws 0"
timeout 600 $B > $out/plain_k2_$tag.log 2>&1 && cat $out/plain_k2_$tag.log &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $out/launches_k2_$tag.csv $B > $out/ncu_list_k2_$tag.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:objective_fast -s 1 -c 1 \
    -f -o $out/prof_k2_$tag $B > $out/ncu_full_k2_$tag.log 2>&1
echo "ncu full exit $?"
