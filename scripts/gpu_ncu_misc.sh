#!/bin/bash
# ncu launch lists (gpu__time_duration + DRAM bytes + instruction counts) of the kernels next to
# the headline one: K3 (time aggregates), the float64 fast path, the annual-sum epilogue, the
# storage-format variants.  usage: gpu_ncu_misc.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
python scripts/bench_aggregate.py > $out/k3_plain_$tag.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:time_aggregate -c 4 --csv --log-file $out/k3_ncu_$tag.csv python scripts/bench_aggregate.py > /dev/null 2>&1
echo "k3 exit $?"
python scripts/bench_misc.py > $out/misc_plain_$tag.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:'forward_fast64|forward_pipe|elementwise' -c 40 --csv --log-file $out/misc_ncu_$tag.csv python scripts/bench_misc.py > /dev/null 2>&1
echo "misc exit $?"; cat $out/misc_plain_$tag.log
