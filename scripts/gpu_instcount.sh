#!/bin/bash
# Dynamic instruction count and duration of forward_pipe_kernel for library variants.
# usage: gpu_instcount.sh <T> <variant...>
T=$1; shift
for v in "$@"; do
  echo -n "$v: "
  MOD16_B200_LIB=$PWD/mod16_b200/lib/variants/lib$v.so timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none -k regex:forward_pipe -s 4 -c 1 python bench.py --steps 2 --warmup 3 --T $T --no-cpu --no-e2e --no-secondary 2>&1 | grep -E "inst_executed|duration|issue_active|grid_size" | awk '{printf "%s %s  ", $1, $NF} END {print ""}'
done
