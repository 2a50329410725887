#!/bin/bash
# K2 on the GPU: parity tests, then the objective-batch throughput at a reduced and full size.
# usage: scripts/gpu_k2.sh <tag>
tag=${1:-k2}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -15 $out/pytest_$tag.log
timeout 600 python bench.py --workload calibration --draws 8192 --steps 2 > $out/k2_cal_small_$tag.json 2> $out/k2_cal_small_$tag.err; echo "k2 small exit $?"; cat $out/k2_cal_small_$tag.json; tail -5 $out/k2_cal_small_$tag.err
timeout 900 python bench.py --workload calibration --steps 3 > $out/k2_cal_$tag.json 2> $out/k2_cal_$tag.err; echo "k2 cal exit $?"; cat $out/k2_cal_$tag.json; tail -5 $out/k2_cal_$tag.err
timeout 900 python bench.py --workload sobol --steps 3 > $out/k2_sobol_$tag.json 2> $out/k2_sobol_$tag.err; echo "k2 sobol exit $?"; cat $out/k2_sobol_$tag.json; tail -5 $out/k2_sobol_$tag.err
