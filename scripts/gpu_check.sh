#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, launch list, one full ncu capture.
# usage: scripts/gpu_check.sh <tag>
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi > $out/smi_$tag.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" | tee -a $out/pytest_$tag.log
timeout 300 python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; echo "smoke exit $?" | tee -a $out/smoke_$tag.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
tail -3 $out/bench_$tag.json
B="python bench.py --steps 2 --warmup 3 --T 128 --no-cpu --no-e2e --no-secondary"
timeout 600 $B > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches_$tag.csv $B > $out/ncu_list_$tag.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_pipe -s 4 -c 1 \
    -f -o $out/prof_$tag $B > $out/ncu_full_$tag.log 2>&1
echo "ncu full exit $?"
