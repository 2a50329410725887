#!/bin/bash
# The driver's multi-GPU bench invocation.  usage: gpu_bench_n.sh <tag> <N>
tag=$1; N=$2; out=gpurun_out; mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N \
  bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_${tag}_n$N.json 2> $out/bench_${tag}_n$N.err
echo "bench n$N exit $?"; tail -c 600 $out/bench_${tag}_n$N.err
