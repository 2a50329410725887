#!/bin/bash
# K2 on N GPUs, both shardings.  usage: gpu_k2_multi.sh <tag> <N>
tag=$1; N=$2; out=gpurun_out; mkdir -p $out
for shard in draw site; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
    bench.py --workload calibration --steps 3 --cpu-draws 1 --shard $shard \
    > $out/k2_${tag}_n${N}_$shard.json 2> $out/k2_${tag}_n${N}_$shard.err
  echo "k2 n$N $shard exit $?"; tail -c 900 $out/k2_${tag}_n${N}_$shard.json; tail -2 $out/k2_${tag}_n${N}_$shard.err
done
