#!/bin/bash
# Multi-GPU bench pass: N ranks under torchrun (one per GPU), plus the reference arm.  Outputs -> gpurun_out/
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
for n in $(echo $N | tr ',' ' '); do
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  fi
  tail -1 gpurun_out/${TAG}_bench_n$n.json | cut -c1-400
done
