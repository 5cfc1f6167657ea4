#!/bin/bash
# BASELINE configs[3]: 2^20 shape-A proofs sharded over 8 B200s (131 072 per GPU, 20.7 GB resident each), accept bitmap gathered
# through sv_allgather_bitmap.  usage (gpurun --gpus 8): bash tools/gpu_config4.sh TAG [N]
mkdir -p gpurun_out
TAG=${1:-r2}; N=${2:-8}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_config4_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${TAG}_config4_smi.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --total-proofs 1048576 --steps 5 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${TAG}_bench_config4_${N}gpu.json 2> gpurun_out/${TAG}_bench_config4_${N}gpu.err
echo "rc=$?"
tail -1 gpurun_out/${TAG}_bench_config4_${N}gpu.json | cut -c1-600
tail -5 gpurun_out/${TAG}_bench_config4_${N}gpu.err
