#!/bin/bash
# On the GPU box: run every lab variant, the pipe microbenchmark, then (optionally) tests + bench.
mkdir -p gpurun_out
TAG=${1:-lab}
( cd tools/lab && for f in pb_*; do timeout 120 ./$f ${LAB_DEPTH:-20}; done ) > gpurun_out/${TAG}_variants.txt 2>&1
timeout 120 tools/microbench/pipes2 > gpurun_out/${TAG}_pipes2.txt 2>&1
if [ "$2" = full ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_A.json 2> gpurun_out/${TAG}_bench_A.err
  tail -3 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_bench_A.json
fi
cat gpurun_out/${TAG}_variants.txt gpurun_out/${TAG}_pipes2.txt
