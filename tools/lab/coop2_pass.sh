#!/bin/bash
# GPU pass for the latency form of the cooperative permutation: latencies, coopbench, the transcript tests, wire-path sweep
mkdir -p gpurun_out
TAG=${1:-c2}
timeout 60 tools/microbench/lat > gpurun_out/${TAG}_lat.txt 2>&1
timeout 200 tools/lab/coopbench > gpurun_out/${TAG}_coopbench.txt 2>&1
cat gpurun_out/${TAG}_lat.txt; grep -v "transcript-like:  *[28]" gpurun_out/${TAG}_coopbench.txt
timeout 900 python -m pytest tests/test_gpu_wire.py tests/test_gpu_verify_full.py tests/test_gpu_parity.py -m gpu -x -q -k "transcript or wire or full or challenges" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for cfg in "2 2 2 0" "2 1 2 0" "2 3 2 0" "2 4 2 0" "3 1 3 0" "3 1 2 0" "3 2 2 0" "2 2 2 1" "2 1 2 1" "1 1 2 1" "3 1 3 1"; do
  set -- $cfg
  SVB_TRACE=1 SVB_FS_PARTS=$1 SVB_FS_LEAD=$2 SVB_FS_MID=$3 SVB_FS_REST_COOP=$4 timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 10 \
     > gpurun_out/${TAG}_wire.json 2> gpurun_out/${TAG}_trace.txt
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_wire.json')); print('parts $1 lead $2 mid $3 restcoop $4: wire',round(d['value']),'full',round(d['full_verifier'].get('value',0)))"
  grep "svb trace" gpurun_out/${TAG}_trace.txt | tail -14 | head -1 | cut -c1-260
done
