#!/bin/bash
# wire path timeline (SVB_TRACE=1) under transcript-part schedules.  usage: bash tools/lab/wire_trace.sh TAG
mkdir -p gpurun_out
TAG=${1:-wt}
for cfg in "2 2 64" "2 1 64" "2 3 64" "2 2 32" "2 4 32" "3 2 64"; do
  set -- $cfg
  echo "== parts $1 lead $2 chunk $3" >> gpurun_out/${TAG}_trace.txt
  SVB_TRACE=1 SVB_FS_PARTS=$1 SVB_FS_LEAD=$2 SVB_CHUNK_MB=$3 timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 6 \
     > gpurun_out/${TAG}_wire.json 2>> gpurun_out/${TAG}_trace.txt
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_wire.json')); print('parts $1 lead $2 chunk $3: wire',round(d['value']),'full',round(d['full_verifier'].get('value',0)))"
  grep "svb trace" gpurun_out/${TAG}_trace.txt | tail -9 | head -1
done
