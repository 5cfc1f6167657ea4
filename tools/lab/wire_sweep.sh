#!/bin/bash
# wire path: chunk size x lead chunks x compute streams (SVB_TRACE timeline of the last call of each setting)
mkdir -p gpurun_out
TAG=${1:-ws}
for cfg in "2 2 64 4" "2 1 64 4" "2 3 64 4" "2 2 48 4" "2 3 48 4" "2 2 96 4" "2 1 96 4" "2 2 64 3" "2 2 64 2" "2 4 32 4" "1 1 64 4"; do
  set -- $cfg
  SVB_TRACE=1 SVB_FS_PARTS=$1 SVB_FS_LEAD=$2 SVB_CHUNK_MB=$3 SVB_KSTREAMS=$4 timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 10 \
     > gpurun_out/${TAG}_wire.json 2> gpurun_out/${TAG}_trace.txt
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_wire.json')); print('parts $1 lead $2 chunk $3 streams $4: wire',round(d['value']),'full',round(d['full_verifier'].get('value',0)))"
  grep "svb trace" gpurun_out/${TAG}_trace.txt | tail -14 | head -1 | cut -c1-260
done
