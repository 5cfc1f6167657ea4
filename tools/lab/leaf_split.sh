#!/bin/bash
# leaf-digest split (fri_leaf_kernel before the transcript) on / off: parity tests, wire-path timeline, bench line.
# usage: bash tools/lab/leaf_split.sh TAG
mkdir -p gpurun_out
TAG=${1:-ls}
timeout 900 python -m pytest tests -m gpu -q -x -k "wire or full or transcript or fs or pyref or arity or python_prover" > gpurun_out/${TAG}_pytest_on.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_on.log
tail -3 gpurun_out/${TAG}_pytest_on.log
SVB_LEAF_SPLIT=0 timeout 900 python -m pytest tests -m gpu -q -x -k "wire or full or transcript" > gpurun_out/${TAG}_pytest_off.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_off.log
tail -3 gpurun_out/${TAG}_pytest_off.log
for cfg in "1 2 2 64" "0 2 2 64" "1 2 1 64" "1 2 2 32" "1 2 1 32" "1 3 1 32" "1 2 3 32"; do
  set -- $cfg
  echo "== split $1 parts $2 lead $3 chunk $4" >> gpurun_out/${TAG}_trace.txt
  SVB_TRACE=1 SVB_LEAF_SPLIT=$1 SVB_FS_PARTS=$2 SVB_FS_LEAD=$3 SVB_CHUNK_MB=$4 timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 8 \
     > gpurun_out/${TAG}_wire.json 2>> gpurun_out/${TAG}_trace.txt
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_wire.json')); print('split $1 parts $2 lead $3 chunk $4: wire',round(d['value']),'full',round(d['full_verifier'].get('value',0)))"
  grep "svb trace" gpurun_out/${TAG}_trace.txt | tail -9 | head -1
done
for sp in 1 0; do
  SVB_LEAF_SPLIT=$sp timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/${TAG}_bench_split$sp.json 2> gpurun_out/${TAG}_bench_split$sp.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_split$sp.json")); e=d["e2e"]
print("split $sp: value",round(d["value"]),"e2e",round(e["value"]),"frac",round(e.get("frac_of_h2d_only",0),3),"large",round(e.get("large_batch",{}).get("value",0)),
      "rec_fs",round(e.get("record_path_device_transcript",{}).get("value",0)),"full",round(e.get("full_verifier",{}).get("value",0)),
      "resident_fs",round(d.get("resident_with_device_transcript",{}).get("value",0)))
PY
done
