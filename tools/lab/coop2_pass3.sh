#!/bin/bash
# third GPU pass: high-priority prep stream (unpack + prepare) x ramp-down of the chunk schedule
mkdir -p gpurun_out
TAG=${1:-c5}
run() {  # label, env...
  local label="$1"; shift
  env SVB_TRACE=1 "$@" timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 10 > gpurun_out/${TAG}_wire.json 2> gpurun_out/${TAG}_trace.txt
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_wire.json')); print('$label: wire',round(d['value']),'full',round(d['full_verifier'].get('value',0)))"
  grep "svb trace" gpurun_out/${TAG}_trace.txt | head -20 | tail -2 | cut -c1-420
}
run "prep0" SVB_PREP_STREAM=0
run "prep1" SVB_PREP_STREAM=1
run "prep1 ramp64" SVB_RAMP=1
run "prep1 ramp128" SVB_RAMP=1 SVB_RAMP_MIN=128
run "prep1 chunk48" SVB_CHUNK_MB=48
run "prep1 chunk48 ramp96" SVB_CHUNK_MB=48 SVB_RAMP=1 SVB_RAMP_MIN=96
run "prep1 chunk32 ramp64" SVB_CHUNK_MB=32 SVB_RAMP=1
run "prep1 streams2" SVB_KSTREAMS=2
for prep in 0 1; do
  SVB_PREP_STREAM=$prep timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-wire > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); e=d.get('e2e') or {}
print('prep $prep: value',round(d['value']),'resident_fs',round(d.get('resident_with_device_transcript')['value']),'record',round(e.get('record_path')['value']),'record_fs',round(e.get('record_path_device_transcript')['value']), 'e2e', round(e.get('value',0)), 'full', e.get('full_verifier',{}).get('value'))"
done
