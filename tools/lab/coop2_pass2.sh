#!/bin/bash
# second GPU pass after the latency form of the transcript: chunk schedule of the wire path (ramp-down, chunk size), lead size
mkdir -p gpurun_out
TAG=${1:-c3}
run() {  # label, env...
  local label="$1"; shift
  env SVB_TRACE=1 "$@" timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 10 > gpurun_out/${TAG}_wire.json 2> gpurun_out/${TAG}_trace.txt
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_wire.json')); print('$label: wire',round(d['value']),'full',round(d['full_verifier'].get('value',0)))"
  grep "svb trace" gpurun_out/${TAG}_trace.txt | tail -14 | head -1 | cut -c1-250
}
run "default" SVB_X=0
run "ramp min64" SVB_RAMP=1
run "ramp min128" SVB_RAMP=1 SVB_RAMP_MIN=128
run "ramp min224" SVB_RAMP=1 SVB_RAMP_MIN=224
run "chunk48" SVB_CHUNK_MB=48
run "chunk48 ramp128" SVB_CHUNK_MB=48 SVB_RAMP=1 SVB_RAMP_MIN=128
run "chunk32" SVB_CHUNK_MB=32
run "chunk32 ramp128" SVB_CHUNK_MB=32 SVB_RAMP=1 SVB_RAMP_MIN=128
run "lead1248" SVB_FS_LEAD_PROOFS=1248
run "lead2080" SVB_FS_LEAD_PROOFS=2080
run "lead1248 ramp128" SVB_FS_LEAD_PROOFS=1248 SVB_RAMP=1 SVB_RAMP_MIN=128
# record path with device transcript and resident batch with device transcript: lead size
for lead in 1024 1664 2304; do
  SVB_FS_LEAD_PROOFS=$lead timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-wire > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); e=d.get('e2e') or {}
print('lead $lead: value',round(d['value']),'resident_fs',d.get('resident_with_device_transcript'),'record',e.get('record_path'),'record_fs',e.get('record_path_device_transcript'))"
done
