#!/bin/bash
# blocks of fri_query_kernel per SM (SVB_QUERY_BPS, capped through dynamic shared memory): resident throughput, host pipelines, wire timeline
mkdir -p gpurun_out
TAG=${1:-bps}
for cfg in "0 0" "4 0" "3 0" "2 0" "3 1" "2 1" "4 1"; do
  set -- $cfg
  export SVB_QUERY_BPS=$1
  if [ "$2" = 1 ]; then export SVB_RAMP=1; else unset SVB_RAMP; fi
  SVB_TRACE=1 timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); e=d.get('e2e') or {}
g=lambda k: round((e.get(k) or {}).get('value',0))
print('bps $1 ramp $2: value',round(d['value']),'resident_fs',round(d['resident_with_device_transcript']['value']),'e2e',round(e.get('value',0)),'large',g('large_batch'),'record',g('record_path'),'record_fs',g('record_path_device_transcript'),'full',g('full_verifier'))"
  grep "svb trace\] wire" gpurun_out/${TAG}_bench.err | head -12 | tail -1 | cut -c1-250
done
