#!/bin/bash
# the bench contract on hardware: default line, reference arm, and (with --gpus 2) the two-rank line.  usage: bash tools/gpu_bench_check.sh TAG [N]
mkdir -p gpurun_out
TAG=${1:-bc}; N=${2:-1}
timeout 900 python bench.py --steps 10 > gpurun_out/${TAG}_bench_A.json 2> gpurun_out/${TAG}_bench_A.err; echo "bench rc=$?"
tail -5 gpurun_out/${TAG}_bench_A.err
python - <<'PY' ${TAG}
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}_bench_A.json"))
    e=d.get("e2e") or {}
    print("value",round(d["value"]),"resident_fs",d["resident_with_device_transcript"],"\ne2e",round(e.get("value",0)),"single",e.get("single_call"),"h2d_only",round(e.get("h2d_only_proofs_per_s",0)),
          "frac",e.get("frac_of_h2d_only"),"\nrecord",e.get("record_path"),"\nrecord_fs",e.get("record_path_device_transcript"),"\nfull",e.get("full_verifier"),"\ncpu",d.get("cpu_baseline"))
except Exception as ex: print("parse failed", ex)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
if [ "$N" -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "n$N rc=$?"
  tail -3 gpurun_out/${TAG}_bench_n$N.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench_n$N.json') if l.startswith('{')][-1]); print('N=$N value',round(d['value']),'e2e',round(d['e2e']['value']),d['config']['bitmap_gather'])"
fi
