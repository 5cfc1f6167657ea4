#!/bin/bash
# timeline of sv_verify_proofs_wire / _full (SVB_TRACE=1).  usage: bash tools/lab/wire_trace.sh TAG
mkdir -p gpurun_out
TAG=${1:-wt}
for mb in 64; do
  echo "== chunk ${mb} MiB" >> gpurun_out/${TAG}_trace.txt
  SVB_TRACE=1 SVB_CHUNK_MB=$mb timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 6 \
     > gpurun_out/${TAG}_wire_${mb}.json 2>> gpurun_out/${TAG}_trace.txt
done
grep "svb trace" gpurun_out/${TAG}_trace.txt | sed -n '4,6p;12,14p'
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-wire > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<'PY' ${TAG}
import json,sys
d=json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json")); e=d["e2e"]
print("e2e two-in-flight",round(e["value"]),"single",round(e["single_call"]["value"]),"h2d_only",round(e["h2d_only_proofs_per_s"]),"full",e.get("full_verifier"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_wire_launches.csv \
  python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 1 > /dev/null 2>&1
python - <<'PY' ${TAG}
import csv,sys
rows=[r for r in csv.reader(open(f"gpurun_out/{sys.argv[1]}_wire_launches.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[:60]:
    print(r[4][:60], r[-1])
PY
