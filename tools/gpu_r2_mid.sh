#!/bin/bash
# mid-round GPU pass: every -m gpu test except the two full-size configs, transforms bench, default bench line
mkdir -p gpurun_out
TAG=${1:-mid}
timeout 1500 python -m pytest tests -m gpu -q -k "not config3_full_size and not config5_full_size" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --transforms-leg --steps 10 > gpurun_out/${TAG}_transforms.json 2> gpurun_out/${TAG}_transforms.err
cat gpurun_out/${TAG}_transforms.json; tail -3 gpurun_out/${TAG}_transforms.err
timeout 900 python bench.py --steps 20 > gpurun_out/${TAG}_bench_A.json 2> gpurun_out/${TAG}_bench_A.err
python - <<'PY' ${TAG}
import json,sys
d=json.load(open(f"gpurun_out/{sys.argv[1]}_bench_A.json"))
e=d.get("e2e") or {}
w=e.get("wire") or {}
print("value",d["value"],"e2e",e.get("value"),"h2d_only",e.get("h2d_only_proofs_per_s"),"wire",w.get("value"),"full",(w.get("full_verifier") or {}).get("value"), w.get("error"))
PY
tail -3 gpurun_out/${TAG}_bench_A.err
