#!/bin/bash
# Lab variants of the permutation (tools/lab/variants.txt) + the fused FRI bench, Merkle bench and parity of the built library.
mkdir -p gpurun_out
TAG=${1:-perm}
OUT=gpurun_out/${TAG}_perm.txt
( cd tools/lab && for f in pb_*; do timeout 120 ./$f 20; done ) > $OUT 2>&1
for rep in 1 2; do
  echo "== fused bench" >> $OUT
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" >> $OUT
done
echo "== parity" >> $OUT
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "poseidon or fri or merkle or transcript" 2>&1 | tail -2 >> $OUT
timeout 300 python bench.py --workload merkle --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('merkle', d['value'])" >> $OUT
cat $OUT
