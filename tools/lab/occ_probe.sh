#!/bin/bash
# Occupancy variants of the library (tools/lab/libsvb200_mb<N>.so built with -DSVB_MINBLOCKS=N): fused FRI bench + Merkle bench.
mkdir -p gpurun_out
TAG=${1:-occ}
OUT=gpurun_out/${TAG}_occ.txt
: > $OUT
cp stark-verifier_b200/libsvb200.so /tmp/default.so
for v in default mb5 mb6; do
  if [ $v != default ]; then cp tools/lab/libsvb200_$v.so stark-verifier_b200/libsvb200.so; fi
  echo "== $v" >> $OUT
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fused', d['value'], d['ms_per_step'])" >> $OUT
  timeout 300 python bench.py --workload merkle --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('merkle', d['value'])" >> $OUT
done
cp /tmp/default.so stark-verifier_b200/libsvb200.so
cat $OUT
