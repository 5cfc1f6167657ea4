#!/bin/bash
# DRAM traffic of fri_query_kernel<0> (4 096 shape-A proofs, one launch) under the prefetch / L2-fetch-granularity knobs.
# Output: gpurun_out/<tag>_traffic.txt  (bytes read from DRAM, L2 read sectors requested by the SMs, duration)
mkdir -p gpurun_out
TAG=${1:-traffic}
OUT=gpurun_out/${TAG}_traffic.txt
: > $OUT
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_ltcfabric.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,gpu__time_duration.sum
i=0
run() {
  i=$((i+1))
  echo "== $*" >> $OUT
  env "$@" timeout 300 ncu --metrics $M --clock-control none -k regex:fri_query -s 2 -c 1 --csv --log-file gpurun_out/${TAG}_ncu_$i.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python - gpurun_out/${TAG}_ncu_$i.csv >> $OUT <<'PY'
import csv, sys
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 3 and 'fri_query' in ''.join(r):
        print('  ', r[-3], r[-2], r[-1])
PY
}
if [ "$2" = knobs ]; then
  run SVB_PREFETCH=1
  run SVB_PREFETCH=0
  run SVB_PREFETCH=2
  run SVB_PREFETCH=1 SVB_L2_FETCH=32
  run SVB_PREFETCH=0 SVB_L2_FETCH=32
  run SVB_PREFETCH=1 SVB_L2_FETCH=128
fi
# block order: class-major over the whole batch (0) against unit groups of 8 / 32 / 128 blocks per class
for g in ${GROUPS_LIST:-0 16 32 64}; do run SVB_GROUP_BLOCKS=$g; done
for g in ${GROUPS_LIST:-0 16 32 64}; do
  echo "== bench SVB_GROUP_BLOCKS=$g" >> $OUT
  env SVB_GROUP_BLOCKS=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" >> $OUT
done
cat $OUT
