#!/bin/bash
# kernels r2.5 (partial-round loops of the cooperative permutation unrolled by two): the tests that run it, its ncu capture, the default bench line
mkdir -p gpurun_out
TAG=${1:-r25}
NCU="ncu --set full --clock-control none --import-source on"
python -c "import stark_verifier_b200 as s; print(s.version())" > gpurun_out/${TAG}_version.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/${TAG}_version.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 $NCU -k regex:fri_challenges_coop -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_coop python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 2 > gpurun_out/${TAG}_ncu_coop.log 2>&1
{ cat gpurun_out/${TAG}_version.txt; python tools/ncu_summary.py gpurun_out/${TAG}_prof_coop.ncu-rep; } > gpurun_out/${TAG}_ncu_summary_fri_challenges_coop.txt 2>&1
rm -f gpurun_out/${TAG}_prof_coop.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench_config2_shapeA.json 2> gpurun_out/${TAG}_bench_config2.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_config2_shapeA.json')); e=d['e2e']
g=lambda k: round((e.get(k) or {}).get('value',0)) if isinstance(e.get(k),dict) else e.get(k)
print('value',round(d['value']),'resident_fs',round(d['resident_with_device_transcript']['value']),'e2e',round(e['value']),'frac',e.get('frac_of_h2d_only'),'large',g('large_batch'),'record',g('record_path'),'record_fs',g('record_path_device_transcript'),'full',g('full_verifier'),'traffic',d['roofline']['traffic'])"
head -12 gpurun_out/${TAG}_ncu_summary_fri_challenges_coop.txt
