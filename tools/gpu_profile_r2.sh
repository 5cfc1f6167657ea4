#!/bin/bash
# Round-2 profile pass of the SHIPPED kernels (run at HEAD; every summary is stamped with sv_version()): bench lines of configs 2, 3,
# 5 and the outer configuration, launch list of the default bench, one `ncu --set full` capture per kernel summarised ON THE BOX
# (the .ncu-rep files carry the sources and are 15-50 MB each: they stay there).
# usage (on the GPU box): bash tools/gpu_profile_r2.sh TAG   -> gpurun_out/TAG_*
mkdir -p gpurun_out
TAG=${1:-r2}
NCU="ncu --set full --clock-control none --import-source on"
python -c "import stark_verifier_b200 as s; print(s.version())" > gpurun_out/${TAG}_version.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/${TAG}_version.txt 2>&1
summarise() {   # .ncu-rep -> text summary + raw metric csv
  { cat gpurun_out/${TAG}_version.txt; python tools/ncu_summary.py gpurun_out/${TAG}_prof_$1.ncu-rep; } > gpurun_out/${TAG}_ncu_summary_$1.txt 2>&1
  ncu -i gpurun_out/${TAG}_prof_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_$1.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_$1.ncu-rep
}
# ---- bench lines ----
timeout 900 python bench.py > gpurun_out/${TAG}_bench_config2_shapeA.json 2> gpurun_out/${TAG}_bench_config2.err; echo "config2 rc=$?"
timeout 900 python bench.py --workload merkle --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config5_merkle.json 2> gpurun_out/${TAG}_bench_config5.err; echo "config5 rc=$?"
timeout 1500 python bench.py --workload B --proofs 65536 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_config3_shapeB_65536.json 2> gpurun_out/${TAG}_bench_config3.err; echo "config3 rc=$?"
timeout 900 python bench.py --workload outer --proofs 1024 --steps 5 --no-wire > gpurun_out/${TAG}_bench_outer_hash_b.json 2> gpurun_out/${TAG}_bench_outer.err; echo "outer rc=$?"
for f in config2_shapeA config5_merkle config3_shapeB_65536 outer_hash_b; do cut -c1-260 gpurun_out/${TAG}_bench_$f.json; done
# ---- launch list of the default bench (shares of the step) ----
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
# ---- full captures ----
timeout 600 $NCU -k regex:fri_query -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_fri_query \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_fri.log 2>&1
summarise fri_query
timeout 600 $NCU -k regex:merkle_verify -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_merkle_verify \
  python bench.py --workload merkle --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_merkle.log 2>&1
summarise merkle_verify
for k in wire_unpack wire_header_unpack wire_pi_hash fri_challenges_coop fri_challenges_kernel plonk_check; do
  timeout 600 $NCU -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$k \
    python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 2 > gpurun_out/${TAG}_ncu_$k.log 2>&1
  echo "$k rc=$?"
  summarise $k
done
# the commit-phase kernels: launches 0-5 = the LDE (one pass), then the 2^22 forward transform (two passes per call), the inverse, the commitment
timeout 600 $NCU -k regex:"ntt_pass|merkle_leaf_hash_cols|merkle_level" -s 20 -c 24 -f -o gpurun_out/${TAG}_prof_transforms \
   python bench.py --transforms-leg --steps 3 > gpurun_out/${TAG}_ncu_transforms.log 2>&1
summarise transforms
ls -la gpurun_out | grep ${TAG} | head -60
