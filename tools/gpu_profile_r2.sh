#!/bin/bash
# Round-2 profile pass of the SHIPPED kernels: bench line, launch list, one `ncu --set full` capture per kernel.
# usage (on the GPU box): bash tools/gpu_profile_r2.sh TAG   -> gpurun_out/TAG_*
mkdir -p gpurun_out
TAG=${1:-r2}
NCU="ncu --set full --clock-control none --import-source on"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python bench.py > gpurun_out/${TAG}_bench_A.json 2> gpurun_out/${TAG}_bench_A.err
cut -c1-600 gpurun_out/${TAG}_bench_A.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 $NCU -k regex:fri_query -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_fri_query \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_fri.log 2>&1
summarise() {   # .ncu-rep -> text summary + raw metric csv; the report itself (15-50 MB with sources) stays on the box
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_$1.ncu-rep > gpurun_out/${TAG}_ncu_summary_$1.txt 2>&1
  ncu -i gpurun_out/${TAG}_prof_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_$1.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_prof_$1.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_details_$1.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_$1.ncu-rep
}
summarise fri_query
for k in wire_unpack wire_pi_hash fri_challenges plonk_check ntt_pass lde_scale_pad; do
  C=1; S=1; if [ $k = ntt_pass ]; then C=22; S=0; fi
  timeout 400 $NCU -k regex:$k -s $S -c $C -f -o gpurun_out/${TAG}_prof_$k \
    python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 2 > gpurun_out/${TAG}_ncu_$k.log 2>&1
  echo "$k rc=$?"
  summarise $k
done
ls -la gpurun_out | tail -20
