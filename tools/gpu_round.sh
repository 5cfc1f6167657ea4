#!/bin/bash
# One GPU-box pass: parity tests, bench lines, ncu launch list + one full capture.  Outputs -> gpurun_out/
mkdir -p gpurun_out
TAG=${1:-run}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
# the files whose kernels were written without a GPU, one by one and without -x: a failure in one must not hide the others
for f in plonk transforms wire verify_full; do
  timeout 600 python -m pytest tests/test_gpu_${f}.py -m gpu -q > gpurun_out/${TAG}_pytest_gpu_${f}.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_gpu_${f}.log
  tail -2 gpurun_out/${TAG}_pytest_gpu_${f}.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_A.json 2> gpurun_out/${TAG}_bench_A.err
timeout 600 python bench.py --workload merkle --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_merkle.json 2> gpurun_out/${TAG}_bench_merkle.err
if [ -x tools/microbench/pipes2 ]; then timeout 120 tools/microbench/pipes2 > gpurun_out/${TAG}_pipes2.txt 2>&1; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fri_query -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_fri \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_bench_A.json; cat gpurun_out/${TAG}_bench_merkle.json; cat gpurun_out/${TAG}_smoke.log | tail -2
# hash family B (outer wrapped-proof configuration): bench line + one full capture of its query kernel
timeout 900 python bench.py --workload outer --proofs 1024 --steps 5 > gpurun_out/${TAG}_bench_outer.json 2> gpurun_out/${TAG}_bench_outer.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fri_query -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_fri_b \
  python bench.py --workload outer --proofs 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full_b.log 2>&1
cat gpurun_out/${TAG}_bench_outer.json | cut -c1-400
# wire format: the serialised-proof leg on its own, and one full capture of the gather kernel
timeout 600 python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 10 > gpurun_out/${TAG}_bench_wire.json 2> gpurun_out/${TAG}_bench_wire.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wire_unpack -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_wire \
  python bench.py --wire-leg --workload A --proofs 4096 --distinct 64 --steps 2 > gpurun_out/${TAG}_ncu_full_wire.log 2>&1
cat gpurun_out/${TAG}_bench_wire.json | cut -c1-400
