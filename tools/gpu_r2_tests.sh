#!/bin/bash
# GPU pass for the round-2 functionality: the new parity tests, a regression subset, one bench line.  usage: bash tools/gpu_r2_tests.sh TAG
mkdir -p gpurun_out
TAG=${1:-r2t}
timeout 900 python -m pytest tests/test_gpu_pyref_golden.py tests/test_gpu_arity.py -m gpu -q > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_new.log
tail -5 gpurun_out/${TAG}_pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_pyref_golden.py --deselect tests/test_gpu_arity.py \
   -k "not config3_full_size and not config5_full_size" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 > gpurun_out/${TAG}_bench_A.json 2> gpurun_out/${TAG}_bench_A.err
cut -c1-300 gpurun_out/${TAG}_bench_A.json
