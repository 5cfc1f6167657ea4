#!/bin/bash
# end-of-round pass at HEAD: compute-sanitizer on the kernels new in r2.4, the complete `pytest -m gpu` the driver runs, smoke(), the default
# bench line and the reference arm.  usage (on the GPU box): bash tools/gpu_final_r2.sh TAG
mkdir -p gpurun_out
TAG=${1:-fin}
S=gpurun_out/${TAG}_sanitizer.txt
{
echo "# compute-sanitizer on the kernels of revision r2.4 (cooperative permutation in its latency form, transcript, wire path with the priority stream)"
python -c "import stark_verifier_b200 as s; print(s.version())"
echo "## memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py -m gpu -q -x -k "cooperative or transcript or wire_verify or challenges" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -20
echo "memcheck rc=${PIPESTATUS[0]}"
echo "## racecheck (shared-memory exchange buffers of poseidon_g_coop2: one __syncwarp per round, double-buffered)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cooperative or transcript" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | head -20
echo "racecheck rc=${PIPESTATUS[0]}"
echo "## synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cooperative or transcript" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | head -20
echo "synccheck rc=${PIPESTATUS[0]}"
} > $S 2>&1
cat $S
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; tail -5 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_config2_shapeA.json 2> gpurun_out/${TAG}_bench_config2.err; tail -4 gpurun_out/${TAG}_bench_config2.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -4 gpurun_out/${TAG}_bench_reference.err
cut -c1-400 gpurun_out/${TAG}_bench_config2_shapeA.json; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
