#!/bin/bash
# NTT / LDE on the GPU: parity with the host twin, throughput, one ncu capture per pass kind.  usage: bash tools/gpu_ntt.sh TAG
mkdir -p gpurun_out
TAG=${1:-ntt}
timeout 900 python -m pytest tests/test_gpu_transforms.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --transforms-leg --steps 10 > gpurun_out/${TAG}_transforms.json 2> gpurun_out/${TAG}_transforms.err
cat gpurun_out/${TAG}_transforms.json; tail -3 gpurun_out/${TAG}_transforms.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 6 -c 12 -f -o gpurun_out/${TAG}_prof \
   python bench.py --transforms-leg --steps 3 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_summary.txt 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_prof.ncu-rep
grep -E "kernel:|dram__bytes|gpu__time_duration|grid_size|issue_active|bank|stalls" gpurun_out/${TAG}_ncu_summary.txt | head -60
