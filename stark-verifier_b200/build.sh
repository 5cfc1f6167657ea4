#!/bin/bash
# Build libsvb200.so (CUDA kernels + C ABI + host side) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC,-O3,-pthread -Xptxas -v -shared \
  -o libsvb200.so csrc/capi.cu csrc/host_side.cpp csrc/wire_host.cpp csrc/plonk_host.cpp csrc/ntt_host.cpp -ldl 2>&1 | tee build_ptxas.log | grep -E "error|warning|registers|spill|Compiling entry|bytes stack" || true
# registers / stack / spill bytes of every kernel: build_ptxas.log (a copy per round is committed as profiles/ptxas_rN.txt)
ls -la libsvb200.so
