#!/bin/bash
# AddressSanitizer + UBSan over the HOST twins of the device functions (wire_record_word, plonk_check_one with every gate,
# ntt_pass_block) and the rest of the host side: the same sources built with -fsanitize, loaded through SVB200_LIB, driven
# by the CPU test-suite.  No GPU needed; an out-of-bounds index in the shared __host__ __device__ code shows up here.
set -e
cd "$(dirname "$0")/.."
OUT=${1:-/tmp/asan}
mkdir -p "$OUT"
(cd stark-verifier_b200 && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC,-pthread,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer -shared -o "$OUT/libsvb200.so" \
  csrc/capi.cu csrc/host_side.cpp csrc/wire_host.cpp csrc/plonk_host.cpp csrc/ntt_host.cpp -ldl 2>&1 | grep -iE "error" || true)
SVB200_LIB="$OUT/libsvb200.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
  python -m pytest tests/test_wire_format.py tests/test_plonk_check.py tests/test_full_proof.py tests/test_ntt.py tests/test_host_logic.py \
  tests/test_oracle_kat.py -q -p no:cacheprovider -s 2>&1 | grep -iE "runtime error|AddressSanitizer|passed|failed" | sort | uniq -c
