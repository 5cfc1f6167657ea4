#!/bin/bash
# Build every variant listed in tools/lab/variants.txt ("name|-Dflags") here (no GPU needed), or run them on the box.
cd "$(dirname "$0")"
mode=${1:-build}
while IFS='|' read -r name flags; do
  [ -z "$name" ] && continue
  if [ "$mode" = build ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -DVARIANT="\"$name\"" $flags -o pb_$name permbench.cu 2>&1 | grep -E "error" 
  else
    timeout 120 ./pb_$name 20
  fi
done < variants.txt
