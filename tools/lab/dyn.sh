#!/bin/bash
# dynamic SASS opcode count of poseidon_permute_kernel<0> for one set of -D flags (no GPU needed). usage: dyn.sh name "-Dflags"
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -DVARIANT="\"$1\"" $2 -cubin -o /tmp/pb_$1.cubin permbench.cu 2>&1 | grep -E "error" 
cuobjdump -sass /tmp/pb_$1.cubin | python3 ../sass_dyn.py | head -${3:-14}
