#!/bin/bash
# compile one CUDA unit of radiosity_b200/csrc with the library's flags; print registers / spills and the SASS size per kernel
cd /root/repo/radiosity_b200/csrc || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC,-O2,-Wall -cudart static -Xptxas -v -c $1.cu -o $1.o 2>&1 | grep -E "error|warning|registers|spill|Compiling" | sed 's/ptxas info    : //' | paste - - - 2>/dev/null | sed 's/Compiling entry function//' | cut -c1-260
cuobjdump -sass $1.o | grep -E "Function|^\s+/\*[0-9a-f]{4}\*/" | awk '/Function/{name=$3; next} {c[name]++} END{for(n in c) print c[n], n}' | c++filt | cut -c1-160
