#!/bin/bash
# builds build/variants/<name>.so: the library with resolve.cu compiled with extra flags.  usage: variant_resolve.sh name "-DLH_RESOLVE_MIN_CTAS=4 ..."
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants build/v_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++20 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -Ilighthugger_b200/csrc "$@" -c -o build/v_$name/resolve.o lighthugger_b200/csrc/resolve.cu
nvcc -shared -o build/variants/$name.so build/api.o build/comm.o build/cull.o build/depth.o build/raster.o build/v_$name/resolve.o -lcudart_static -ldl -lrt -lpthread 2>/dev/null
echo built $name
