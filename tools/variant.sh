#!/bin/bash
# builds build/variants/<name>.so: the library with raster.cu compiled with extra flags.  usage: variant.sh name "-DLH_BANDS=8 ..."
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants build/v_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++20 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -Ilighthugger_b200/csrc "$@" -c -o build/v_$name/raster.o lighthugger_b200/csrc/raster.cu
nvcc -shared -o build/variants/$name.so build/api.o build/comm.o build/cull.o build/depth.o build/resolve.o build/v_$name/raster.o -lcudart_static -ldl -lrt -lpthread 2>/dev/null
echo built $name
