#!/bin/bash
# source-level capture of the raster kernels of one frame (28 launches per frame; frame 4 is captured)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-prof}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_meshlets_kernel" --launch-skip 84 -c 28 \
    -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
gzip -9 gpurun_out/${T}_full.ncu-rep
ls -la gpurun_out
