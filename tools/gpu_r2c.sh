#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2c}
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -12 > gpurun_out/${T}_pytest_fullsize.log
cat gpurun_out/${T}_pytest_fullsize.log
bash tools/ab.sh > gpurun_out/${T}_ab.txt 2>&1
cat gpurun_out/${T}_ab.txt
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"raster_|cull_meshlets|resolve_kernel|reduce_depth" --csv --log-file gpurun_out/${T}_counters.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_counters_bench.log 2>&1
du -sm gpurun_out
