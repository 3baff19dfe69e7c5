#!/bin/bash
# round-end evidence on one GPU: the whole -m gpu suite, smoke, both bench arms, launch list, counters, one --set full capture
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2z}
rm -f gpurun_out/${T}_color.jsonl
LH_COLOR_LOG=gpurun_out/${T}_color.jsonl timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_multigpu.py 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 600 gpurun_out/${T}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launch_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"raster_|cull_|resolve_kernel|reduce_depth" --csv --log-file gpurun_out/${T}_counters.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_counters_bench.log 2>&1
# --set full of the frame's raster and resolve kernels (frame 4: launches 3 * 119 .. of all kernels; -k filters first)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_meshlets_kernel|resolve_kernel|cull_meshlets" --launch-skip 102 -c 34 \
    -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i gpurun_out/${T}_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active > gpurun_out/${T}_full_raw.csv 2>/dev/null
gzip -9 -f gpurun_out/${T}_full.ncu-rep
ls -la gpurun_out | tail -12
