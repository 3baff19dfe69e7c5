#!/bin/bash
# Round-2 call B: the pipelined raster front end (bulk copies + records).  Tests first (a wrong kernel stops here).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2b}
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_pytest_fullsize.log
cat gpurun_out/${T}_pytest_fullsize.log
timeout 60 tools/texfilter_probe > gpurun_out/${T}_texfilter.txt 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
bash tools/ab.sh > gpurun_out/${T}_ab.txt 2>&1
for t in "4,6,16,3,1,16,0" "4,6,16,4,1,16,1" "4,6,16,4,1,16,0" "4,8,16,4,1,16,1" "4,6,16,2,1,16,1"; do
  LH_TUNE=$t timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']
print('$t', round(d['value'],2), 'fps vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'])"
done > gpurun_out/${T}_tune.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"raster_|cull_meshlets|resolve_kernel|reduce_depth" --csv --log-file gpurun_out/${T}_counters.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_counters_bench.log 2>&1
cat gpurun_out/${T}_ab.txt gpurun_out/${T}_tune.txt
tail -c 600 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
du -sm gpurun_out
