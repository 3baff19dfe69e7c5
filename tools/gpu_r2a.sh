#!/bin/bash
# Round-2 diagnostic call: sanity tests, the L2-atomic ceiling, a bench line, the rasteriser's
# atomic / issue counters, one `--set full` source capture of the raster kernels, and a few launch-shape A/Bs.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
timeout 60 tools/atomic_probe > gpurun_out/${T}_atomic_probe.txt 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
# counters of every rasteriser launch of one frame (3 warm frames + 1 timed + 2 e2e warm + 1 e2e; the last frame is used)
timeout 400 ncu --metrics gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"raster_|cull_meshlets|resolve_kernel|reduce_depth" --csv --log-file gpurun_out/${T}_counters.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_counters_bench.log 2>&1
# source-level capture: the cascades' raster kernel launches of one frame (28 raster_meshlets launches per frame)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_meshlets_kernel" --launch-skip 84 -c 28 \
    -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2>/dev/null
# keep only 4 launches' worth of source pages to stay under the size limit: cascade 3 band 3, cascade 2 band 3, main band 0/1
gzip -9 gpurun_out/${T}_full.ncu-rep
ls -la gpurun_out/
bash tools/ab.sh > gpurun_out/${T}_ab.txt 2>&1
for t in "4,4,16,3,1,16" "4,8,16,3,1,16" "4,6,16,2,1,16" "4,6,16,4,1,16" "4,6,16,6,1,16" "4,6,8,3,1,16" "4,6,24,3,1,16" "4,6,16,3,0,16"; do
  LH_TUNE=$t timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']
print('$t', round(d['value'],2), 'fps vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'])"
done > gpurun_out/${T}_tune.txt 2>&1
cat gpurun_out/${T}_pytest.log gpurun_out/${T}_atomic_probe.txt gpurun_out/${T}_ab.txt gpurun_out/${T}_tune.txt
tail -c 1500 gpurun_out/${T}_bench.json
du -sm gpurun_out
