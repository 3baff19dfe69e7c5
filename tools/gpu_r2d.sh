#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2d}
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest_fullsize.log
cat gpurun_out/${T}_pytest_fullsize.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 2500 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
export BENCH_EXTRA="--tex-size 256"
bash tools/ab.sh > gpurun_out/${T}_ab.txt 2>&1
cat gpurun_out/${T}_ab.txt
for t in "4,8,16,3,1,16,1" "4,6,16,3,1,16,0" "4,6,16,3,0,16,1" "4,4,16,3,1,16,1" "6,6,16,3,1,16,1"; do
  LH_TUNE=$t timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --tex-size 256 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']
print('$t', round(d['value'],2), 'fps vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'])"
done > gpurun_out/${T}_tune.txt 2>&1
cat gpurun_out/${T}_tune.txt
du -sm gpurun_out
