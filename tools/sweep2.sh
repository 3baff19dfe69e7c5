#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload c2 --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']
print('$1', round(d['value'],2), 'fps vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'])"; }
for t in 4,6,16,3,1,16 3,6,16,3,1,16 5,6,16,3,1,16 4,6,12,3,1,16 4,6,24,3,1,16 4,5,16,3,1,16 4,7,16,3,1,16; do LH_TUNE=$t run "tune=$t"; done
for f in 0.05,0.2,0.5 0.02,0.08,0.3 0.03,0.15,0.5 0.04,0.12,0.3 0.06,0.25,0.6; do LH_TUNE_FRACTIONS=$f run "frac=$f"; done
