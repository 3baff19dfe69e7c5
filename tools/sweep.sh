#!/bin/bash
# Runs bench.py once per LH_TUNE setting ("main_bands,shadow_bands,main_grid,shadow_grid") and prints the pass times.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-sweep}; shift
for t in "$@"; do
  LH_TUNE=$t timeout 300 python bench.py --workload ${WL:-c2} --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']
print('$t', round(d['value'],2), 'fps vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'])"
done | tee gpurun_out/${TAG}.txt
