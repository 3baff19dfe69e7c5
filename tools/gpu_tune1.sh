#!/bin/bash
# LH_TUNE variants on one GPU.  usage: gpu_tune1.sh tag "tune1 tune2 ..."
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-tune}
for tune in default $2; do
  if [ $tune = default ]; then unset LH_TUNE; else export LH_TUNE=$tune; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']
print('$tune', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],1), 'vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'], 'frame', p['frame'])" | tee -a gpurun_out/${T}_tune.txt
done
