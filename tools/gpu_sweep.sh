#!/bin/bash
# BASELINE configs 1, 4, 5 and the culling-ratio x resolution sweep of config 5 on one GPU (one JSON line each).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2sweep}
for wl in c1 c4 c5; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_${wl}.json 2> gpurun_out/${T}_${wl}.err
done
for wl in c2 c5; do
  for vis in 0.1 0.4 0.9; do
    timeout 400 python bench.py --workload $wl --visible $vis --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_${wl}_vis${vis}.json 2> gpurun_out/${T}_${wl}_vis${vis}.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/*sweep*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1]); p=d['pass_ms']; c=d['config']
        print(f.split('/')[-1], 'fps', round(d['value'],1), 'Mtri/s', round(d['culled_mtri_per_s']), 'visible', d['frame']['visible_instances'], 'frac', c.get('visible_instance_fraction'), 'vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'], 'cull', p['cull_meshlets'])
    except Exception as e:
        print('ERR', f, e)
PY
