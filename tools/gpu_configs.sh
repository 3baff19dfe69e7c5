# the other BASELINE configurations, one short run each (parity cases, not the bench line)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for wl in c1 c4 c5; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/cfg_$wl.json 2> gpurun_out/cfg_$wl.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/cfg_$wl.json').read().strip().split('\n')[-1]); p=d['pass_ms']; f=d['frame']
    print('$wl', round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms vis', p['visbuffer'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'], 'tris', d['config']['input_triangles'], 'res', d['config']['resolution'], 'visible', f['visible_instances'], 'rasterised', f['rasterised_meshlets'])
except Exception as e: print('ERR $wl', e, open('gpurun_out/cfg_$wl.err').read()[-800:])"
done
