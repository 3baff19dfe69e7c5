#!/bin/bash
# A/B: runs bench.py once per library variant under build/variants/ (plus the in-tree build) and prints the pass times.
wl=${WL:-c2}
shopt -s nullglob
for lib in "" build/variants/*.so; do
  LH_CUDA_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline $BENCH_EXTRA 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']; f=d['frame']
print('${lib:-intree}', round(d['value'],2), 'fps vis', p['visbuffer'], 'rasterK', p['raster_meshlets_kernel'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'], 'occl', f['meshlets_skipped_occluded'], 'empty', f['meshlets_skipped_no_sample'], 'n', [f['rasterised_meshlets']]+f['shadow_rasterised_meshlets'])"
done
