#!/bin/bash
# multi-GPU: depth-band counts (LH_TUNE main,shadow,...) at N GPUs.  usage: gpu_mgpu_tune.sh N tag "tune1 tune2 ..."
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}; T=${2:-mgt}
for tune in $3; do
  name=$(echo $tune | tr , _)
  timeout 400 env LH_TUNE=$tune python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${name}_n$N.json 2> gpurun_out/${T}_bench_${name}_n$N.err
  python - "$tune" gpurun_out/${T}_bench_${name}_n$N.json <<'PY'
import json,sys
name,f=sys.argv[1:3]
try:
    d=json.loads(open(f).read().strip().split('\n')[-1]); p=d['pass_ms']
    print(name, 'fps', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'vis', p['visbuffer'], 'comp', p['composite'], 'shadow', p['shadowmaps'], 'shx', p.get('shadow_exchange'), 'resolve', p['render_geometry'], 'bar', p.get('frame_barrier'))
except Exception as e:
    print('ERR', name, e); print(open(f.replace('.json','.err')).read()[-800:])
PY
done
