#!/bin/bash
# multi-GPU call: correctness (tests/test_multigpu.py) then bench variants.  usage: gpu_mgpu.sh N tag [test]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}; T=${2:-mg}
nvidia-smi topo -m > gpurun_out/${T}_topo_n$N.txt 2>&1
if [ "$3" = "test" ]; then
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/${T}_mgpu_pytest_n$N.log
fi
run() {  # name, extra bench args, env
  local name=$1; shift
  timeout 400 env $ENVX python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${T}_bench_${name}_n$N.json 2> gpurun_out/${T}_bench_${name}_n$N.err
  python - "$name" gpurun_out/${T}_bench_${name}_n$N.json <<'PY'
import json,sys
name,f=sys.argv[1:3]
try:
    d=json.loads(open(f).read().strip().split('\n')[-1]); p=d['pass_ms']
    print(name, 'fps', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'rank0 frame', p['frame'], 'vis', p['visbuffer'], 'comp', p['composite'], 'depth', p['depth_reduction'], 'shadow', p['shadowmaps'], 'shx', p.get('shadow_exchange'), 'resolve', p['render_geometry'], 'bar', p.get('frame_barrier'), 'e2e', round(d['e2e']['value'],1))
    print('   spread', {k:(v['min'],v['max']) for k,v in (d.get('pass_ms_over_ranks') or {}).items() if k in ('frame','visbuffer','shadowmaps','composite','shadow_exchange','frame_barrier','render_geometry','cull_meshlets')})
    print('   balance', (d.get('stripe_balance') or {}).get('imbalance_main_shadow_per_iteration'))
except Exception as e:
    print('ERR', name, e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
}
ENVX="" run balanced
ENVX="" run equal_rows --equal-rows
if [ "$3" != "quick" ]; then
ENVX="" run equal --balance 0
fi
if [ "$4" = "more" ]; then
ENVX="LH_TUNE=3,4,16,3,1,16,1,128" run bands34
ENVX="LH_TUNE=4,6,16,3,1,16,1,32" run area32
ENVX="" run cascades --shadow-split cascades
ENVX="" run nccl --exchange nccl
fi
