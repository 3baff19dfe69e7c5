cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}
TAG=scale
if [ "$2" = "test" ]; then
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/${TAG}_mgpu_pytest_n$N.log
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
if [ "$3" = "more" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --shadow-split cascades > gpurun_out/${TAG}_bench_cascades_n$N.json 2> gpurun_out/${TAG}_bench_cascades_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 10 --warmup 3 --exchange nccl > gpurun_out/${TAG}_bench_nccl_n$N.json 2> gpurun_out/${TAG}_bench_nccl_n$N.err
fi
for f in gpurun_out/${TAG}_bench*_n$N.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().split('\n')[-1]); p=d['pass_ms']; print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'vis',p['visbuffer'],'comp',p['composite'],'depth',p['depth_reduction'],'shadow',p['shadowmaps'],'resolve',p['render_geometry'],'cullm',p['cull_meshlets'], 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('ERR $f', e); print(open('${f%.json}.err').read()[-1500:])"; done
