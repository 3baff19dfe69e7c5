cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r1d}
N=${2:-2}
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py 2>&1 | tail -25
bash tools/sweep.sh ${TAG}_sweep 4,6,16,3,1 4,6,16,3,0
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/${TAG}_mgpu_pytest.log
for part in stripes range; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --partition $part > gpurun_out/${TAG}_bench_${part}_n$N.json 2> gpurun_out/${TAG}_bench_${part}_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --shadow-split cascades > gpurun_out/${TAG}_bench_cascades_n$N.json 2> gpurun_out/${TAG}_bench_cascades_n$N.err
for f in gpurun_out/${TAG}_bench_*_n$N.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().split('\n')[-1]); p=d['pass_ms']; print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'vis',p['visbuffer'],'comp',p['composite'],'shadow',p['shadowmaps'],'resolve',p['render_geometry'],'cullm',p['cull_meshlets'], 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('ERR $f', e)"; done
