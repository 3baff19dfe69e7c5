cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
for spr in ${SPRS:-8 4 16 32}; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 8 --warmup 3 --stripes-per-rank $spr > gpurun_out/stripes_${spr}_n$N.json 2> gpurun_out/stripes_${spr}_n$N.err
python -c "
import json
try:
    d=json.loads(open('gpurun_out/stripes_${spr}_n$N.json').read().strip().split('\n')[-1]); p=d['pass_ms']
    print('spr=$spr', round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms vis', p['visbuffer'], 'comp', p['composite'], 'shadow', p['shadowmaps'], 'resolve', p['render_geometry'])
except Exception as e: print('ERR', e)"
done
