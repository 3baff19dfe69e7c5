cd $GRAFT_REPO_ROOT
bash tools/run_quick.sh ${1:-both}
N=2
TAG=${1:-both}
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/${TAG}_mgpu_pytest.log
cat gpurun_out/${TAG}_mgpu_pytest.log
for part in stripes range; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --partition $part > gpurun_out/${TAG}_bench_${part}_n$N.json 2> gpurun_out/${TAG}_bench_${part}_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --shadow-split cascades > gpurun_out/${TAG}_bench_cascades_n$N.json 2> gpurun_out/${TAG}_bench_cascades_n$N.err
