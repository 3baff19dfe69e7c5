set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/mgpu_topo.txt 2>&1
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/mgpu_pytest.log
for ex in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --exchange $ex > gpurun_out/mgpu_bench_${ex}_n$N.json 2> gpurun_out/mgpu_bench_${ex}_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --shadow-split cascades > gpurun_out/mgpu_bench_cascades_n$N.json 2> gpurun_out/mgpu_bench_cascades_n$N.err
tail -3 gpurun_out/mgpu_*.err
cat gpurun_out/mgpu_pytest.log
