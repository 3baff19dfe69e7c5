set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_pytest.log
tail -c 600 gpurun_out/${TAG}_bench.err
