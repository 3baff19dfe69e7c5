cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-abl}
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py 2>&1 | tail -25
bash tools/ab.sh 2>&1 | cut -c1-150
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch_bench.log 2>&1
