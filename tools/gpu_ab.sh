cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-ab}
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_multigpu.py 2>&1 | tail -25
bash tools/ab.sh > gpurun_out/${TAG}_ab.txt 2>&1
cat gpurun_out/${TAG}_ab.txt
