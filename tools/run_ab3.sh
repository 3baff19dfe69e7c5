cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-ab}
bash tools/ab.sh > gpurun_out/${TAG}_ab.txt 2>&1
cat gpurun_out/${TAG}_ab.txt
