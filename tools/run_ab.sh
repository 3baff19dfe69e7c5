cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-ab}
bash tools/ab.sh > gpurun_out/${TAG}_ab.txt 2>&1
cat gpurun_out/${TAG}_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch_bench.log 2>&1
for lib in build/variants/*.so; do
n=$(basename $lib .so)
LH_CUDA_LIB=$PWD/$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_$n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
