set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1c_pytest.log
timeout 400 python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_meshlets_kernel|resolve_kernel" --launch-skip 93 -c 31 -o gpurun_out/r1c_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1c_ncu.log 2>&1
gzip -9 gpurun_out/r1c_full.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_launch_bench.log 2>&1
ls -la gpurun_out
du -sm gpurun_out
