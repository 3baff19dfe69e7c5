set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r1d
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_meshlets_kernel|cull_meshlets_kernel|resolve_kernel" --launch-skip 105 -c 35 -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2>/dev/null
gzip -9 gpurun_out/${T}_full.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
cat gpurun_out/${T}_pytest.log gpurun_out/${T}_smoke.log
du -sm gpurun_out
