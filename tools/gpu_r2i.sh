#!/bin/bash
# colour diagnostics, then the launch list of one emulated rank of an 8-GPU world (one device), then a bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2i}
bash tools/gpu_diag.sh $T
for R in 3/8 0/8; do
  tag=$(echo $R | tr / of)
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-rank $R > gpurun_out/${T}_bench_rank${tag}.json 2> gpurun_out/${T}_bench_rank${tag}.err
  python - gpurun_out/${T}_bench_rank${tag}.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1]); print(sys.argv[1], 'fps', round(d['value'],1), d['pass_ms'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_rank3of8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --emulate-rank 3/8 > gpurun_out/${T}_launch_bench.log 2>&1
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
