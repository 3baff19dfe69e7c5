#!/bin/bash
# GPU suite, A/B of the in-tree build against build/variants (base = previous commit), emulated rank 3/8
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2n}
rm -f gpurun_out/${T}_color.jsonl
LH_COLOR_LOG=gpurun_out/${T}_color.jsonl timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_multigpu.py 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_color.jsonl | cut -c1-400
bash tools/ab.sh 2>&1 | tee gpurun_out/${T}_ab.txt
for lib in "" build/variants/base.so; do
  LH_CUDA_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-rank 3/8 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']; print('rank3of8 ${lib:-intree}', round(d['value'],1), {k:p[k] for k in ('cull_meshlets','visbuffer','shadowmaps','render_geometry','frame')})" | tee -a gpurun_out/${T}_ab.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_rank3of8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --emulate-rank 3/8 > gpurun_out/${T}_launch_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_c2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launch_bench_c2.log 2>&1
