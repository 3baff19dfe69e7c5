#!/bin/bash
# A/B of the in-tree build against build/variants at N=1 and as emulated rank 3/8; quick parity check first
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2o}
timeout 900 python -m pytest tests/test_resolve_gpu.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
bash tools/ab.sh 2>&1 | tee gpurun_out/${T}_ab.txt
for lib in "" build/variants/*.so; do
  LH_CUDA_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-rank 3/8 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pass_ms']; print('rank3of8 ${lib:-intree}', round(d['value'],1), {k:p[k] for k in ('cull_meshlets','visbuffer','shadowmaps','render_geometry','frame')})" | tee -a gpurun_out/${T}_ab.txt
done
