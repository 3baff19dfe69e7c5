#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2g}
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_multigpu.py --deselect tests/test_gpu_fullsize.py 2>&1 | grep -E "Error|FAILED|passed|failed" > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | grep -E "Error|FAILED|passed|failed" > gpurun_out/${T}_pytest_fullsize.log
cat gpurun_out/${T}_pytest_fullsize.log
timeout 60 tools/texfilter_probe > gpurun_out/${T}_texfilter.txt 2>&1
timeout 600 python tools/resolve_diag.py tiny c1q c4s 2>&1 | grep -E "^==|max \|diff\|" > gpurun_out/${T}_resolve_diag.txt
cat gpurun_out/${T}_resolve_diag.txt
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 700 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
bash tools/gpu_sweep.sh ${T}sweep
