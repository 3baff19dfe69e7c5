#!/bin/bash
# colour diagnostics on the GPU box: texture-unit value probe + fit, per-term attribution, then the colour tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-diag}
if [ -x build/texvalue_probe ]; then
  timeout 120 build/texvalue_probe gpurun_out && python tools/texvalue_fit.py gpurun_out > gpurun_out/${T}_texvalue_fit.txt 2>&1
  cat gpurun_out/${T}_texvalue_fit.txt
fi
timeout 900 python tools/resolve_diag.py tiny c4s c2 > gpurun_out/${T}_resolve_diag.txt 2>&1
grep -v "^   px" gpurun_out/${T}_resolve_diag.txt
timeout 900 python -m pytest tests/test_resolve_gpu.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_color.log
