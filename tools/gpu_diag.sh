#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-diag}
timeout 900 python tools/resolve_diag.py c1q c4s c2 > gpurun_out/${T}_resolve_diag.txt 2>&1
cat gpurun_out/${T}_resolve_diag.txt
