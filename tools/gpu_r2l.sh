#!/bin/bash
# whole GPU suite with the colour statistics logged, smoke, one bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2l}
rm -f gpurun_out/${T}_color.jsonl
LH_COLOR_LOG=gpurun_out/${T}_color.jsonl timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_multigpu.py 2>&1 | tail -15 | tee gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_color.jsonl
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - gpurun_out/${T}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1]); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['pass_ms'])
PY
tail -3 gpurun_out/${T}_bench.err
