# compute-sanitizer over one small frame (smoke: tiny scene, all passes) -- SURVEY.md 4, test plan item 8
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
