#!/bin/bash
# compute-sanitizer racecheck + memcheck over the WHOLE GPU suite (full-size configs included)
cd "${GRAFT_REPO_ROOT:-.}"
for tool in racecheck memcheck; do
  echo "==== compute-sanitizer --tool $tool: pytest tests -m gpu"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_${tool}_suite.txt 2>&1
  echo "exit $?"
  grep -c "Race reported\|Invalid\|Error:" gpurun_out/r2_${tool}_suite.txt
  tail -5 gpurun_out/r2_${tool}_suite.txt
done
