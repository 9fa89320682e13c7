#!/bin/bash
# compute-sanitizer over the hot path (one B200): memcheck + racecheck + synccheck on smoke() and on the raster stress /
# ragged-input / pool-overflow tests.  Logs go to gpurun_out/<tag>_sanitize_*.log; the summaries are kept in profiles/.
TAG=${1:-r2}
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
SMOKE='python -c "import __graft_entry__ as g; g.smoke()"'
TESTS='python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "raster_stress or empty_and_ragged or pool_overflow or degenerate or config1_anchor"'
run() {  # name, tool flags, command
  local name=$1; shift; local flags=$1; shift
  timeout 900 $CS $flags bash -c "$*" > gpurun_out/${TAG}_sanitize_${name}.log 2>&1
  echo "$name rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_${name}.log | tail -1)"
}
run memcheck_smoke  "--tool memcheck --leak-check no" "$SMOKE"
run racecheck_smoke "--tool racecheck --racecheck-report all" "$SMOKE"
run synccheck_smoke "--tool synccheck" "$SMOKE"
run memcheck_tests  "--tool memcheck --leak-check no" "$TESTS"
run racecheck_tests "--tool racecheck --racecheck-report all" "$TESTS"
for f in gpurun_out/${TAG}_sanitize_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke ok|Error|hazard" $f | sort | uniq -c | head -12; done > gpurun_out/${TAG}_sanitize_summary.txt
cat gpurun_out/${TAG}_sanitize_summary.txt
