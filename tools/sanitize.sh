#!/bin/bash
# compute-sanitizer passes over small instances of every kernel of the path (run on a B200: `gpurun -- bash tools/sanitize.sh`).
# Output: gpurun_out/r02_sanitizer_*.log; the summary lines are copied into profiles/r02_sanitizer.md.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
SMOKE='import __graft_entry__ as g; g.smoke()'
SC='tests/test_gpu_scan_context.py -k "bit_exact or ringkey_exact or ties or empty or generate or not_multiple"'
run() { name=$1; lim=$2; shift 2
  ( time timeout "$lim" "$@" ) > gpurun_out/r02_sanitizer_$name.log 2>&1
  echo "$name: rc=$? $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke ok" gpurun_out/r02_sanitizer_$name.log | tr '\n' ' ')"
}
run memcheck_smoke 240 $CS --tool memcheck python -c "$SMOKE"
run memcheck_pyramid 300 $CS --tool memcheck python -m pytest -x -q -m gpu tests/test_gpu_pyramid.py
eval "run memcheck_sc 420 $CS --tool memcheck python -m pytest -x -q -m gpu $SC"
run memcheck_pe 300 $CS --tool memcheck python -m pytest -x -q -m gpu tests/test_gpu_pose_estimator.py -k "eval_matches or errors or regrow"
run racecheck_smoke 300 $CS --tool racecheck python -c "$SMOKE"
run synccheck_smoke 240 $CS --tool synccheck python -c "$SMOKE"
run initcheck_smoke 240 $CS --tool initcheck python -c "$SMOKE"
run memcheck_umma 120 $CS --tool memcheck python tools/sc_umma_check.py 8192 150
run racecheck_umma 120 $CS --tool racecheck python tools/sc_umma_check.py 8192 50
