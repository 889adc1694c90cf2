#!/bin/bash
# compute-sanitizer over every kernel family on the small shapes of the GPU tests (run on a B200 box: gpurun -- bash scripts/sanitize.sh).
# Logs land in gpurun_out/r2_sanitizer_*.log; the summary lines are copied to profiles/r2_sanitizer.md by hand.
mkdir -p gpurun_out
san() { # name tool timeout pytest-args...
  local name=$1 tool=$2 lim=$3; shift 3
  local t0=$(date +%s)
  timeout $lim compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_${tool}_$name.log \
      python -m pytest "$@" -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_sanitizer_${tool}_$name.pytest 2>&1
  local rc=$?
  echo "$name $tool rc=$rc $(( $(date +%s) - t0 ))s | $(tail -1 gpurun_out/r2_sanitizer_${tool}_$name.pytest) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_$name.log | tail -1)"
}
san scoring memcheck ${SAN_LIMIT:-240} tests/test_gpu_scoring.py
san match_nms memcheck ${SAN_LIMIT:-150} tests/test_gpu_matching.py tests/test_gpu_nms.py
san match_nms racecheck ${SAN_LIMIT:-150} tests/test_gpu_matching.py tests/test_gpu_nms.py
san fit memcheck ${SAN_LIMIT:-240} tests/test_gpu_fit.py -k "vec_score or percentile or labels_match or seed_ or tcgen05_step or centring or pair_cluster or silhouette"
