#!/bin/bash
# compute-sanitizer passes over the single-GPU kernel tests (the reference has no race detection at all).
# Run on a GPU box:  scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]   (default: all four)
set -u
tools="${1:-memcheck racecheck synccheck initcheck}"
mkdir -p gpurun_out
for t in $tools; do
  echo "== compute-sanitizer --tool $t"
  compute-sanitizer --tool "$t" --error-exitcode 1 --target-processes all \
    python -m pytest tests/test_gpu_kernels.py -x -q -k "kth_abs or fused_sgd or oktopk_single_gpu_matches_oracle and 4096" \
    > "gpurun_out/sanitize_$t.log" 2>&1
  echo "exit $? (log gpurun_out/sanitize_$t.log)"; tail -3 "gpurun_out/sanitize_$t.log"
done
