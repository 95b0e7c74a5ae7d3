#!/bin/bash
# compute-sanitizer passes (the reference has no race detection at all).
#   scripts/sanitize.sh [tools] [single|multi]
# single: the single-GPU kernel tests; multi: a tiny 2-GPU Ok-Topk / gTopk run (cross-GPU mailboxes, TMA pulls from peer
# memory, last-CTA tickets) -- needs a 2-GPU box.  Logs land in gpurun_out/sanitize_<scope>_<tool>.log.
set -u
tools="${1:-memcheck racecheck synccheck initcheck}"
scope="${2:-single}"
mkdir -p gpurun_out
for t in $tools; do
  echo "== compute-sanitizer --tool $t ($scope)"
  if [ "$scope" = "multi" ]; then
    timeout 900 compute-sanitizer --tool "$t" --error-exitcode 1 --target-processes all \
      python -m pytest tests/test_multigpu.py -x -q -k "small_bucket_for_sanitizer" \
      > "gpurun_out/sanitize_multi_$t.log" 2>&1
  else
    timeout 900 compute-sanitizer --tool "$t" --error-exitcode 1 --target-processes all \
      python -m pytest tests/test_gpu_kernels.py -x -q -k "kth_abs or fused_sgd or (oktopk_single_gpu_matches_oracle and 4096) or land_grads_kernel or fused_update_is_skipped" \
      > "gpurun_out/sanitize_${scope}_$t.log" 2>&1
  fi
  echo "exit $? (log gpurun_out/sanitize_${scope}_$t.log)"; tail -3 "gpurun_out/sanitize_${scope}_$t.log"
done
