#!/bin/bash
# compute-sanitizer over every kernel family of libsnb200.so on small shapes (scripts/sanitize_driver.py).
#   scripts/gpu_sanitize.sh [memcheck|racecheck|synccheck|initcheck]      (default memcheck; ONE tool per invocation)
# NOTE: the round-2 GPU pool refuses compute-sanitizer ("closed on this pool"): the recipe is for a box where the tool is allowed;
# tests/test_gpu_guard_bands.py is the out-of-bounds-write check that runs everywhere.
# Run on a GPU box:  gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh memcheck'
# The plain driver runs first (the tool is only worth its time on a program that exits 0 by itself); the log goes to
# gpurun_out/sanitize_<tool>.log and its summary lines to stdout.  CUDA graphs are captured with memcheck only: racecheck /
# synccheck instrument every launch and replaying a captured graph under them multiplies the run time.
set -u
TOOL="${1:-memcheck}"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG="gpurun_out/sanitize_${TOOL}.log"
timeout 600 python scripts/sanitize_driver.py > "gpurun_out/sanitize_plain.log" 2>&1 || { echo "plain driver failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
echo "plain driver: $(tail -1 gpurun_out/sanitize_plain.log)"
GRAPH=1; [ "$TOOL" != "memcheck" ] && GRAPH=0
SANITIZE_GRAPH=$GRAPH timeout 1200 compute-sanitizer --tool "$TOOL" --print-limit 50 --error-exitcode 86 \
  python scripts/sanitize_driver.py > "$LOG" 2>&1
RC=$?
echo "compute-sanitizer --tool $TOOL: exit code $RC"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver|Internal Sanitizer Error|=========     at|========= (Invalid|Race|Barrier|Uninitialized)" "$LOG" | sort | uniq -c | sort -rn | head -40
exit $RC
