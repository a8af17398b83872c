#!/usr/bin/env bash
# compute-sanitizer over every kernel family (scripts/sanitize_target.py).  ONE tool per invocation and per gpurun call
# (B200_PROFILING.md: several tools in one call have wedged the GPU):
#     gpurun --timeout 1500 -- 'bash scripts/sanitize.sh memcheck'      # or racecheck | synccheck | initcheck
# The plain run must exit 0 first; the log goes to gpurun_out/sanitize_<tool>.log, its summary to stdout.
set -u
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 3 \
    python scripts/sanitize_target.py > "gpurun_out/sanitize_${tool}.log" 2>&1
rc=$?
echo "compute-sanitizer --tool $tool: exit code $rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|loss|ok" "gpurun_out/sanitize_${tool}.log" | tail -12
grep -E "=========.*(Invalid|Race|hazard|Barrier|Uninitialized)" "gpurun_out/sanitize_${tool}.log" | sort | uniq -c | sort -rn | head -20
exit $rc
