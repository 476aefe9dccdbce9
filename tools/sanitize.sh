#!/usr/bin/env bash
# compute-sanitizer over tools/sanitizer_workload.py (SURVEY.md section 5 "race detection"): memcheck, racecheck,
# initcheck, synccheck.  Summaries go to gpurun_out/sanitizer_<tool>.txt; copy them to profiles/ to commit.
#   bash tools/sanitize.sh [tool ...]        (default: all four)
set -uo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
OUT="$ROOT/gpurun_out"
mkdir -p "$OUT"
TOOLS=("$@")
[[ ${#TOOLS[@]} -eq 0 ]] && TOOLS=(memcheck racecheck initcheck synccheck)
export MTB_GC_FREEZE=0 CUDA_MODULE_LOADING=LAZY
for t in "${TOOLS[@]}"; do
  log="$OUT/sanitizer_$t.txt"
  echo "== compute-sanitizer --tool $t" | tee "$log"
  timeout "${SANITIZE_TIMEOUT:-900}" compute-sanitizer --tool "$t" --print-limit 20 --launch-timeout 0 \
      --kernel-name kns=3mtb \
      python "$ROOT/tools/sanitizer_workload.py" ${SANITIZE_MODES:-} >> "$log" 2>&1
  echo "exit code $?" >> "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok|exit code|\[(fp32|tf32|bf16)\] ok" "$log" | tail -8
done
