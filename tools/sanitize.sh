#!/bin/bash
# compute-sanitizer over one small case per kernel family (tools/sanitize_cases.py).
#   tools/sanitize.sh [out_dir]        # default gpurun_out/sanitize
# memcheck over every family; racecheck (shared-memory hazards) over the families whose kernels
# hand shared memory between warps: the tcgen05 convolutions (TMA producer / MMA issuer /
# epilogue warps, CTA-pair mbarrier protocol), the weight gradient and the fusion kernels.
# Summaries (the tools' own RESULT lines) go to <out_dir>/summary.txt.
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
SAN=${SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
: > "$OUT/summary.txt"
run() {   # tool family
  local log="$OUT/$1_$2.log"
  timeout 900 "$SAN" --tool "$1" --error-exitcode 9 --print-limit 20 \
      python tools/sanitize_cases.py "$2" > "$log" 2>&1
  local rc=$?
  echo "$1 $2: rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1)" | tee -a "$OUT/summary.txt"
}
for fam in ${FAMILIES_MEM:-conv2cta convt conv1 fcn tails wgrad fusion confusion}; do run memcheck "$fam"; done
for fam in ${FAMILIES_RACE:-conv2cta convt conv1 tails wgrad fusion confusion}; do run racecheck "$fam"; done
