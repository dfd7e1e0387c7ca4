#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in memcheck synccheck racecheck; do
  echo "== $t"; timeout 110 compute-sanitizer --tool $t python tools/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all kernels ran|Error|error" | head -5
done 2>&1 | tee gpurun_out/r02_compute_sanitizer.txt
echo run25 done
