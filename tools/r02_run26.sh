#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
{ for t in memcheck synccheck racecheck; do echo "== $t"; timeout 110 compute-sanitizer --tool $t python tools/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all kernels ran|Race reported|Read access|Write access" | cut -c1-260 | head -8; done; } 2>&1 | tee gpurun_out/r02_compute_sanitizer.txt
AB_WORKLOADS=ea_128k,adakv_ea_128k timeout 120 python tools/ab_env.py default KVP_EA_PAIR=0 2>&1 | head -2 | tee gpurun_out/r02_ab_ea_pair_final.txt
echo run26 done
