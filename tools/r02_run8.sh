#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "compress_matches_its_score_path_large" > gpurun_out/r02_sanitizer_ea.txt 2>&1; grep -E "=========|passed|failed" gpurun_out/r02_sanitizer_ea.txt | head -40
for mb in 0 80 110; do KVP_KNORM_L2_KEEP_MB=$mb AB_EXTRA=0 AB_WORKLOADS=knorm_128k timeout 100 python tools/ab_variants.py 2>&1 | grep default | head -1 | sed "s/^/keep_mb=$mb /"; done > gpurun_out/r02_ab_knorm_l2.txt; cat gpurun_out/r02_ab_knorm_l2.txt
echo run8 done
