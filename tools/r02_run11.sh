#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T='python -m pytest tests/test_gpu_parity.py tests/test_gpu_wrappers.py -q -x -p no:cacheprovider'
echo "== EA tests"; timeout 90 $T -k "expected_attention or tova or large32k or stats_press" 2>&1 | grep -E "passed|failed|FAILED" | head -3
echo "== knorm/decoding tests"; timeout 90 $T -k "knorm" 2>&1 | grep -E "passed|failed|FAILED" | head -3
echo "== full suite x2"; for i in 1 2; do timeout 240 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED" | head -3; done
AB_EXTRA=0 AB_WORKLOADS=ea_128k,decoding_knorm timeout 200 python tools/ab_variants.py > gpurun_out/r02_ab_run11.txt 2>&1; cat gpurun_out/r02_ab_run11.txt
timeout 60 python tools/cluster_profile.py > gpurun_out/r02_cluster_phases4.txt 2>&1; cat gpurun_out/r02_cluster_phases4.txt
echo run11 done
