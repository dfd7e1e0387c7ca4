#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_c9_tests_$i.txt 2>&1; tail -3 gpurun_out/r02_c9_tests_$i.txt | cut -c1-200; done
for i in 1 2 3 4 5 6; do timeout 100 python -m pytest tests/test_gpu_parity.py -q -x -k "compress_matches_its_score_path_large or eight_heads or any_group" 2>&1 | tail -1; done
AB_EXTRA=0 AB_WORKLOADS=ea_128k,decoding_knorm timeout 300 python tools/ab_variants.py > gpurun_out/r02_ab_run9.txt 2>&1; cat gpurun_out/r02_ab_run9.txt
timeout 60 python tools/cluster_profile.py > gpurun_out/r02_cluster_phases3.txt 2>&1; cat gpurun_out/r02_cluster_phases3.txt
echo run9 done
