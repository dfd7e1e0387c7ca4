#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_c5_tests.txt 2>&1; tail -4 gpurun_out/r02_c5_tests.txt | cut -c1-220
timeout 120 python tools/ea_experiments.py > gpurun_out/r02_ea_experiments2.txt 2>&1; cat gpurun_out/r02_ea_experiments2.txt
AB_EXTRA=0 AB_WORKLOADS=ea_128k,decoding_knorm,adakv_ea_128k timeout 300 python tools/ab_variants.py > gpurun_out/r02_ab_heads.txt 2>&1; cat gpurun_out/r02_ab_heads.txt
echo run5 done
