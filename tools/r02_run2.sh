#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/r02_c2_tests.txt 2>&1; tail -25 gpurun_out/r02_c2_tests.txt | cut -c1-220
AB_EXTRA=0 AB_WORKLOADS=ea_128k,snapkv_32k,snapkv_128k_70b timeout 400 python tools/ab_variants.py > gpurun_out/r02_ab_variants.txt 2>&1; cat gpurun_out/r02_ab_variants.txt
PROFILE_FULL=0 PROFILE_WORKLOADS="ea_128k decoding_knorm snapkv_32k" bash tools/r02_profile.sh
for w in ea_128k decoding_knorm snapkv_32k; do python tools/ncu_launches.py gpurun_out/r02_launches_$w.csv | tail -12; done
echo run2 done
