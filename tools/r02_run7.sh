#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_c7_tests.txt 2>&1; tail -4 gpurun_out/r02_c7_tests.txt | cut -c1-220
timeout 120 python tools/ea_experiments.py > gpurun_out/r02_ea_experiments4.txt 2>&1; cat gpurun_out/r02_ea_experiments4.txt
AB_EXTRA=0 AB_WORKLOADS=ea_128k,decoding_knorm timeout 300 python tools/ab_variants.py > gpurun_out/r02_ab_vstages.txt 2>&1; cat gpurun_out/r02_ab_vstages.txt
timeout 60 python tools/cluster_profile.py > gpurun_out/r02_cluster_phases2.txt 2>&1; cat gpurun_out/r02_cluster_phases2.txt
echo run7 done
