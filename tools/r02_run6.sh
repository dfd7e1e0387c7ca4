#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_c6_tests.txt 2>&1; tail -4 gpurun_out/r02_c6_tests.txt | cut -c1-220
timeout 120 python tools/ea_experiments.py > gpurun_out/r02_ea_experiments3.txt 2>&1; cat gpurun_out/r02_ea_experiments3.txt
AB_EXTRA=0 AB_WORKLOADS=ea_128k,adakv_ea_128k timeout 300 python tools/ab_variants.py > gpurun_out/r02_ab_inkernel_v.txt 2>&1; cat gpurun_out/r02_ab_inkernel_v.txt
timeout 60 python tools/cluster_profile.py > gpurun_out/r02_cluster_phases.txt 2>&1; cat gpurun_out/r02_cluster_phases.txt
timeout 100 python tools/snap_profile.py > gpurun_out/r02_snap_roles.txt 2>&1; cat gpurun_out/r02_snap_roles.txt
KVP_KNORM_CLUSTER=4 timeout 60 python bench.py --workload decoding_knorm --steps 200 --no-cpu --no-e2e --no-extras > gpurun_out/r02_bench_decoding_c4.json 2>/dev/null
python tools/summarize_bench.py gpurun_out/r02_bench_decoding_c4.json
echo run6 done
