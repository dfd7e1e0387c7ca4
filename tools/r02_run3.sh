#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_c3_tests.txt 2>&1; tail -4 gpurun_out/r02_c3_tests.txt | cut -c1-220
timeout 100 python tools/ea_profile.py > gpurun_out/r02_ea_roles.txt 2>&1; cat gpurun_out/r02_ea_roles.txt
AB_EXTRA=0 AB_WORKLOADS=ea_128k,knorm_128k,snapkv_32k,decoding_knorm timeout 300 python tools/ab_variants.py > gpurun_out/r02_ab_pdl.txt 2>&1; cat gpurun_out/r02_ab_pdl.txt
KVP_KNORM_CLUSTER=0 timeout 60 python bench.py --workload decoding_knorm --steps 200 --no-cpu --no-e2e --no-extras > gpurun_out/r02_bench_decoding_nocluster.json 2>/dev/null
timeout 60 python bench.py --workload decoding_knorm --steps 200 --no-cpu --no-e2e --no-extras > gpurun_out/r02_bench_decoding.json 2>/dev/null
python tools/summarize_bench.py gpurun_out/r02_bench_decoding_nocluster.json gpurun_out/r02_bench_decoding.json
echo run3 done
