#!/bin/bash
# validation of HEAD: GPU suite twice (flakiness check), then the default bench line with all extras
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_c12_tests_$i.txt 2>&1; tail -3 gpurun_out/r02_c12_tests_$i.txt | cut -c1-200; done
for i in 1 2 3; do timeout 100 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "expected_attention" 2>&1 | tail -1; done
timeout 500 python bench.py > gpurun_out/r02_bench_ea_128k_run12.json 2> gpurun_out/r02_bench_ea_128k_run12.err; tail -c 600 gpurun_out/r02_bench_ea_128k_run12.err
python tools/summarize_bench.py gpurun_out/r02_bench_ea_128k_run12.json 2>&1 | tail -30
AB_EXTRA=0 AB_WORKLOADS=ea_128k,knorm_128k,snapkv_32k,snapkv_128k_70b,decoding_knorm timeout 200 python tools/ab_variants.py 2>&1 | tee gpurun_out/r02_ab_run12.txt
echo run12 done
