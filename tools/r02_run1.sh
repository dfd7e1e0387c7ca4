#!/bin/bash
# first comprehensive GPU call of round 2 (one box): tests, host overhead, A/B of the kernel variants, bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpu.txt
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02_c1_tests.txt 2>&1; tail -5 gpurun_out/r02_c1_tests.txt
timeout 120 python tools/host_overhead.py > gpurun_out/r02_c1_host_overhead.txt 2>&1; cat gpurun_out/r02_c1_host_overhead.txt
AB_EXTRA=0 AB_WORKLOADS=ea_128k,snapkv_32k,snapkv_128k_70b timeout 300 python tools/ab_variants.py > gpurun_out/r02_ab_variants.txt 2>&1; cat gpurun_out/r02_ab_variants.txt
timeout 200 python tools/ab_knorm_fused.py > gpurun_out/r02_ab_knorm_fused.txt 2>&1; cat gpurun_out/r02_ab_knorm_fused.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 3000 gpurun_out/r02_bench_default.json; tail -3 gpurun_out/r02_bench_default.err
timeout 60 python bench.py --workload decoding_knorm --steps 200 --no-cpu --no-e2e --no-extras > gpurun_out/r02_bench_decoding.json 2> gpurun_out/r02_bench_decoding.err
timeout 60 python bench.py --workload knorm_128k --steps 50 --no-cpu --no-e2e --no-extras > gpurun_out/r02_bench_knorm.json 2> gpurun_out/r02_bench_knorm.err
echo run1 done
