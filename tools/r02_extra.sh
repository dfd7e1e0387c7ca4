#!/bin/bash
# extra measurements piggy-backed on a GPU call (time-boxed by the caller)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( AB_EXTRA=0 AB_WORKLOADS=snapkv_32k,snapkv_128k_70b,decoding_knorm timeout 200 python tools/ab_variants.py > gpurun_out/r02_ab_snap.txt 2>&1 )
( timeout 200 python tools/ab_knorm_fused.py > gpurun_out/r02_ab_knorm_fused.txt 2>&1 )
( timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err )
( timeout 60 python bench.py --workload decoding_knorm --steps 200 --no-cpu --no-e2e --no-extras > gpurun_out/r02_bench_decoding.json 2> gpurun_out/r02_bench_decoding.err )
( timeout 60 python bench.py --workload knorm_128k --steps 50 --no-cpu --no-e2e --no-extras > gpurun_out/r02_bench_knorm.json 2> gpurun_out/r02_bench_knorm.err )
echo extra done
