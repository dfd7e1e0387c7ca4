#!/bin/bash
# final validation + records of HEAD: GPU suite, smoke, bench line of every workload, ncu launch lists + full captures,
# measured parity distances of the 32k reference fixture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests_final.txt 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.txt | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_wrappers.py -q -s -p no:cacheprovider -k "large32k_scores_and_kept" 2>&1 | grep -E "^\{|passed|failed" > gpurun_out/r02_parity_large32k.txt; tail -3 gpurun_out/r02_parity_large32k.txt | cut -c1-200
bash tools/final_bench.sh 2>&1 | tail -12
PROFILE_WORKLOADS="ea_128k knorm_128k snapkv_32k snapkv_128k_70b decoding_knorm streaming_128k" bash tools/r02_profile.sh 2>&1 | tail -20
echo run19 done
