#!/bin/bash
# final records of HEAD: GPU suite (with the measured ulp distances printed), smoke, bench line of every workload, launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r02_gpu_tests_final.txt 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.txt | cut -c1-200
grep -E "^\[measured\]" gpurun_out/r02_gpu_tests_final.txt | sort | uniq -c | sort -rn | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
bash tools/final_bench.sh 2>&1 | tail -9
timeout 120 python bench.py --workload adakv_ea_128k --no-cpu > gpurun_out/final/adakv_ea_128k.json 2> gpurun_out/final/adakv_ea_128k.err
PROFILE_FULL=0 PROFILE_WORKLOADS="ea_128k knorm_128k snapkv_32k" bash tools/r02_profile.sh 2>&1 | tail -4
echo run23 done
