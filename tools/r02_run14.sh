#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PROFILE_FULL=0 PROFILE_WORKLOADS="ea_128k knorm_128k snapkv_32k decoding_knorm" bash tools/r02_profile.sh
timeout 60 python tools/cluster_profile.py > gpurun_out/r02_cluster_phases_v4.txt 2>&1; cat gpurun_out/r02_cluster_phases_v4.txt
timeout 100 python tools/ea_profile.py > gpurun_out/r02_ea_roles_v4.txt 2>&1; cat gpurun_out/r02_ea_roles_v4.txt
timeout 200 python -m pytest tests/test_gpu_wrappers.py -q -x -p no:cacheprovider -k "outside_the_tensor_core or shape_limits" 2>&1 | tail -3
echo run14 done
