#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
AB_WORKLOADS=ea_128k,knorm_128k,snapkv_32k,keydiff_128k,rerotate_knorm_128k,adakv_ea_128k,snapkv_128k_70b,decoding_knorm timeout 400 python tools/ab_env.py default 2>&1 | tee gpurun_out/r02_ab_run20.txt
echo run20 done
