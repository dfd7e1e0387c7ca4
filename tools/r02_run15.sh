#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T='python -m pytest -q -x -p no:cacheprovider'
echo "== EA parity tests (pair kernel default)"; timeout 200 $T tests/test_gpu_parity.py -k "expected_attention" 2>&1 | tail -15 | cut -c1-300
echo "== wrappers"; timeout 200 $T tests/test_gpu_wrappers.py 2>&1 | tail -5 | cut -c1-300
echo "== A/B"; AB_WORKLOADS=ea_128k,adakv_ea_128k timeout 400 python tools/ab_env.py default KVP_EA_PAIR=0 2>&1 | tee gpurun_out/r02_ab_ea_pair.txt
echo "== full suite"; timeout 400 $T tests -m gpu 2>&1 | tail -4 | cut -c1-300
echo run15 done
