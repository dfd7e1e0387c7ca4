#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_EXTRA=0 AB_WORKLOADS=ea_128k timeout 600 python tools/ab_variants.py 2>&1 | tee gpurun_out/r02_ab_ea_pair_tmem.txt
echo "== ld32 parity"; KVPRESS_B200_LIB=$PWD/tools/bin/libv_ld32.so timeout 200 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_parity.py -k "expected_attention" 2>&1 | tail -2
echo run17 done
