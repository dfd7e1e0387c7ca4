#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T='python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider'
echo "== default lib, EA tests"; timeout 200 $T -k "expected_attention" 2>&1 | grep -E "passed|failed|FAILED" | head -3
echo "== default lib, the large test alone"; timeout 200 $T -k "compress_matches_its_score_path_large" 2>&1 | grep -E "passed|failed|FAILED" | head -3
echo "== default lib, launch blocking"; CUDA_LAUNCH_BLOCKING=1 timeout 200 $T -k "expected_attention" 2>&1 | grep -E "passed|failed|FAILED|Error" | head -5
for v in vst0 vst2 nohint; do echo "== $v"; KVPRESS_B200_LIB=$PWD/tools/bin/libv_$v.so timeout 200 $T -k "expected_attention" 2>&1 | grep -E "passed|failed|FAILED" | head -3; done
echo "== default lib full suite"; timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED" | head -3
echo run10 done
