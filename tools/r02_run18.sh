#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/ea_experiments.py 2>&1 | tee gpurun_out/r02_ea_pair_experiments.txt
echo run18 done
