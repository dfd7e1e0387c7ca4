#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/ea_experiments.py > gpurun_out/r02_ea_experiments.txt 2>&1; cat gpurun_out/r02_ea_experiments.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:knorm_cluster -s 4 -c 1 -f -o gpurun_out/r02_prof_knorm_cluster \
   python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph --workload decoding_knorm > /dev/null 2>&1
ls -la gpurun_out/r02_prof_knorm_cluster.ncu-rep
echo run4 done
