#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python tools/ea_profile.py 2>&1 | tee gpurun_out/r02_ea_roles_pair.txt
KVP_EA_PAIR=0 timeout 100 python tools/ea_profile.py 2>&1 | tee gpurun_out/r02_ea_roles_onecta.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ea_logits_pair -s 3 -c 1 -f -o gpurun_out/r02_prof_ea_logits_pair \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph --workload ea_128k > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
echo run16 done
