#!/bin/bash
# Runs on the GPU box (via gpurun): (1) ncu launch lists (time + DRAM bytes per launch) of one eager bench command per
# workload -> gpurun_out/r02_launches_<w>.csv, parsed locally by tools/ncu_launches.py into profiles/traffic.json;
# (2) one `--set full` capture of the dominant kernels. Numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in ${PROFILE_WORKLOADS:-ea_128k knorm_128k snapkv_32k snapkv_128k_70b decoding_knorm streaming_128k}; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --cache-control none --csv --log-file gpurun_out/r02_launches_${w}.csv \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph --workload $w > /dev/null 2>&1
  echo "$w: $(grep -c -E 'kernel' gpurun_out/r02_launches_${w}.csv) metric rows"
done
if [ "${PROFILE_FULL:-1}" = "1" ]; then
  NCU="timeout 300 ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
  B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph"
  $NCU -k regex:ea_logits_pair -o gpurun_out/r02_prof_ea_logits_pair $B --workload ea_128k > /dev/null 2>&1
  KVP_EA_PAIR=0 $NCU -k regex:ea_logits_kernel -o gpurun_out/r02_prof_ea_logits_onecta $B --workload ea_128k > /dev/null 2>&1
  $NCU -k regex:select_compact -o gpurun_out/r02_prof_ea_select $B --workload ea_128k > /dev/null 2>&1
  $NCU -k regex:ea_finalize -o gpurun_out/r02_prof_ea_finalize $B --workload ea_128k > /dev/null 2>&1
  $NCU -k regex:knorm_score -o gpurun_out/r02_prof_knorm_score $B --workload knorm_128k > /dev/null 2>&1
  $NCU -k regex:select_compact -o gpurun_out/r02_prof_knorm_select $B --workload knorm_128k > /dev/null 2>&1
  $NCU -k regex:snap_stats -o gpurun_out/r02_prof_snap_stats $B --workload snapkv_32k > /dev/null 2>&1
  $NCU -k regex:snap_colsum -o gpurun_out/r02_prof_snap_colsum $B --workload snapkv_32k > /dev/null 2>&1
  $NCU -k regex:knorm_cluster -o gpurun_out/r02_prof_knorm_cluster $B --workload decoding_knorm > /dev/null 2>&1
  ls -la gpurun_out/r02_prof_*.ncu-rep
fi
