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
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:ea_logits -s 3 -c 1 -f -o gpurun_out/r02_prof_ea_logits \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph --workload ea_128k > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:select_compact -s 3 -c 1 -f -o gpurun_out/r02_prof_ea_select \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph --workload ea_128k > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:knorm_cluster -s 3 -c 1 -f -o gpurun_out/r02_prof_knorm_cluster \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph --workload decoding_knorm > /dev/null 2>&1
  ls -la gpurun_out/r02_prof_*.ncu-rep
fi
