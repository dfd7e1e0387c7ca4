#!/bin/bash
# Runs on the GPU box (via gpurun): ncu launch lists (time + DRAM bytes per launch) of one bench command per
# workload, written to gpurun_out/; parsed locally by tools/ncu_launches.py into profiles/traffic.json.
# bench does >= 3 warm-up + 6 timed calls; skip the first 4 calls' kernels, list the next 4 calls.
mkdir -p gpurun_out
for spec in ea_128k:3 snapkv_32k:5 snapkv_128k_70b:5 knorm_128k:2 streaming_128k:1 decoding_knorm:1; do
  w=${spec%%:*}; n=${spec##*:}
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --cache-control none -s $((4 * n + 4)) -c $((4 * n)) --csv --log-file gpurun_out/launches_${w}_final.csv \
      python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --workload $w > /dev/null 2>&1
  echo "$w: $(grep -c -E 'kvp|kernel' gpurun_out/launches_${w}_final.csv) metric rows"
done
