#!/bin/bash
# e2e leg of bench.py in every host-path mode (same box, back to back): gpurun_out/e2e_modes.txt
cd "$(dirname "$0")/.."
out=gpurun_out/e2e_modes.txt; : > $out
for wl in ea_128k knorm_128k snapkv_32k streaming_128k; do
  for mode in serial staged zero_copy; do
    if [ $wl = ea_128k ] && [ $mode = zero_copy ]; then continue; fi
    line=$(timeout 200 python bench.py --workload $wl --steps 20 --no-cpu --e2e-mode $mode 2>&1 | tail -1)
    echo "$wl $mode $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print("e2e_ms=%.3f e2e_tok_s=%.3e h2d=%d d2h=%d dev_us=%.1f" % (e["ms_per_step"], e["value"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], d["ms_per_step"]*1e3))' 2>&1 | tail -1)" >> $out
  done
done
cat $out
