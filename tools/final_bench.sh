#!/bin/bash
# bench line of every workload (gpurun_out/final/<workload>.json); the default workload also times the CPU leg
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/final
timeout 200 python bench.py > gpurun_out/final/ea_128k.json 2> gpurun_out/final/ea_128k.err
for wl in knorm_128k snapkv_32k streaming_128k keydiff_128k rerotate_knorm_128k snapkv_128k_70b decoding_knorm; do
  timeout 120 python bench.py --workload $wl --no-cpu > gpurun_out/final/$wl.json 2> gpurun_out/final/$wl.err
done
for f in gpurun_out/final/*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%-22s %8.1f us  frac %.3f  e2e %.3f ms (%s)  launches/step %d" % (
        d["config"]["workload"], d["ms_per_step"] * 1e3, d["roofline"]["frac"], e.get("ms_per_step", float("nan")),
        e.get("mode"), d["gpu_launches"] // d["steps"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
