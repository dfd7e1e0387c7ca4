#!/bin/bash
# multi-rank check (N = number of visible GPUs): torchrun bench line (weak-scaled headline + configs[4] layer split + e2e
# with NUMA binding) and the reference arm under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_ea_128k_n$N.json 2> gpurun_out/r02_bench_ea_128k_n$N.err
tail -c 400 gpurun_out/r02_bench_ea_128k_n$N.err
python - gpurun_out/r02_bench_ea_128k_n$N.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
print("n_gpus", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["numa_binding"])
for x in d["extras"]: print(" extra", x["workload"], {k: x[k] for k in ("ms_per_prefill_pass", "us_per_layer", "layers_per_rank", "tokens_per_s") if k in x})
print(d["extras_errors"])
PY
if [ "${REF_ARM:-1}" = "1" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>&1 | grep -E '^\{' | cut -c1-600
fi
echo run13 done
