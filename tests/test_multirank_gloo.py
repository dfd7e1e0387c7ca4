"""world_size-2 checks of bench.py's multi-rank bookkeeping on CPU (gloo): shards are disjoint, the step
time is the max over ranks, the reported value is the whole-job aggregate. The data path itself has no
collective to test: every (batch, layer, kv-head) row is compressed independently."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    w = dict(bench.WORKLOADS["knorm_128k"], S=512)
    K, V, _ = bench.make_inputs(w, "cpu", bench.rank_seed(rank))
    local_ms = 10.0 + 5.0 * rank  # rank 1 is the slow one
    job_ms = bench.max_over_ranks(local_ms, dist, "cpu")
    value = bench.whole_job_tokens_per_s(w["B"] * w["S"], world, job_ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, float(K.float().sum()))
    # configs[4]: layers pipeline-split over the ranks; every rank learns every rank's range
    owned = [None] * world
    dist.all_gather_object(owned, list(bench.layer_range(80, world, rank)))
    torch.save({"job_ms": job_ms, "value": value, "sums": gathered, "owned": owned}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bookkeeping(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    results = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for res in results:
        assert res["job_ms"] == 15.0                      # max over ranks, identical on every rank
        assert abs(res["value"] - 2 * 512 / 15e-3) < 1e-6   # aggregate of both shards
        assert res["sums"][0] != res["sums"][1]            # ranks synthesise different shards
    assert results[0]["sums"] == results[1]["sums"]
    owned = results[0]["owned"]
    assert owned == results[1]["owned"] and owned[0] == list(range(0, 40)) and owned[1] == list(range(40, 80))


def test_layer_ranges_partition_the_model():
    import bench

    for n_layers in (80, 36, 32, 7):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in bench.layer_range(n_layers, world, r)]
            assert got == list(range(n_layers))                       # contiguous, disjoint, complete, in rank order
    assert len(bench.layer_range(80, 8, 3)) == 10 and list(bench.layer_range(7, 8, 7)) == []


def test_reference_arm_prints_what_it_ran():
    """`bench.py --impl reference`: exactly --steps timed calls after --warmup, ms_per_step = the mean call time (the
    driver checks steps x ms_per_step against its own clock), config identical to the B200 arm's, kind = reference
    when oracle/_ref was built."""
    import json
    import subprocess
    import time

    t0 = time.time()
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--workload", "knorm_128k", "--cpu-step-budget", "0.02"], capture_output=True, text=True, check=True)
    wall = time.time() - t0
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 3 and line["warmup"] == 1 and line["gpu_launches"] == 0
    assert line["steps"] * line["ms_per_step"] * 1e-3 < wall
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(line["config"]) >= {"workload", "S", "n_kept", "l2", "sharding"}
