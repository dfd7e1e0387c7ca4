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
    torch.save({"job_ms": job_ms, "value": value, "sums": gathered}, os.path.join(out_dir, f"r{rank}.pt"))
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
