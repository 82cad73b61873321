"""N>1 path on CPU (gloo, world_size 2): the data-parallel-replica bench aggregation (times MAX over ranks, token
counts SUM) and the per-rank workload sharding give the whole-job numbers the driver expects."""
import os
import socket

import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q) -> None:
    import torch.distributed as dist

    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r "measured" r+1 seconds and generated 100*(r+1) tokens
    maxed, summed = bench.reduce_over_ranks([1.0 + rank, 10.0 - rank], [100.0 * (rank + 1), 7.0], "cpu")
    q.put((rank, maxed, summed))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_over_ranks_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, maxed, summed in res:
        assert maxed == [2.0, 10.0] and summed == [300.0, 14.0]
    # whole-job throughput = all ranks' tokens / slowest rank's time
    assert summed[0] / maxed[0] == 150.0


def test_single_process_is_identity():
    import bench

    assert bench.reduce_over_ranks([3.0], [5.0], "cpu") == ([3.0], [5.0])


def test_reference_arm_json_contract_on_cpu():
    """`bench.py --impl reference` (the CPU port of the reference path on the host cores) must print ONE JSON line with
    the driver-contract keys: impl, the same metric / unit / higher_is_better as our arm, a cpu_baseline describing the
    run and an e2e block with zero copy bytes.  Run here on a tiny shape so that the contract is checked without a GPU."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--model", "tiny", "--batch", "4",
                        "--prompt-len", "32", "--gen-len", "8", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["metric"].startswith("decode tokens/sec")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
