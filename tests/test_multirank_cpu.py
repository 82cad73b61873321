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
