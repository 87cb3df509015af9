"""N > 1 path of bench.py on CPU: world_size-2 gloo run of the rank aggregation (replicas only: MAX of times, SUM of work)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r: device leg took 100 + 50 r ms for 10 + r audio seconds, e2e leg 200 - 20 r ms for 5 audio seconds
    times, totals = bench.aggregate_over_ranks([100.0 + 50.0 * rank, 200.0 - 20.0 * rank], [10.0 + rank, 5.0], "cpu", world)
    out[rank] = (times, totals)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_aggregation_over_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        times, totals = out[r]
        assert times == [150.0, 200.0]          # max over ranks, per leg
        assert totals == [21.0, 10.0]           # whole-job audio seconds
    # value = whole-job audio seconds / max time, as bench.py reports it
    assert abs(out[0][1][0] / (out[0][0][0] * 1e-3) - 140.0) < 1e-9


def test_single_rank_is_identity():
    sys.path.insert(0, ROOT)
    import bench

    times, totals = bench.aggregate_over_ranks([12.5], [3.0], "cpu", 1)
    assert times == [12.5] and totals == [3.0]
