"""world_size-2 gloo test of the multi-GPU host logic (frame sharding, max-over-ranks timing, in-order merge)."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ros_gpu_stereo_processor_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_of_rank(11, rank, world)
    digests = [1000 + f for f in mine]                      # stands in for the per-frame results
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    merged = sharding.merge_in_frame_order(gathered, world)
    fps, ms = sharding.aggregate_throughput(len(mine), 10.0 * (rank + 1), dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, merged, fps, ms))


def test_two_rank_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, merged0, fps0, ms0), (r1, mine1, merged1, fps1, ms1) = res
    assert mine0 == [0, 2, 4, 6, 8, 10] and mine1 == [1, 3, 5, 7, 9]
    assert merged0 == merged1 == [1000 + i for i in range(11)]
    assert ms0 == ms1 == 20.0                                  # MAX over ranks
    assert abs(fps0 - 11 / 0.020) < 1e-6 and fps0 == fps1      # all frames over the slowest rank


def test_single_rank_paths():
    sys.path.insert(0, ROOT)
    from ros_gpu_stereo_processor_b200 import sharding
    assert sharding.frames_of_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.frames_of_rank(3, 2, 4) == [2] and sharding.frames_of_rank(2, 3, 4) == []
    assert sharding.merge_in_frame_order([[0, 3], [1], [2]], 3) == [0, 1, 2, 3]
    assert sharding.aggregate_throughput(16, 8.0)[0] == 2000.0
