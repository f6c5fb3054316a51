"""The N>1 path on CPU: two gloo ranks shard a batch by image (no data-path collective) and
agree on the max-over-ranks step time the way bench.py does."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from picha_b200 import _native as N
    from picha_b200.shard import plan_batch, shard_range
    import bench
    n = 37
    # the batch of the 8-GPU thumbnail pipeline in small: runs of same-shape images with a few odd ones in between
    shapes = [(1920, 1080, 0)] * 14 + [(1921, 1080, 0)] + [(1920, 1080, 0)] * 9 + [(640, 480, 1)] * 13
    srcs = [N.CImage(1, w * N.lib.picha_b200_pixel_bytes(p), w, h, p) for (w, h, p) in shapes]
    dsts = [N.CImage(1, 256 * N.lib.picha_b200_pixel_bytes(p), 256, 256, p) for (_, _, p) in shapes]
    lo, hi = shard_range(n, rank, world)        # the library's own split (picha_b200_*_batch(..., device=-1))
    owned = torch.zeros(n, dtype=torch.int64)
    for first, count in plan_batch(srcs[lo:hi], dsts[lo:hi]):   # ... and its chunks of one launch each
        assert count >= 1 and len({shapes[lo + first + i] for i in range(count)}) == 1
        owned[lo + first:lo + first + count] += 1
    dist.all_reduce(owned)                      # test-only: every image is in exactly one chunk of one rank
    ms = bench.max_over_ranks(10.0 + rank)      # what bench.py reports as the step time
    total = bench.sum_over_ranks(hi - lo)
    q.put((rank, owned.tolist(), ms, total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_by_image_and_take_max_time():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, owned, ms, total in res:
        assert owned == [1] * 37
        assert ms == 11.0
        assert total == 37
