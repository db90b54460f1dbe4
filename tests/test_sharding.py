"""Host-side multi-GPU logic: frame shards, strip partition, max-over-ranks timing — including a real
world_size-2 run over the gloo backend (CPU)."""
import os
import socket

import pytest

from pixel_art_remaster_gpu_b200 import sharding


def test_frame_shards_tile_the_stream():
    for n in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            if n < world:
                continue
            cuts = [sharding.frame_shard(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_shard(8, 2, 2)


def test_strips_cover_image_with_exact_apron():
    assert sharding.APRON_ROWS >= 37  # SURVEY App. A.8 dependency radius
    strips = sharding.strip_rows(4096, 8)
    assert [s[:2] for s in strips] == [(512 * k, 512 * (k + 1)) for k in range(8)]
    assert strips[0][2] == 0 and strips[-1][3] == 4096
    for k, (b, e, lb, le) in enumerate(strips):
        assert lb == max(0, b - sharding.APRON_ROWS) and le == min(4096, e + sharding.APRON_ROWS)
    assert sharding.strip_rows(10, 1) == [(0, 10, 0, 10)]
    with pytest.raises(ValueError):
        sharding.strip_rows(3, 4)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = sharding.frame_shard(4097, rank, world)
        mine = [b, e, sharding.stream_seed(rank, 4096, 0xC0FFEE)]
        got = [None] * world
        dist.all_gather_object(got, mine)
        slow = sharding.max_over_ranks_ms(10.0 + 5.0 * rank, dist)
        q.put((rank, got, slow))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo():
    """world_size 2 on CPU: shards are disjoint and cover the stream, seeds do not collide, and the
    step time every rank reports is the slowest rank's."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, got, slow in res:
        assert got[0][:2] == [0, 2049] and got[1][:2] == [2049, 4097]
        assert got[1][2] - got[0][2] == 4096
        assert slow == 15.0
