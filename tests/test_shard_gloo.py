"""Scene sharding on CPU with the gloo backend, world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gauspcc_b200 import shard


def test_assign_scenes_lpt():
    sizes = [600, 100, 500, 300, 300, 200, 100, 100]
    a = shard.assign_scenes(sizes, 2)
    assert sorted(sum(a, [])) == list(range(8))
    loads = [sum(sizes[i] for i in r) for r in a]
    assert abs(loads[0] - loads[1]) <= 100
    assert shard.assign_scenes(sizes, 2) == a
    assert shard.assign_scenes([], 4) == [[], [], [], []]
    one = shard.assign_scenes(sizes, 1)
    assert sorted(one[0]) == list(range(8))


def test_block_ranges_and_pinning():
    for n, k in [(10, 3), (1_000_000, 8), (5, 5), (2, 5), (0, 4), (7, 1)]:
        r = shard.block_ranges(n, k)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert len(r) == max(1, min(k, max(n, 1))) and max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    assert shard.block_path("/a/b/xyz_pcc.bin", 3) == "/a/b/xyz_pcc_blk3.bin"
    before = os.sched_getaffinity(0)
    try:
        got = shard.pin_rank(0, 1)
        assert got == len(before) and os.sched_getaffinity(0) == before          # one rank keeps every CPU
        if len(before) >= 4:
            assert shard.pin_rank(1, 2) == len(before) // 2
            assert os.sched_getaffinity(0) == set(sorted(before)[len(before) // 2:2 * (len(before) // 2)])
    finally:
        os.sched_setaffinity(0, before)


def _worker(rank, world, port, sizes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a = shard.assign_scenes(sizes, world)
        mine = a[rank]
        # per-scene result rows: {n_points, file_bytes, scene id}
        local = torch.tensor([[sizes[i], sizes[i] * 7 + 3, i] for i in mine], dtype=torch.int64).reshape(-1, 3)
        out = shard.gather_results(local, len(sizes), a)
        q.put((rank, out.numpy()))
    finally:
        dist.destroy_process_group()


def test_gather_results_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    sizes = [600, 100, 500, 300, 250]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[n, n * 7 + 3, i] for i, n in enumerate(sizes)])
    assert np.array_equal(res[0], want) and np.array_equal(res[1], want)
