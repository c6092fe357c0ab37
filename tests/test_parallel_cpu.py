"""Multi-process host logic of the batch sharding (gloo, world_size 2, CPU): slice bounds and the result gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stlpose_b200.parallel import gather_keypoints, shard_bounds


def test_shard_bounds_match_torch_chunk():
    for n in (0, 1, 2, 5, 32, 511, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            chunks = torch.arange(n).chunk(world) if n else []
            for r in range(world):
                lo, hi = shard_bounds(n, world, r)
                expect = chunks[r].tolist() if r < len(chunks) else []
                assert list(range(lo, hi)) == expect, (n, world, r)
            assert sum(shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)) == n
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        J = 17
        full_p = torch.arange(n * J * 2, dtype=torch.float32).view(n, J, 2)
        full_m = torch.arange(n * J, dtype=torch.float32).view(n, J, 1) * 0.5
        lo, hi = shard_bounds(n, world, rank)
        p, m = gather_keypoints(full_p[lo:hi].clone(), full_m[lo:hi].clone(), n)
        ok = torch.equal(p, full_p) and torch.equal(m, full_m)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8, 1])
def test_gather_keypoints_gloo_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
