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
        # in a multi-process job the training path keeps the two-launch BatchNorm kernels (no cooperative grid spinning
        # next to in-flight collectives): training._use_coop_bn, "auto" policy
        from stlpose_b200 import training
        if training.COOP_BN == "auto":
            ok = ok and not training._use_coop_bn(torch.empty(2, 9, 7, 32, dtype=torch.bfloat16), backward=True)
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


def _grad_worker(rank, world, port, hooks, q):
    """Each rank: tiny CPU conv net, its slice of an 8-sample batch (5 / 3 split), local mean loss, GradientReducer."""
    from stlpose_b200.parallel import GradientReducer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1))
        unused = torch.nn.Parameter(torch.ones(3))                        # a parameter that gets no gradient
        params = list(net.parameters()) + [unused]
        x = torch.randn(8, 3, 6, 5)
        y = torch.randn(8, 4, 6, 5)
        lo, hi = (0, 5) if rank == 0 else (5, 8)
        red = GradientReducer(params, local_batch=hi - lo, bucket_bytes=64)    # several buckets
        assert red.global_batch == 8 and len(red.buckets) > 1
        if hooks:
            red.attach_hooks()
        for _ in range(2):                                                 # two steps: state is reset between them
            for p in params:
                p.grad = None
            loss = 0.5 * ((net(x[lo:hi]) - y[lo:hi]) ** 2).mean()
            loss.backward()
            if hooks:
                red.finish_step()
            else:
                red.reduce_all()
        # single-process reference: one loss over the whole batch (what DataParallel + loss.py:87 compute)
        ref = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1))
        ref.load_state_dict(net.state_dict())
        (0.5 * ((ref(x) - y) ** 2).mean()).backward()
        ok = all(torch.allclose(p.grad, r.grad, atol=1e-6) for p, r in zip(net.parameters(), ref.parameters()))
        ok = ok and unused.grad is not None and float(unused.grad.abs().max()) == 0.0
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("hooks", [False, True])
def test_gradient_reducer_matches_whole_batch_loss_gloo_world2(hooks):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, hooks, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
