"""Data-parallel fine-tuning step on 2 GPUs (NCCL): per-rank forward/backward on a slice of the batch, GradientReducer
all-reduce weighted by the slice sizes, identical SGD update on both ranks.  Needs >= 2 GPUs (gpurun --gpus 2);
skipped on the single-GPU box of the regular GPU tier."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import faulthandler
    import traceback
    import torch.distributed as dist
    faulthandler.dump_traceback_later(200, exit=True)          # a rank stuck in a collective: show where, then die
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _worker_body(rank, world, q, dist)
    except BaseException:
        q.put((rank, "error", traceback.format_exc()))
        q.close()
        q.join_thread()
        os._exit(1)                                            # do not wait for the peer in destroy_process_group
    dist.destroy_process_group()


def _worker_body(rank, world, q, dist):
    if True:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import test_train_gpu as T
        import stlpose_b200 as S
        from stlpose_b200.parallel import GradientReducer
        B, lr = 6, 0.05
        bounds = [(0, 4), (4, 6)]
        S_, sd0, x, tgt, tw, m = T._setup(B, seed=11)          # same checkpoint and batch on both ranks
        lo, hi = bounds[rank]
        opt = torch.optim.SGD(m.parameters(), lr=lr)
        red = GradientReducer(m.parameters(), local_batch=hi - lo)
        assert red.global_batch == B
        step = S.TrainStep(m, opt, S.PersonMSELoss(), batch=hi - lo, reducer=red)
        step(x[lo:hi], tgt[lo:hi], tw[lo:hi])
        torch.cuda.synchronize()
        names = ["conv1.weight", "layer1.0.conv1.weight", "stage3.1.branches.2.1.bn2.weight", "final_layer.weight"]
        sd = dict(m.named_parameters())
        same = True
        for n in names:
            mine = sd[n].detach().clone()
            both = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(both, mine)
            same = same and torch.equal(both[0], both[1])
        ok_ref = True
        errs = {}
        if rank == 0:                                          # expected update from the two slices' own gradients
            grads = []
            for (a, b) in bounds:
                ref = S.PoseHighResolutionNet(width=32)
                ref.load_state_dict(sd0, strict=True)
                ref = ref.cuda().train()
                S.PersonMSELoss()(ref(x[a:b].cuda()), tgt[a:b].cuda(), tw[a:b].cuda()).backward()
                grads.append({k: p.grad.clone() for k, p in ref.named_parameters()})
            for n in names:
                g = grads[0][n] * (4 / B) + grads[1][n] * (2 / B)
                want = sd0[n].cuda() - lr * g
                err = (sd[n].detach() - want).norm() / (lr * g).norm().clamp_min(1e-12)
                ok_ref = ok_ref and err.item() < 2e-2
                errs["ref:" + n] = round(err.item(), 5)
        # eager step with bucket all-reduces launched from autograd hooks (overlapping the rest of backward) must give
        # the same update as the graph-replayed step followed by reduce_all()
        m2 = S.PoseHighResolutionNet(width=32)
        m2.load_state_dict(sd0, strict=True)
        m2 = m2.cuda().train()
        opt2 = torch.optim.SGD(m2.parameters(), lr=lr)
        red2 = GradientReducer(m2.parameters(), local_batch=hi - lo, bucket_bytes=8 << 20).attach_hooks()
        assert len(red2.buckets) > 4
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            loss2 = S.PersonMSELoss()(m2(x[lo:hi].cuda()), tgt[lo:hi].cuda(), tw[lo:hi].cuda())
            opt2.zero_grad()
            loss2.backward()
            red2.finish_step()
            opt2.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        sd2 = dict(m2.named_parameters())
        ok_hooks = True
        for n in names:
            upd = (sd[n].detach() - sd0[n].cuda())
            err = (sd2[n].detach() - sd[n].detach()).norm() / upd.norm().clamp_min(1e-12)
            ok_hooks = ok_hooks and err.item() < 1e-3
            errs["hooks:" + n] = round(err.item(), 5)
        q.put((rank, bool(same), bool(ok_ref and ok_hooks), errs))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_train_step_matches_weighted_slice_gradients():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in procs:
            results.append(q.get(timeout=300))                 # a failing rank reports its traceback instead
            assert results[-1][1] != "error", results[-1][2]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert sorted(r[:3] for r in results) == [(0, True, True), (1, True, True)], results
