"""fine-tuning step (train-mode forward, PersonMSELoss, backward, SGD step) on one GPU."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stlpose_b200 as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="32,64,128")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--width", type=int, default=32)
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as one CUDA graph (TrainStep)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:                                   # torchrun: one process per GPU, per-rank batch fixed (weak scaling)
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    crit = S.PersonMSELoss()
    flops = 45.8e9 if a.width == 32 else None
    for B in [int(b) for b in a.batches.split(",")]:
        torch.manual_seed(0)
        m = S.PoseHighResolutionNet(width=a.width).cuda().train()        # fresh per batch size (one TrainStep per optimizer)
        opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)      # model_setup.py:138-139
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, 3, 256, 192, device="cuda", generator=g)
        tgt = torch.rand(B, 17, 64, 48, device="cuda", generator=g)
        tw = torch.tensor([0.0, 1.0, 1.2, 1.5], device="cuda")[torch.randint(0, 4, (B, 17, 1), device="cuda", generator=g)]
        fwd = bwd = upd = None
        red = None
        if world > 1:
            from stlpose_b200.parallel import GradientReducer
            red = GradientReducer(m.parameters(), local_batch=B)
        if a.graph or red is not None:
            gstep = S.TrainStep(m, opt, crit, batch=B, reducer=red, use_graph=bool(a.graph))
            step = lambda: gstep(x, tgt, tw)
        else:
            def step():
                out = S.forward_pass(m, x, "HRNet", device="cuda", flip=False)
                loss = crit(out, tgt, tw)
                opt.zero_grad()
                loss.backward()
                opt.step()
                return loss.detach()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if not a.graph and world == 1:                                   # phase split on one extra eager step
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(); out = S.forward_pass(m, x, "HRNet", device="cuda", flip=False); loss = crit(out, tgt, tw)
            ev[1].record(); opt.zero_grad(); loss.backward(); ev[2].record(); opt.step(); ev[3].record()
            torch.cuda.synchronize()
            fwd, bwd, upd = (round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(3))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            loss = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        exch = None
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            # (exposed exchange time: bench.py --workload train measures it against a collective-free replay)
        line = dict(workload=f"hrnet_w{a.width}_256x192 train step (fwd + PersonMSELoss + bwd + SGD)", graph=bool(a.graph),
                    batch=B * world, n_gpus=world, ms_per_step=round(ms, 2), crops_per_s=round(B * world / ms * 1e3, 1),
                    allreduce_ms=exch, fwd_ms=fwd, bwd_ms=bwd, opt_ms=upd,
                    tflops=round(flops * B * world / ms / 1e9, 1) if flops else None, loss=float(loss))
        if rank == 0:
            print(json.dumps(line))
        del m, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
