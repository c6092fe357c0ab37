#!/usr/bin/env python
"""Benchmarks of the HRNet keypoint hot path (BASELINE.json configs 2-5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload infer|train|decode] [--width 32|48] [--batch B | --global-batch G]

One JSON line on stdout (rank 0).  Default = the configuration the headline metric is quoted on (config 2): HRNet-W32
256x192, 512 crops per GPU per step, flip test + get_final_preds decode.

  --workload infer  : person crops/s, forward (+ mirrored forward) + flip-average + decode.  A "step" is one pass over
                      one batch of synthetic crops.  --width 48 --global-batch 1024 is config 3 (strong scaling: the
                      global batch is split over the ranks).
  --workload train  : config 4, one fine-tuning step (train-mode forward, PersonMSELoss, backward, SGD) per step, data
                      parallel over the ranks with the NCCL gradient all-reduce inside the captured step; the line
                      carries `collective` = the exposed all-reduce time.
  --workload decode : config 5, fused flip-average + decode of resident heatmaps (HBM-bound), with the sweep over batch
                      1 Ki - 64 Ki and both heatmap sizes in `sweep`.
  value    : units/s with inputs resident in HBM, device-timed with CUDA events (max over ranks)
  e2e      : the same through the public call with pinned HOST buffers (H2D of the step's inputs, D2H of its result
             inside the timed region)
  roofline : the dominant kernel against the measured peak of MEASURED_PEAKS.json
  cpu_baseline : the reference's own modules (oracle/_ref, staged by oracle/stage_ref.py; `kind: "reference"`) - or the
             oracle port when they are absent (`kind: "port"`) - on this box's host cores, bounded sample
--impl reference times that CPU path alone (rank 0; the other ranks exit).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {32: (256, 192), 48: (384, 288)}
FLOPS_FWD = {32: 15.290007552e9, 48: 70.6132e9}   # 2*MACs of one forward at SHAPES[width] (oracle.conv_flops_per_crop)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    captures (profiles/ncu_traffic.json: one entry per workload / width, each naming its .md summary)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get(key)
        return (e["bytes_per_launch"], e["note"]) if e else (None, f"no ncu capture committed for {key}")
    except (OSError, ValueError, KeyError):
        return None, "profiles/ncu_traffic.json missing"


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.samples, self.proc, self.gpu_index = [], None, gpu_index

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.time(), line.strip()))
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, smax, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.samples:
            if t < t0 or t > t1:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                smax = max(smax, float(parts[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuPath:
    """The reference's CPU implementation of one workload.  kind "reference": the unmodified modules staged under
    oracle/_ref (models.HRnet.PoseHighResolutionNet, lib.inference.forward_pass, lib.pose_parsing.get_final_preds_hrnet,
    lib.loss.PersonMSELoss) through oracle/ref_shim.py; kind "port": the oracle restatement (same torch-CPU / NumPy calls)
    when the staged copy is absent."""

    def __init__(self, workload, width, threads=None):
        import torch
        from oracle import hrnet_oracle, pose_oracle, ref_shim
        torch.set_num_threads(threads or _host_threads())      # torchrun exports OMP_NUM_THREADS=1
        self.torch, self.ho, self.po = torch, hrnet_oracle, pose_oracle
        self.workload, self.width, self.image = workload, width, SHAPES[width]
        self.kind = "reference" if ref_shim.available() else "port"
        self.cores = torch.get_num_threads()
        self.sd = hrnet_oracle.synth_state_dict(width, seed=0)
        if self.kind == "reference":
            self.lib = ref_shim.lib()
            if workload != "decode":
                self.model = ref_shim.build_reference_hrnet(width, self.image)
                self.model.load_state_dict(self.sd, strict=True)
        if workload == "train" and self.kind == "reference":
            self.model.train()
            self.opt = torch.optim.SGD(self.model.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)
            self.crit = self.lib.loss.PersonMSELoss()
        elif workload == "infer" and self.kind == "reference":
            self.model.eval()

    def describe(self):
        what = {"infer": "forward_pass(flip=True) + get_final_preds_hrnet", "decode": "flip_back + average + get_final_preds_hrnet",
                "train": "train-mode forward + PersonMSELoss + backward + SGD step"}[self.workload]
        who = ("unmodified reference modules (oracle/_ref: models/HRnet.py, lib/inference.py, lib/pose_parsing.py, "
               "lib/loss.py)") if self.kind == "reference" else "oracle port (same torch-CPU fp32 / NumPy calls)"
        return f"{what}, {who}, fp32"

    def inputs(self, n):
        torch, po = self.torch, self.po
        g = torch.Generator().manual_seed(0)
        H, W = self.image
        if self.workload == "decode":
            heat = po.blob_heatmaps(n, 17, H // 4, W // 4, seed=1, noise=0.01).astype("float32")
            heat_f = heat[:, :, :, ::-1].copy()
            c, s = po.synth_boxes(n, seed=0)
            return heat, heat_f, c, s
        x = torch.randn(n, 3, H, W, generator=g)
        if self.workload == "train":
            tgt = torch.from_numpy(po.blob_heatmaps(n, 17, H // 4, W // 4, seed=1, noise=0.0)).float()
            tw = torch.ones(n, 17, 1)
            return x, tgt, tw
        c, s = po.synth_boxes(n, seed=0)
        return x, c, s

    def step(self, inp):
        torch, ho, po = self.torch, self.ho, self.po
        if self.workload == "infer":
            x, c, s = inp
            if self.kind == "reference":
                with torch.no_grad():
                    out = self.lib.inference.forward_pass(self.model, x, "HRNet", device="cpu", flip=True)
                return self.lib.pose_parsing.get_final_preds_hrnet(out.numpy(), c, s)
            h0 = ho.hrnet_forward(self.sd, x, self.width).numpy()
            h1 = ho.hrnet_forward(self.sd, x.flip(3), self.width).numpy()
            return po.get_final_preds(po.flip_average(h0, h1), c, s)
        if self.workload == "decode":
            heat, heat_f, c, s = inp
            if self.kind == "reference":
                fb = self.lib.transforms.flip_back(torch.from_numpy(heat_f), self.lib.CONSTANTS.FLIP_PAIRS)
                fb[:, :, :, 1:] = fb.clone()[:, :, :, 0:-1]                       # lib/inference.py:25
                avg = ((torch.from_numpy(heat) + fb) * 0.5).numpy()
                return self.lib.pose_parsing.get_final_preds_hrnet(avg, c, s)
            return po.get_final_preds(po.flip_average(heat, heat_f), c, s)
        x, tgt, tw = inp
        if self.kind == "reference":
            out = self.lib.inference.forward_pass(self.model, x, "HRNet", device="cpu", flip=False)
            loss = self.crit(out, tgt, tw)
            self.opt.zero_grad()
            loss.backward()
            self.opt.step()
            return float(loss.detach())
        sd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone())
              for k, v in self.sd.items()}
        heat = ho.hrnet_forward_train(sd, x, self.width)
        d = (heat - tgt).reshape(x.shape[0], 17, -1) * tw
        loss = 0.5 * (d * d).mean(dim=(0, 2)).sum() / 17
        loss.backward()
        with torch.no_grad():
            for v in sd.values():
                if v.grad is not None:
                    v -= 1e-3 * v.grad
        return float(loss)

    def time(self, n, passes):
        inp = self.inputs(n)
        small = self.inputs(min(n, 2))
        self.step(small)                                     # warm-up (thread pool, lazy init)
        best = float("inf")
        for _ in range(passes):
            t = time.perf_counter()
            self.step(inp)
            best = min(best, time.perf_counter() - t)
        return n / best, best


CPU_SAMPLE = {("infer", 32): 32, ("infer", 48): 8, ("train", 32): 16, ("train", 48): 4, ("decode", 32): 2048,
              ("decode", 48): 1024}
METRIC = {"infer": "person crops/sec HRNet-W{w} {h}x{v} fwd+flip+decode",
          "train": "person crops/sec HRNet-W{w} {h}x{v} fine-tuning step (fwd + PersonMSELoss + bwd + SGD)",
          "decode": "person crops/sec heatmap decode {hh}x{hv} (flip-average + get_final_preds)"}
WORKLOAD = {"infer": "HRNet-W{w} {h}x{v} inference + flip-test + get_final_preds decode",
            "train": "HRNet-W{w} {h}x{v} fine-tuning step (train-mode BatchNorm, PersonMSELoss fwd/bwd, SGD), data-parallel",
            "decode": "heatmap decode only: 17-joint {hh}x{hv} heatmaps, flip-average + argmax + refinement + back-projection"}


def _names(args):
    H, W = SHAPES[args.width]
    f = dict(w=args.width, h=H, v=W, hh=H // 4, hv=W // 4)
    return METRIC[args.workload].format(**f), WORKLOAD[args.workload].format(**f)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    metric, workload = _names(args)
    cpu = CpuPath(args.workload, args.width)
    sample = args.cpu_sample or CPU_SAMPLE[(args.workload, args.width)]
    inp, small = cpu.inputs(sample), cpu.inputs(min(sample, 2))
    for _ in range(args.warmup):
        cpu.step(small)
    times = []
    for _ in range(args.steps):
        t = time.perf_counter()
        cpu.step(inp)
        times.append(time.perf_counter() - t)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    desc = f"{sample} crops per step: {cpu.describe()}"
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "crops_per_step": sample, "device": "host CPU",
                   "note": "a bounded sample of the GPU arm's workload on the host cores, same metric (32 crops = BASELINE.json "
                           "config 1's batch for the default workload)"},
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cpu.cores, "kind": cpu.kind, "sample": desc},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
class Harness:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        try:
            gpu_id = "GPU-" + str(torch.cuda.get_device_properties(self.local).uuid)
        except Exception:
            gpu_id = str(self.local)
        self.sampler = ClockSampler(gpu_id)
        self.sampler.start()
        time.sleep(0.3)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item() / steps, t0, t1

    def base_line(self, metric, workload, value, ms_step, clocks, dtype, scaling):
        a = self.args
        return {"metric": metric, "value": value, "unit": "crops/s", "n_gpus": self.world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": {"workload": workload},
                "clocks": clocks}

    def cpu_baseline(self, out):
        a = self.args
        if self.world != 1 or a.no_cpu_baseline:
            return
        cpu = CpuPath(a.workload, a.width)
        sample = a.cpu_sample or CPU_SAMPLE[(a.workload, a.width)]
        v, dt = cpu.time(sample, passes=2)
        out["cpu_baseline"] = {"value": v, "unit": "crops/s", "cores": cpu.cores, "kind": cpu.kind,
                               "sample": f"{sample} crops, best of 2 passes ({dt:.1f} s each): {cpu.describe()}"}

    def finish(self, out, ops=None):
        if self.rank == 0:
            print(json.dumps(out))
            if self.args.dump_ops and ops is not None:
                with open(self.args.dump_ops, "w") as f:
                    json.dump(ops, f)
        self.sampler.stop()
        if self.world > 1:
            self.dist.destroy_process_group()


def per_rank_batch(args, world, rank):
    """(this rank's crops per step, total crops per step, scaling).  --global-batch: contiguous DataParallel-style slices
    (parallel.shard_bounds) of a fixed global batch = strong scaling; else --batch per GPU = weak scaling."""
    if args.global_batch:
        from stlpose_b200.parallel import shard_bounds
        lo, hi = shard_bounds(args.global_batch, world, rank)
        return hi - lo, args.global_batch, "strong"
    return args.batch, args.batch * world, "weak"


def run_infer(args):
    import stlpose_b200 as S
    from stlpose_b200.pipeline import KeypointPipeline
    h = Harness(args)
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    IMAGE, WIDTH = SHAPES[args.width], args.width
    B, total, scaling = per_rank_batch(args, world, rank)
    Bmax = -(-total // world)

    # BASELINE.json config 1/2: random-init weights, torch.manual_seed(0), default nn.Conv2d init, BatchNorm (1, 0, 0, 1)
    torch.manual_seed(0)
    model = S.PoseHighResolutionNet(width=WIDTH, image_size=IMAGE).to(dev).eval()
    pipe = KeypointPipeline(model, B, IMAGE, flip=True, use_graph=not args.no_graph)

    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.randn(B, 3, *IMAGE, generator=gen).pin_memory()
    c_host = (torch.rand(B, 2, generator=gen) * torch.tensor([400.0, 300.0]) + 100.0).pin_memory()   # box centres
    hgt = torch.rand(B, 1, generator=gen) * 320.0 + 80.0                                                # box heights
    s_host = (torch.cat([0.75 * hgt, hgt], dim=1) / 200.0 * 1.25).pin_memory()                          # _xywh2cs scale
    p_host = torch.empty(B, 17, 2).pin_memory()
    m_host = torch.empty(B, 17, 1).pin_memory()
    pipe.x.copy_(x_host); pipe.center.copy_(c_host); pipe.scale.copy_(s_host)

    # the only exchange of the sharded pipeline: every rank ends up with all keypoints (204 B per crop)
    gathered = [torch.empty((Bmax, 17, 3), device=dev) for _ in range(world)] if world > 1 else None
    packed = torch.zeros((Bmax, 17, 3), device=dev) if world > 1 else None

    def gather_results():
        if world > 1:
            packed[:B, :, :2].copy_(pipe.preds)
            packed[:B, :, 2:].copy_(pipe.maxvals)
            dist.all_gather(gathered, packed)

    def step_resident():
        pipe.step()
        gather_results()

    def step_e2e():
        # the public call: pinned host crops/boxes in, pinned host keypoints out (H2D of the next batch overlaps
        # the current pass inside KeypointPipeline)
        pipe(x_host, c_host, s_host, p_host, m_host)
        gather_results()

    ms_step, t0, t1 = h.timed(step_resident, args.steps, args.warmup)
    clocks = h.sampler.summary(t0, t1)
    ms_e2e, _, _ = h.timed(step_e2e, args.steps, max(args.warmup, 3))

    # live per-kernel timing of the dominant kernels (tcgen05 convolutions) over one step, CUDA events on the launch stream
    ops = model.profile_ops(pipe.x, flip_pair=True)
    tc_kinds = ("conv_tc", "block_tc", "link_tc")   # block_tc / link_tc = two convolutions fused in one kernel
    conv_ms = sum(o["ms"] for o in ops if o["kind"] in tc_kinds)
    conv_flops = sum(o["flops"] for o in ops if o["kind"] in tc_kinds)
    conv_n = sum(1 for o in ops if o["kind"] in tc_kinds)
    other_ms = sum(o["ms"] for o in ops if o["kind"] not in tc_kinds)
    peaks, peak_kind = measured_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

    out = None
    if rank == 0:
        metric, workload = _names(args)
        value, e2e = total / (ms_step * 1e-3), total / (ms_e2e * 1e-3)
        h2d = x_host.numel() * 4 + c_host.numel() * 4 + s_host.numel() * 4
        d2h = p_host.numel() * 4 + m_host.numel() * 4
        traffic, traffic_note = ncu_traffic(f"infer_w{WIDTH}")
        out = h.base_line(metric, workload, value, ms_step, clocks, "bf16", scaling)
        out["config"].update({
            "crops_per_step_total": total, "crops_per_gpu_per_step": B, "forwards_per_crop": 2,
            "cuda_graph": pipe.graph is not None,
            "l2": f"inputs ({B * 3 * IMAGE[0] * IMAGE[1] * 4 / 1e6:.0f} MB of crops per step) and activations exceed the 126 MB L2; no flush needed",
            "partition": f"batch sharded over {world} GPU(s); only exchange = all-gather of keypoints "
                         f"(204 B/crop, inside the timed step when n_gpus > 1)"})
        out["e2e"] = {"value": e2e, "unit": "crops/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h}
        out["gpu_launches"] = pipe.launches_per_step * args.steps
        out["roofline"] = {
            "bound": "tensor", "kernel": "conv_tc_kernel + basic_block_kernel + bottleneck_link_kernel (tcgen05 convolutions)",
            "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)",
            "launches_per_step": conv_n, "conv_ms_per_step": conv_ms, "other_kernels_ms": other_ms,
            "whole_step_frac": (2 * FLOPS_FWD[WIDTH] * B / (ms_step * 1e-3) / 1e12) / peak_tf}
        h.cpu_baseline(out)
    if not args.no_extras:
        # config 4 next to the headline, so that the scaling runs (which use the default command line) also record the
        # data-parallel fine-tuning step: 32 crops per GPU, gradients all-reduced inside the captured step
        del pipe, model
        torch.cuda.empty_cache()
        r = measure_train(h, args, steps=10, warmup=3, e2e=False)
        if out is not None:
            out["train"] = {
                "metric": "person crops/sec fine-tuning step (fwd + PersonMSELoss + bwd + SGD), HRNet-W%d" % WIDTH,
                "value": r["total"] / (r["ms_step"] * 1e-3), "unit": "crops/s", "ms_per_step": r["ms_step"],
                "crops_per_gpu_per_step": r["B"], "n_gpus": world, "scaling": "weak", "steps": 10,
                "tensor_frac": r["achieved_tf"] / r["peak_tf"], "kernels_per_step": r["kernel_nodes"],
                "collective": r["collective"], "loss": r["loss"]}
    h.finish(out, ops)


def measure_train(h, args, steps, warmup, e2e=True):
    """One data-parallel fine-tuning step per step (config 4) on the ranks of harness h -> dict of measurements (rank 0
    fills in the derived numbers).  Used by --workload train and, briefly, as the `train` object of the default line."""
    import stlpose_b200 as S
    from stlpose_b200.parallel import GradientReducer
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    IMAGE, WIDTH = SHAPES[args.width], args.width
    B, total, scaling = per_rank_batch(args, world, rank) if args.workload == "train" else (32, 32 * world, "weak")
    H, W = IMAGE
    crit = S.PersonMSELoss()

    def build(dry):
        torch.manual_seed(0)
        m = S.PoseHighResolutionNet(width=WIDTH, image_size=IMAGE).to(dev).train()
        opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)      # lib/model_setup.py:138-139
        red = GradientReducer(m.parameters(), local_batch=B) if world > 1 else None
        if red is not None and dry:
            red.world = 1                       # same stream choreography, no ncclAllReduce: the collective-free step
        return m, S.TrainStep(m, opt, crit, batch=B, image_size=IMAGE, reducer=red, use_graph=not args.no_graph,
                               count_kernels=not dry)

    model, step = build(dry=False)
    gen = torch.Generator().manual_seed(2000 + rank)
    x_host = torch.randn(B, 3, H, W, generator=gen).pin_memory()
    t_host = torch.rand(B, 17, H // 4, W // 4, generator=gen).pin_memory()
    w_host = torch.tensor([0.0, 1.0, 1.2, 1.5])[torch.randint(0, 4, (B, 17, 1), generator=gen)].pin_memory()
    loss_host = torch.empty(()).pin_memory()
    x_dev, t_dev, w_dev = x_host.to(dev), t_host.to(dev), w_host.to(dev)

    def step_resident():
        step(x_dev, t_dev, w_dev)

    def step_e2e():                              # pinned host batch in (H2D inside the call), loss scalar out
        loss_host.copy_(step(x_host, t_host, w_host), non_blocking=True)

    ms_step, t0, t1 = h.timed(step_resident, steps, warmup)
    clocks = h.sampler.summary(t0, t1)
    ms_e2e = h.timed(step_e2e, steps, max(warmup, 3))[0] if e2e else None
    r = {"B": B, "total": total, "scaling": scaling, "ms_step": ms_step, "ms_e2e": ms_e2e, "clocks": clocks,
         "loss": float(step.loss), "kernel_nodes": step.kernel_nodes, "graph": step.graph is not None,
         "optimizer_in_graph": step.optimizer_in_graph, "collective": None,
         "h2d": (x_host.numel() + t_host.numel() + w_host.numel()) * 4}
    if world > 1:
        n_bytes = sum(f.numel() for f in step.reducer.flat) * 4
        n_buckets = len(step.reducer.flat)
        del step, model
        torch.cuda.empty_cache()
        model, step = build(dry=True)            # the same step without the collective: the difference is what is exposed
        ms_dry, _, _ = h.timed(step_resident, steps, warmup)
        r["collective"] = {"op": "ncclAllReduce (fp32 gradient buckets, nodes of the captured step, on a communication "
                                 "stream forked when backward has filled a bucket)", "bytes_per_step": n_bytes,
                           "buckets": n_buckets, "ms_per_step_without_collective": ms_dry,
                           "exposed_ms": ms_step - ms_dry}
    del step, model
    torch.cuda.empty_cache()
    peaks, peak_kind = measured_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
    r["achieved_tf"] = 3 * FLOPS_FWD[WIDTH] * B / (ms_step * 1e-3) / 1e12     # forward + dgrad + wgrad of every conv
    r["peak_tf"], r["peak_kind"] = peak_tf, peak_kind
    return r


def run_train(args):
    h = Harness(args)
    r = measure_train(h, args, args.steps, args.warmup)
    out = None
    if h.rank == 0:
        WIDTH = args.width
        metric, workload = _names(args)
        traffic, traffic_note = ncu_traffic(f"train_w{WIDTH}")
        out = h.base_line(metric, workload, r["total"] / (r["ms_step"] * 1e-3), r["ms_step"], r["clocks"], "bf16", r["scaling"])
        out["config"].update({
            "crops_per_step_total": r["total"], "crops_per_gpu_per_step": r["B"],
            "optimizer": "SGD momentum 0.9 weight decay 5e-4", "cuda_graph": r["graph"],
            "optimizer_in_graph": r["optimizer_in_graph"],
            "l2": "activations of a step (150 MB per crop) exceed the 126 MB L2; no flush needed",
            "partition": f"batch sharded over {h.world} GPU(s), per-rank BatchNorm statistics, gradients all-reduced"})
        out["e2e"] = {"value": r["total"] / (r["ms_e2e"] * 1e-3), "unit": "crops/s", "ms_per_step": r["ms_e2e"],
                      "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4}
        out["gpu_launches"] = r["kernel_nodes"] * args.steps if r["kernel_nodes"] else None   # kernel nodes of the captured step
        out["loss"] = r["loss"]
        out["roofline"] = {
            "bound": "tensor", "kernel": "whole step: conv_tc (forward, dgrad) + wgrad_tc + BatchNorm kernels",
            "achieved": r["achieved_tf"], "peak": r["peak_tf"], "unit": "TFLOP/s", "frac": r["achieved_tf"] / r["peak_tf"],
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": f"{r['peak_kind']} bf16_tflops_sustained", "flops_per_crop": 3 * FLOPS_FWD[WIDTH]}
        if r["collective"]:
            out["collective"] = r["collective"]
        h.cpu_baseline(out)
    h.finish(out)


def run_decode(args):
    import stlpose_b200 as S  # noqa: F401
    from stlpose_b200 import pose_parsing
    from stlpose_b200.transforms import FLIP_PAIRS
    h = Harness(args)
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    H, W = SHAPES[args.width]
    hh, hw = H // 4, W // 4
    J = 17
    B = args.batch if args.batch != 512 else 16384      # config 5 headline point: 16 Ki crops per GPU
    total, scaling = B * world, "weak"
    gen = torch.Generator(device=dev).manual_seed(rank)

    def make(b, hh_, hw_):
        heat = torch.randn(b, J, hh_, hw_, device=dev, generator=gen)
        heat2 = torch.randn(b, J, hh_, hw_, device=dev, generator=gen)
        c = torch.rand(b, 2, device=dev, generator=gen) * 300 + 100
        s = torch.rand(b, 2, device=dev, generator=gen) * 2 + 0.5
        return heat, heat2, c, s

    heat, heat2, c, s = make(B, hh, hw)
    per_crop = 2 * J * hh * hw * 4 + 340                 # both heatmap sets read once, keypoints written

    def step_resident():
        pose_parsing._decode(heat, c, s, True, heat_flipped=heat2, pairs=FLIP_PAIRS)

    ms_step, t0, t1 = h.timed(step_resident, args.steps, args.warmup)
    clocks = h.sampler.summary(t0, t1)
    # end to end: a bounded slice of the heatmaps from pinned host memory (PCIe-bound by construction)
    Be = min(B, 2048)
    hh_host, hf_host = heat[:Be].cpu().pin_memory(), heat2[:Be].cpu().pin_memory()
    c_host, s_host = c[:Be].cpu().pin_memory(), s[:Be].cpu().pin_memory()
    stage = [torch.empty_like(heat[:Be]), torch.empty_like(heat2[:Be]), torch.empty_like(c[:Be]), torch.empty_like(s[:Be])]
    p_host = torch.empty(Be, J, 2).pin_memory()

    def step_e2e():
        for d, src in zip(stage, (hh_host, hf_host, c_host, s_host)):
            d.copy_(src, non_blocking=True)
        preds, maxvals, _, _ = pose_parsing._decode(stage[0], stage[2], stage[3], True, heat_flipped=stage[1], pairs=FLIP_PAIRS)
        p_host.copy_(preds, non_blocking=True)

    ms_e2e, _, _ = h.timed(step_e2e, args.steps, max(args.warmup, 3))
    sweep = []
    if rank == 0 and not args.no_sweep:
        for (a_, b_) in ((64, 48), (96, 72)):
            for bb in (1024, 4096, 16384, 65536):
                t_heat, t_heat2, t_c, t_s = make(bb, a_, b_)
                per = J * a_ * b_ * 4
                for name, fn, nbytes in (
                        ("decode", lambda: pose_parsing._decode(t_heat, t_c, t_s, True), per + 340),
                        ("flipavg+decode", lambda: pose_parsing._decode(t_heat, t_c, t_s, True, heat_flipped=t_heat2,
                                                                        pairs=FLIP_PAIRS), 2 * per + 340)):
                    for _ in range(3):
                        fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(10):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    t = e0.elapsed_time(e1) / 10 * 1e-3
                    sweep.append({"heatmap": f"{a_}x{b_}", "batch": bb, "op": name, "ms": t * 1e3,
                                  "crops_per_s": bb / t, "gbps": nbytes * bb / t / 1e9})
                del t_heat, t_heat2
                torch.cuda.empty_cache()
    out = None
    if rank == 0:
        metric, workload = _names(args)
        peaks, peak_kind = measured_peaks()
        peak = peaks["hbm_gbs"]
        gbps = per_crop * B / (ms_step * 1e-3) / 1e9
        traffic, traffic_note = ncu_traffic(f"decode_{hh}x{hw}")
        for r in sweep:
            r["frac"] = r["gbps"] / peak
        out = h.base_line(metric, workload, total / (ms_step * 1e-3), ms_step, clocks, "f32", scaling)
        out["config"].update({"crops_per_gpu_per_step": B, "heatmap": f"{hh}x{hw}", "joints": J,
                              "l2": f"{per_crop * B / 1e6:.0f} MB of heatmaps per step exceed the 126 MB L2; no flush needed",
                              "partition": f"independent crops, {world} GPU(s), no collective"})
        out["e2e"] = {"value": Be * world / (ms_e2e * 1e-3), "unit": "crops/s", "ms_per_step": ms_e2e,
                      "crops_per_gpu_per_step": Be, "h2d_bytes_per_step": (hh_host.numel() * 2 + 4 * Be) * 4,
                      "d2h_bytes_per_step": p_host.numel() * 4}
        out["gpu_launches"] = args.steps
        out["roofline"] = {"bound": "hbm", "kernel": "decode_kernel (flip-average + arg-max + refinement + affine)",
                           "achieved": gbps, "peak": peak, "unit": "GB/s", "frac": gbps / peak, "traffic": traffic,
                           "traffic_note": traffic_note, "peak_source": f"{peak_kind} hbm_gbs (copy kernel)",
                           "bytes_per_crop": per_crop}
        out["sweep"] = sweep
        h.cpu_baseline(out)
    h.finish(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "decode"])
    ap.add_argument("--batch", type=int, default=512,
                    help="crops per GPU per step (config 2: 512; train default 32 = the reference's batch; decode 16384)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="fixed global batch split over the ranks (strong scaling; config 3: --width 48 --global-batch 1024)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="crops in the bounded CPU sample (default per workload)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="decode workload: skip the batch / heatmap-size sweep")
    ap.add_argument("--no-extras", action="store_true", help="infer workload: skip the short fine-tuning measurement (`train`)")
    ap.add_argument("--width", type=int, default=32, choices=[32, 48],
                    help="32: HRNet-W32 256x192 (the configuration the metric is quoted on); 48: HRNet-W48 384x288")
    ap.add_argument("--dump-ops", default="", help="infer workload: write the per-launch timing table (JSON) here")
    args = ap.parse_args()
    if os.environ.get("STL_BENCH_WATCHDOG"):             # debugging aid: dump every thread's stack and exit if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["STL_BENCH_WATCHDOG"]), exit=True)
    if args.workload == "train" and args.batch == 512 and not args.global_batch:
        args.batch = 32                                  # BASELINE config 4 / the reference's default batch per GPU
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 and the OpenMP pool is sized when torch is first imported: give the
        # CPU arm every core this process may run on BEFORE that import
        os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = str(_host_threads())
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    {"infer": run_infer, "train": run_train, "decode": run_decode}[args.workload](args)


if __name__ == "__main__":
    main()
