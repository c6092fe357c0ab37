#!/usr/bin/env python
"""Headline benchmark: person crops/sec, HRNet-W32 256x192 forward + flip test + get_final_preds decode.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of B synthetic person
crops per GPU: two network forwards per crop (plain + mirrored, lib/inference.py:18-22), flip-average and decode.
  value    : crops/s with inputs resident in HBM, device-timed with CUDA events (max over ranks)
  e2e      : the same through the KeypointPipeline call path with pinned HOST buffers (H2D of the crops + boxes,
             D2H of the keypoints inside the timed region)
  roofline : the tcgen05 conv kernel vs the measured bf16 tensor peak (MEASURED_PEAKS.json)
  cpu_baseline : the CPU oracle port of the reference path on this box's host cores (bounded sample)
--impl reference times that CPU port instead (the reference's own CPU path cannot travel to the GPU box; the
oracle executes the same torch/NumPy library calls, see oracle/).
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

WIDTH, IMAGE = 32, (256, 192)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the most expensive kernel shape, from the committed
# `ncu --set full` capture profiles/r01_ncu_conv.md (prof_block): 208.92 MB read + 167.07 MB written
NCU_TRAFFIC_BYTES_PER_LAUNCH = 376.0e6
NCU_TRAFFIC_NOTE = ("basic_block_kernel (two 32->32 3x3 convs @64x48 + residual fused, 1024 images; 32 launches per "
                    "forward, 20 % of step time); algorithmic bytes of that launch: 417 MB (x read once, y written; "
                    "the residual re-read of x hits L2).  The unfused conv_tc_kernel<3,2,9,staged> launch it replaces: "
                    "594.9 MB measured / 604 MB algorithmic")
FLOPS_PER_FORWARD = 15.290007552e9   # HRNet-W32 @256x192, 2*MACs over the 293 convs (oracle.conv_flops_per_crop)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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


def cpu_port_crops_per_sec(n_crops, repeats=1, threads=None):
    """Time the oracle port of the reference path (forward_pass(flip=True) + get_final_preds_hrnet) on host cores."""
    import torch
    from oracle import hrnet_oracle, pose_oracle
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it is allowed to run on
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(threads or avail)
    sd = hrnet_oracle.synth_state_dict(WIDTH, seed=0)
    x = torch.randn(n_crops, 3, *IMAGE, generator=torch.Generator().manual_seed(0))
    center, scale = pose_oracle.synth_boxes(n_crops, seed=0)

    def one(xx, c, s):
        h0 = hrnet_oracle.hrnet_forward(sd, xx, WIDTH).numpy()
        h1 = hrnet_oracle.hrnet_forward(sd, xx.flip(3), WIDTH).numpy()
        return pose_oracle.get_final_preds(pose_oracle.flip_average(h0, h1), c, s)

    one(x[:2], center[:2], scale[:2])  # warm-up
    best = float("inf")
    for _ in range(repeats):
        t = time.perf_counter()
        one(x, center, scale)
        best = min(best, time.perf_counter() - t)
    return n_crops / best, best, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample
    times = []
    for _ in range(args.warmup):
        cpu_port_crops_per_sec(min(sample, 4))
    cores = os.cpu_count()
    for _ in range(args.steps):
        v, dt, cores = cpu_port_crops_per_sec(sample)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    desc = f"{sample} crops per step (oracle port: torch-CPU fp32 HRNet-W{WIDTH} x2 + NumPy flip-average/decode)"
    print(json.dumps({
        "impl": "reference", "metric": f"person crops/sec HRNet-W{WIDTH} {IMAGE[0]}x{IMAGE[1]} fwd+flip+decode", "value": value,
        "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"HRNet-W{WIDTH} {IMAGE[0]}x{IMAGE[1]} inference + flip-test + get_final_preds decode",
                   "crops_per_step": sample, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import stlpose_b200 as S
    from stlpose_b200.pipeline import KeypointPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    # BASELINE.json config 1/2: random-init weights, torch.manual_seed(0), default nn.Conv2d init, BatchNorm (1, 0, 0, 1)
    torch.manual_seed(0)
    model = S.PoseHighResolutionNet(width=WIDTH, image_size=IMAGE)
    model = model.to(dev).eval()
    pipe = KeypointPipeline(model, B, IMAGE, flip=True, use_graph=not args.no_graph)

    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.randn(B, 3, *IMAGE, generator=gen).pin_memory()
    c_host = (torch.rand(B, 2, generator=gen) * torch.tensor([400.0, 300.0]) + 100.0).pin_memory()   # box centres
    hgt = torch.rand(B, 1, generator=gen) * 320.0 + 80.0                                                # box heights
    s_host = (torch.cat([0.75 * hgt, hgt], dim=1) / 200.0 * 1.25).pin_memory()                          # _xywh2cs scale
    p_host = torch.empty(B, 17, 2).pin_memory()
    m_host = torch.empty(B, 17, 1).pin_memory()
    pipe.x.copy_(x_host); pipe.center.copy_(c_host); pipe.scale.copy_(s_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the only exchange of the sharded pipeline: every rank ends up with all keypoints (204 B per crop)
    gathered = [torch.empty((B, 17, 3), device=dev) for _ in range(world)] if world > 1 else None
    packed = torch.empty((B, 17, 3), device=dev) if world > 1 else None

    def gather_results():
        if world > 1:
            packed[..., :2].copy_(pipe.preds)
            packed[..., 2:].copy_(pipe.maxvals)
            dist.all_gather(gathered, packed)

    def step_resident():
        pipe.step()
        gather_results()

    def step_e2e():
        # the public call: pinned host crops/boxes in, pinned host keypoints out (H2D of the next batch overlaps
        # the current pass inside KeypointPipeline)
        pipe(x_host, c_host, s_host, p_host, m_host)
        gather_results()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, t0, t1

    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_id = str(local)
    sampler = ClockSampler(gpu_id)
    sampler.start()
    time.sleep(0.3)
    # device-resident throughput
    ms_step, t0, t1 = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.summary(t0, t1)
    # end to end through the public call with host buffers
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
    sampler.stop()

    # live per-kernel timing of the dominant kernel (tcgen05 conv) over one step, CUDA events on the launch stream
    ops = model.profile_ops(pipe.x, flip_pair=True)
    tc_kinds = ("conv_tc", "block_tc")   # the tcgen05 convolution kernels (block_tc = two convs of a BasicBlock fused)
    conv_ms = sum(o["ms"] for o in ops if o["kind"] in tc_kinds)
    conv_flops = sum(o["flops"] for o in ops if o["kind"] in tc_kinds)
    conv_n = sum(1 for o in ops if o["kind"] in tc_kinds)
    other_ms = sum(o["ms"] for o in ops if o["kind"] not in tc_kinds)
    peaks, peak_kind = measured_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

    out = None
    if rank == 0:
        crops = B * world
        value = crops / (ms_step * 1e-3)
        e2e = crops / (ms_e2e * 1e-3)
        h2d = x_host.numel() * 4 + c_host.numel() * 4 + s_host.numel() * 4
        d2h = p_host.numel() * 4 + m_host.numel() * 4
        out = {
            "metric": f"person crops/sec HRNet-W{WIDTH} {IMAGE[0]}x{IMAGE[1]} fwd+flip+decode", "value": value, "unit": "crops/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"HRNet-W{WIDTH} {IMAGE[0]}x{IMAGE[1]} inference + flip-test + get_final_preds decode",
                       "crops_per_gpu_per_step": B, "forwards_per_crop": 2, "cuda_graph": pipe.graph is not None,
                       "l2": f"inputs ({B * 3 * IMAGE[0] * IMAGE[1] * 4 / 1e6:.0f} MB of crops per step) and activations exceed the 126 MB L2; no flush needed",
                       "partition": f"batch sharded over {world} GPU(s); only exchange = all-gather of keypoints "
                                    f"(204 B/crop, inside the timed step when n_gpus > 1)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "crops/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": pipe.launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel + basic_block_kernel (tcgen05 convolutions)", "achieved": achieved_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                         "traffic_note": NCU_TRAFFIC_NOTE,
                         "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)",
                         "launches_per_step": conv_n, "conv_ms_per_step": conv_ms, "other_kernels_ms": other_ms,
                         "whole_step_frac": (2 * FLOPS_PER_FORWARD * B / (ms_step * 1e-3) / 1e12) / peak_tf},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, dt, cores = cpu_port_crops_per_sec(args.cpu_sample)
            out["cpu_baseline"] = {
                "value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                "sample": f"{args.cpu_sample} crops, 1 pass ({dt:.1f} s): oracle port = torch-CPU fp32 HRNet-W32 x2 "
                          f"+ NumPy flip-average/decode"}
        print(json.dumps(out))
    if args.dump_ops and rank == 0:
        with open(args.dump_ops, "w") as f:
            json.dump(ops, f)
    if world > 1:
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="crops per GPU per step (BASELINE config 2: 512)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="crops in the bounded CPU-baseline sample")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--width", type=int, default=32, choices=[32, 48],
                    help="32: HRNet-W32 256x192 (the configuration the metric is quoted on); 48: HRNet-W48 384x288 (config 3)")
    ap.add_argument("--dump-ops", default="", help="write the per-launch timing table (JSON) to this path")
    args = ap.parse_args()
    if args.width == 48:
        global WIDTH, IMAGE, FLOPS_PER_FORWARD, NCU_TRAFFIC_BYTES_PER_LAUNCH, NCU_TRAFFIC_NOTE
        WIDTH, IMAGE, FLOPS_PER_FORWARD = 48, (384, 288), 70.6132e9
        NCU_TRAFFIC_BYTES_PER_LAUNCH, NCU_TRAFFIC_NOTE = None, "no ncu capture for the W48 shapes"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 and the OpenMP pool is sized when torch is first imported: give the
        # CPU arm every core this process may run on BEFORE that import
        try:
            avail = len(os.sched_getaffinity(0))
        except AttributeError:
            avail = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(avail)
        os.environ["MKL_NUM_THREADS"] = str(avail)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
