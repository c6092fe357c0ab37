"""decode / flip-average+decode / PersonMSELoss sweep, GB/s of algorithmic bytes vs HBM peak.

Algorithmic bytes per crop: decode = J*h*w*4 read + 340 written; fused flip-average + decode = 2x the read;
loss forward+gradient = 2*J*h*w*4 read + J*h*w*4 written (+ J*4 target weights).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stlpose_b200 as S
from stlpose_b200 import _lib, pose_parsing
from stlpose_b200.transforms import FLIP_PAIRS


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbps_burst", "hbm_gbps", "hbm_copy_gbps", "hbm_gbs"):
            if k in d:
                return float(d[k]), "MEASURED_PEAKS.json:" + k
        for k, v in d.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), "MEASURED_PEAKS.json:" + k
    except Exception:
        pass
    return 6533.0, "fallback 6533 GB/s (B200_PROFILING.md)"


def blobs(B, J, h, w, gen):
    cx = torch.rand(B, J, 1, 1, device="cuda", generator=gen) * (w - 1)
    cy = torch.rand(B, J, 1, 1, device="cuda", generator=gen) * (h - 1)
    ys = torch.arange(h, device="cuda").view(1, 1, h, 1); xs = torch.arange(w, device="cuda").view(1, 1, 1, w)
    out = torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 8.0)       # sigma = 2 (JointsDataset.py:248-281)
    out.add_(torch.randn(out.shape, device="cuda", generator=gen), alpha=0.01)
    return out


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-batch", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    peak, src = peaks()
    print(f"# HBM peak {peak:.0f} GB/s ({src}); inputs larger than the 126 MB L2 from B = 1 Ki (213 MB) on")
    print("size     dist   B      op              ms      Mcrops/s   GB/s   frac")
    gen = torch.Generator(device="cuda").manual_seed(0)
    J = 17
    rows = []
    for (h, w) in ((64, 48), (96, 72)):
        for B in (1024, 4096, 16384, 65536):
            if B > a.max_batch:
                continue
            for dist in ("randn", "blobs"):
                if dist == "randn":
                    heat = torch.randn(B, J, h, w, device="cuda", generator=gen)
                    heat2 = torch.randn(B, J, h, w, device="cuda", generator=gen)
                else:
                    chunks = [blobs(min(4096, B - i), J, h, w, gen) for i in range(0, B, 4096)]
                    heat = torch.cat(chunks); del chunks
                    heat2 = heat.flip(3).contiguous()
                c = torch.rand(B, 2, device="cuda", generator=gen) * 300 + 100
                s = torch.rand(B, 2, device="cuda", generator=gen) * 2 + 0.5
                per = J * h * w * 4
                ops = [("decode", lambda: pose_parsing._decode(heat, c, s, True), per + 340),
                       ("flipavg+decode", lambda: pose_parsing._decode(heat, c, s, True, heat_flipped=heat2, pairs=FLIP_PAIRS), 2 * per + 340)]
                if dist == "randn":
                    tw = torch.ones(B, J, 1, device="cuda")
                    crit = S.PersonMSELoss()
                    hg = heat.clone().requires_grad_(True)
                    def loss_step():
                        hg.grad = None
                        crit(hg, heat2, tw).backward()
                    ops.append(("loss fwd+grad", loss_step, 3 * per + J * 4))
                for name, fn, bytes_per in ops:
                    t = timeit(fn, a.iters)
                    gbs = bytes_per * B / t / 1e9
                    print(f"{h}x{w:<4d} {dist:6s} {B:<6d} {name:15s} {t*1e3:7.3f} {B/t/1e6:9.2f} {gbs:7.0f}  {gbs/peak:.3f}")
                    rows.append(dict(h=h, w=w, dist=dist, B=B, op=name, ms=t * 1e3, gbps=gbs, frac=gbs / peak))
                del heat, heat2
                torch.cuda.empty_cache()
    print("JSON " + json.dumps(rows))


if __name__ == "__main__":
    main()
