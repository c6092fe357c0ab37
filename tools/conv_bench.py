"""Time single fused-convolution launches at benchmark scale (BASELINE config 2: 1024 images = 512 crops x 2).

    python tools/conv_bench.py [case ...] [--iters 5] [--impl 0|1] [--n 1024]
Cases: c32 c64 c128 c256 l1c2 l1c3 t1 s2 head (see CASES). Prints us/launch, TFLOP/s and algorithmic GB/s.
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from stlpose_b200 import _lib  # noqa: E402
import gpu_util as G  # noqa: E402

CASES = {
    "c32": dict(cin=32, cout=32, H=64, W=48, k=3, stride=1, res=True),
    "c32nr": dict(cin=32, cout=32, H=64, W=48, k=3, stride=1, res=False),
    "c64": dict(cin=64, cout=64, H=32, W=24, k=3, stride=1, res=True),
    "c64nr": dict(cin=64, cout=64, H=32, W=24, k=3, stride=1, res=False),
    "c128": dict(cin=128, cout=128, H=16, W=12, k=3, stride=1, res=True),
    "c256": dict(cin=256, cout=256, H=8, W=6, k=3, stride=1, res=True),
    "w48c48": dict(cin=48, cout=48, H=96, W=72, k=3, stride=1, res=True),      # HRNet-W48 384x288: the high-resolution branch
    "w48c96": dict(cin=96, cout=96, H=48, W=36, k=3, stride=1, res=True),
    "c128nr": dict(cin=128, cout=128, H=16, W=12, k=3, stride=1, res=False),
    "c256nr": dict(cin=256, cout=256, H=8, W=6, k=3, stride=1, res=False),
    "l1c1": dict(cin=256, cout=64, H=64, W=48, k=1, stride=1, res=False),
    "l1c2": dict(cin=64, cout=64, H=64, W=48, k=3, stride=1, res=False),
    "l1c3": dict(cin=64, cout=256, H=64, W=48, k=1, stride=1, res=True),
    "l1c3nr": dict(cin=64, cout=256, H=64, W=48, k=1, stride=1, res=False),
    "t1": dict(cin=256, cout=32, H=64, W=48, k=3, stride=1, res=False),
    "s2": dict(cin=32, cout=64, H=64, W=48, k=3, stride=2, res=True),
    "stem2": dict(cin=64, cout=64, H=128, W=96, k=3, stride=2, res=False),
    "head": dict(cin=32, cout=17, H=64, W=48, k=1, stride=1, res=False, nchw=True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=["c32", "c64", "c128", "c256", "l1c2", "l1c3"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--mb", type=int, default=0)
    ap.add_argument("--max-ctas", type=int, default=0, help="restrict the launch to this many CTAs (SM partitioning experiments)")
    ap.add_argument("--counters", action="store_true", help="print per-role cycle counters of the last launch")
    args = ap.parse_args()
    L = _lib.lib()
    dev = "cuda"
    for name in args.cases:
        c = CASES[name]
        N, cin, cout, H, W, k, s = args.n, c["cin"], c["cout"], c["H"], c["W"], c["k"], c["stride"]
        Ho, Wo = H // s, W // s
        gen = torch.Generator(device=dev).manual_seed(0)
        w = torch.randn(cout, cin, k, k, device=dev, generator=gen) / (cin * k * k) ** 0.5
        nchw = c.get("nchw", False)
        wp, bp, _, _, cout_pad = G.pack(w, None if nchw else G.rand_bn(cout, gen, dev),
                                        torch.zeros(cout, device=dev) if nchw else None)
        nbytes_in = L.stl_padded_bytes(N, cin, H, W)
        xin = torch.zeros(nbytes_in // 2, dtype=torch.bfloat16, device=dev)
        xin.view(N, H + 1, W + 1, cin)[:, :H, :W].normal_(generator=gen)
        nbytes_out = L.stl_padded_bytes(N, cout, Ho, Wo)
        out = (torch.empty((N, cout, Ho, Wo), dtype=torch.float32, device=dev) if nchw
               else torch.zeros(nbytes_out, dtype=torch.uint8, device=dev))
        res = torch.zeros(nbytes_out // 2, dtype=torch.bfloat16, device=dev).normal_(generator=gen) if c["res"] else None
        d = _lib.ConvDesc()
        d.in_ = xin.data_ptr(); d.N, d.H, d.W, d.Cin = N, H, W, cin
        d.out = out.data_ptr(); d.Cout, d.Cout_pad = cout, cout_pad
        d.ksize, d.stride = k, s
        d.w_packed = wp.data_ptr(); d.bias_packed = bp.data_ptr()
        d.residual = res.data_ptr() if res is not None else None
        d.relu = 1; d.out_nchw = int(nchw); d.impl = args.impl; d.force_mb = args.mb; d.max_ctas = args.max_ctas
        st = _lib.current_stream()
        counters = torch.zeros((148, 3, 4), dtype=torch.int64, device=dev)
        if args.counters:
            d.dbg_counters = counters.data_ptr()
        _lib.check(L.stl_conv2d(ctypes.byref(d), st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            _lib.check(L.stl_conv2d(ctypes.byref(d), st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        flops = 2.0 * N * Ho * Wo * cout * cin * k * k
        byts = N * (H * W * cin * 2 + Ho * Wo * cout * (4 if nchw else 2) * (2 if c["res"] else 1))
        print(f"{name:6s} impl{args.impl} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.0f} GB/s",
              flush=True)
        if args.counters:
            c = counters.float().mean(dim=0).cpu() / 1e3
            print("        kcycles/CTA  producer: wait_a_empty %.0f wait_b_empty %.0f issue %.0f | mma: wait_acc %.0f "
                  "wait_a %.0f wait_b %.0f issue %.0f | epilogue(w2): bar1 %.0f compute %.0f fence+bar2 %.0f acc/res wait %.0f"
                  % (c[0, 0], c[0, 1], c[0, 2], c[1, 0], c[1, 1], c[1, 2], c[1, 3], c[2, 0], c[2, 1], c[2, 2], c[2, 3]),
                  flush=True)


if __name__ == "__main__":
    main()
