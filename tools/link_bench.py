"""Bottleneck junction (stl_bottleneck_link: conv3 + residual + ReLU of a layer1 Bottleneck and conv1 + ReLU of the next
one in one kernel) against the same two 1x1 convolutions as stl_conv2d launches, at benchmark scale.

    python tools/link_bench.py [--n 1024] [--iters 20]
Also times the two-input 1x1 convolution of layer1.0 (conv3 + downsample over K = [conv2 output | block input]) against
its two-launch form.
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


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    L = _lib.lib()
    dev = "cuda"
    n, h, w = a.n, 64, 48
    gen = torch.Generator(device=dev).manual_seed(0)
    st = _lib.current_stream()

    def padded(c):
        t = torch.zeros(L.stl_padded_bytes(n, c, h, w) // 2, dtype=torch.bfloat16, device=dev)
        t.view(n, h + 1, w + 1, c)[:, :h, :w].normal_(generator=gen)
        return t

    t64, x64, x256 = padded(64), padded(64), padded(256)
    out, a64, ds = torch.zeros_like(x256), torch.zeros_like(t64), torch.zeros_like(x256)
    w3 = torch.randn(256, 64, 1, 1, device=dev, generator=gen) / 8
    wd = torch.randn(256, 64, 1, 1, device=dev, generator=gen) / 8
    w1 = torch.randn(64, 256, 1, 1, device=dev, generator=gen) / 16
    wp3, bp3, _, _, _ = G.pack(w3, G.rand_bn(256, gen, dev))
    wpd, bpd, _, _, _ = G.pack(wd, G.rand_bn(256, gen, dev))
    wp1, bp1, _, _, _ = G.pack(w1, G.rand_bn(64, gen, dev))
    wcat = torch.cat([wp3.view(torch.bfloat16).view(256, 64), wpd.view(torch.bfloat16).view(256, 64)], 1).contiguous()
    bcat = (bp3 + bpd).contiguous()

    def conv(src, cin, dst, cout, wp, bp, residual=None, relu=1, src2=None, cin2=0):
        d = _lib.ConvDesc()
        d.in_ = src.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, cin
        d.out = dst.data_ptr(); d.Cout, d.Cout_pad = cout, cout
        d.ksize, d.stride = 1, 1
        d.w_packed = wp.data_ptr(); d.bias_packed = bp.data_ptr()
        d.residual = residual.data_ptr() if residual is not None else None
        d.relu = relu
        if src2 is not None:
            d.in2 = src2.data_ptr(); d.Cin2 = cin2
        _lib.check(L.stl_conv2d(ctypes.byref(d), st))

    px = n * h * w
    gb = lambda chans: px * chans * 2 / 1e9
    us = timed(lambda: (conv(t64, 64, out, 256, wp3, bp3, residual=x256), conv(out, 256, a64, 64, wp1, bp1)), a.iters)
    print(f"junction, two launches : {us:8.1f} us   {gb(64 + 256 + 256 + 256 + 64) / us * 1e3:6.2f} TB/s algorithmic")
    us = timed(lambda: _lib.check(L.stl_bottleneck_link(_lib.ptr(t64), _lib.ptr(x256), _lib.ptr(out), _lib.ptr(a64),
                                                        _lib.ptr(wp3), _lib.ptr(bp3), _lib.ptr(wp1), _lib.ptr(bp1), n, h, w,
                                                        0, st)), a.iters)
    print(f"junction, one kernel   : {us:8.1f} us   {gb(64 + 256 + 256 + 64) / us * 1e3:6.2f} TB/s algorithmic")
    us = timed(lambda: (conv(x64, 64, ds, 256, wpd, bpd, relu=0), conv(t64, 64, out, 256, wp3, bp3, residual=ds)), a.iters)
    print(f"conv3 + downsample, two launches : {us:8.1f} us   {gb(64 + 256 + 64 + 256 + 256) / us * 1e3:6.2f} TB/s algorithmic")
    us = timed(lambda: conv(t64, 64, out, 256, wcat, bcat, src2=x64, cin2=64), a.iters)
    print(f"conv3 + downsample, one launch   : {us:8.1f} us   {gb(64 + 64 + 256) / us * 1e3:6.2f} TB/s algorithmic")


if __name__ == "__main__":
    main()
