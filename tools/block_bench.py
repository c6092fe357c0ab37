"""Fused BasicBlock kernel vs the same block as two conv launches, at benchmark scale (1 024 images)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from stlpose_b200 import _lib

L = _lib.lib()
n, h, w, c = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 64, 48, 32
x = (torch.randn(n, h + 1, w + 1, c, device="cuda") * 0.5).bfloat16()
x[:, h] = 0; x[:, :, w] = 0
def packed():
    wt = torch.randn(c, c, 3, 3, device="cuda") / (c * 9) ** 0.5
    wp = torch.empty(9 * c * c * 2, dtype=torch.uint8, device="cuda"); bp = torch.empty(c, device="cuda")
    _lib.check(L.stl_pack_conv_weights(_lib.ptr(wt), None, None, None, None, None, 0.0, c, c, 3, c, c, _lib.ptr(wp), _lib.ptr(bp), _lib.current_stream()))
    return wp, bp
wp1, bp1 = packed(); wp2, bp2 = packed()
mid = torch.empty_like(x); y = torch.empty_like(x); yf = torch.empty_like(x)
def conv(xin, out, wp, bp, res):
    d = _lib.ConvDesc(); d.in_ = xin.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, c
    d.out = out.data_ptr(); d.Cout, d.Cout_pad = c, c; d.ksize, d.stride = 3, 1
    d.w_packed = wp.data_ptr(); d.bias_packed = bp.data_ptr(); d.residual = res.data_ptr() if res is not None else None; d.relu = 1
    _lib.check(L.stl_conv2d(ctypes.byref(d), _lib.current_stream()))
def two():
    conv(x, mid, wp1, bp1, None); conv(mid, y, wp2, bp2, x)
def fused():
    _lib.check(L.stl_basic_block(_lib.ptr(x), _lib.ptr(yf), _lib.ptr(wp1), _lib.ptr(bp1), _lib.ptr(wp2), _lib.ptr(bp2), n, h, w, c, _lib.current_stream()))
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
a, b = t(two), t(fused)
fl = 2 * 2.0 * n * h * w * c * c * 9
print(f"images {n}: two convs {a:.1f} us ({fl/a/1e6:.0f} TF/s)   fused block {b:.1f} us ({fl/b/1e6:.0f} TF/s)   equal {torch.equal(y, yf)}")
