"""Measure the accumulation error of the tcgen05 conv (fp32 NCHW output path) against an fp64 convolution."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch, torch.nn.functional as F
from gpu_util import bf16_round, to_padded
from stlpose_b200 import _lib
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
L = _lib.lib()
for (cin, cout, k) in ((256, 32, 1), (64, 32, 3), (256, 32, 3), (32, 17, 1)):
    n, h, w = 2, 32, 24
    g = torch.Generator(device="cuda").manual_seed(1)
    x = bf16_round(torch.randn(n, cin, h, w, device="cuda", generator=g).relu())   # non-negative like post-ReLU
    wt = bf16_round(torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5)
    cp = (cout + 15) // 16 * 16
    wp = torch.empty(k * k * cp * cin * 2, dtype=torch.uint8, device="cuda"); bp = torch.zeros(cp, device="cuda")
    _lib.check(L.stl_pack_conv_weights(_lib.ptr(wt), None, None, None, None, None, 0.0, cout, cin, k, cp, cin, _lib.ptr(wp), _lib.ptr(bp), _lib.current_stream()))
    xp = to_padded(x)
    out = torch.empty(n, cout, h, w, device="cuda")
    d = _lib.ConvDesc(); d.in_ = xp.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, cin
    d.out = out.data_ptr(); d.Cout, d.Cout_pad = cout, cp; d.ksize, d.stride = k, 1
    d.w_packed = wp.data_ptr(); d.bias_packed = bp.data_ptr(); d.out_nchw = 1
    _lib.check(L.stl_conv2d(ctypes.byref(d), _lib.current_stream()))
    ref64 = F.conv2d(x.double(), wt.double(), None, 1, k // 2)
    ref32 = F.conv2d(x, wt, None, 1, k // 2)
    scale = ref64.abs().mean().item()
    e_dev = (out.double() - ref64); e_32 = (ref32.double() - ref64)
    flips_dev = (out.bfloat16() != ref64.float().bfloat16()).float().mean().item()
    flips_32 = (ref32.bfloat16() != ref64.float().bfloat16()).float().mean().item()
    print(f"cin {cin} k {k}: |out| mean {scale:.3f}  device err mean {e_dev.mean().item():+.2e} rms {e_dev.pow(2).mean().sqrt().item():.2e} max {e_dev.abs().max().item():.2e} | torch fp32 err rms {e_32.pow(2).mean().sqrt().item():.2e} | bf16 flips dev {flips_dev:.2e} torch32 {flips_32:.2e}")
