"""Per-shape timing of the training kernels (wgrad, BN statistics) at a given batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stlpose_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = _lib.lib()
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
print(f"B={B}")
print("layer                    wgrad_us  TF/s   GB/s | bn_fwd_us bn_bwd_us GB/s(bwd 5 passes)")
for (cin, cout, k, h, w) in ((32, 32, 3, 64, 48), (64, 64, 3, 32, 24), (128, 128, 3, 16, 12), (256, 256, 3, 8, 6),
                             (64, 256, 1, 64, 48), (256, 64, 1, 64, 48), (64, 64, 3, 64, 48), (256, 32, 3, 64, 48),
                             (64, 32, 1, 32, 24), (256, 32, 1, 8, 6)):
    x = torch.randn(B, h + 1, w + 1, cin, device="cuda").bfloat16()
    dz = torch.randn(B, h + 1, w + 1, cout, device="cuda").bfloat16()
    dw = torch.empty(cout, cin, k, k, device="cuda")
    wsb = L.stl_conv_wgrad_workspace_bytes(B, h, w, cin, cout, k, 1, cin)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
    us = t(lambda: _lib.check(L.stl_conv_wgrad(_lib.ptr(x), _lib.ptr(dz), _lib.ptr(dw), B, h, w, cin, cout, k, 1, cin, _lib.ptr(ws), wsb, _lib.current_stream())))
    fl = 2.0 * B * h * w * cin * cout * k * k
    by = (x.numel() + dz.numel()) * 2
    y = torch.empty_like(dz); sums = torch.empty(L.stl_bn_workspace_floats(cout), device="cuda"); mean = torch.empty(cout, device="cuda"); rstd = torch.empty(cout, device="cuda")
    gamma = torch.ones(cout, device="cuda"); beta = torch.zeros(cout, device="cuda"); rm = torch.zeros(cout, device="cuda"); rv = torch.ones(cout, device="cuda")
    f_us = t(lambda: _lib.check(L.stl_bn_train_forward(_lib.ptr(dz), _lib.ptr(gamma), _lib.ptr(beta), None, 1, 1e-5, 0.1, B, h, w, cout,
                                                       _lib.ptr(y), _lib.ptr(sums), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(rm), _lib.ptr(rv), _lib.current_stream())))
    dzz = torch.empty_like(dz)
    b_us = t(lambda: _lib.check(L.stl_bn_train_backward(_lib.ptr(dz), _lib.ptr(y), _lib.ptr(dz), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma), 1,
                                                        B, h, w, cout, _lib.ptr(dzz), None, _lib.ptr(sums), _lib.current_stream())))
    print(f"{cin:3d}->{cout:3d} k{k} {h:3d}x{w:<3d}      {us:8.1f} {fl/us/1e6:6.1f} {by/us/1e3:6.0f} | {f_us:8.1f} {b_us:8.1f}  {dz.numel()*2*7/b_us/1e3:6.0f}")
