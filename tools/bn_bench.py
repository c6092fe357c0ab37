"""Per-kernel time of the train-mode BatchNorm forward / backward kernels at the layer shapes of HRNet-W32 (torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from stlpose_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = _lib.lib()
st = lambda: _lib.current_stream()
for (c, h, w) in ((32, 64, 48), (64, 32, 24), (128, 16, 12), (256, 8, 6), (256, 64, 48), (64, 64, 48)):
    z = torch.randn(B, h + 1, w + 1, c, device="cuda").bfloat16()
    y = torch.empty_like(z); dz = torch.empty_like(z); dres = torch.empty_like(z)
    gamma = torch.ones(c, device="cuda"); beta = torch.zeros(c, device="cuda")
    ws = torch.empty(L.stl_bn_workspace_floats(c), device="cuda"); mean = torch.empty(c, device="cuda"); rstd = torch.empty(c, device="cuda")
    out = torch.empty(2 * c, device="cuda"); tick = torch.zeros(8, dtype=torch.int32, device="cuda")
    def f():
        _lib.check(L.stl_bn_train_forward_ticket(_lib.ptr(z), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(dres), 1, 1e-5, 0.1, B, h, w, c,
                                                 _lib.ptr(y), _lib.ptr(ws), _lib.ptr(mean), _lib.ptr(rstd), None, None, tick.data_ptr(), st()))
        _lib.check(L.stl_bn_train_backward_ticket(_lib.ptr(z), _lib.ptr(y), _lib.ptr(z), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma), 1,
                                                  B, h, w, c, _lib.ptr(dz), _lib.ptr(dres), _lib.ptr(out), _lib.ptr(ws), tick.data_ptr() + 4, st()))
    def f_coop():
        _lib.check(L.stl_bn_train_forward_coop(_lib.ptr(z), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(dres), 1, 1e-5, 0.1, B, h, w, c,
                                               _lib.ptr(y), _lib.ptr(ws), _lib.ptr(mean), _lib.ptr(rstd), None, None, tick.data_ptr(),
                                               tick.data_ptr() + 8, st()))
        _lib.check(L.stl_bn_train_backward_coop(_lib.ptr(z), _lib.ptr(y), _lib.ptr(z), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
                                                _lib.ptr(beta), 1, B, h, w, c, _lib.ptr(dz), _lib.ptr(dres), _lib.ptr(out), _lib.ptr(ws),
                                                tick.data_ptr() + 4, tick.data_ptr() + 16, st()))
    for _ in range(3):
        f(); f_coop()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            f(); f_coop()
        torch.cuda.synchronize()
    row = {}
    for e in prof.key_averages():
        n = e.key
        k = "reduce0" if "channel_reduce_kernel<0>" in n else "reduce1" if "channel_reduce_kernel<1>" in n else \
            "apply" if "bn_apply" in n else "fwd_coop" if "bn_forward_coop" in n else "bwd_coop" if "bn_backward_coop" in n else \
            "backward" if "bn_backward" in n else None
        if k:
            row[k] = e.device_time_total / e.count
    mb = z.numel() * 2 / 1e6
    print(f"B={B} C={c:3d} {h}x{w}  tensor {mb:7.1f} MB  " + "  ".join(f"{k} {row.get(k, 0):6.1f} us" for k in ("reduce0", "apply", "fwd_coop", "reduce1", "backward", "bwd_coop")), flush=True)
