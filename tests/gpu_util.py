"""Helpers for the GPU parity tests: drive single fused convolutions through the C ABI."""
import ctypes

import torch
import torch.nn.functional as F

from stlpose_b200 import _lib

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def to_padded(x, c_pad=None):
    """fp32 NCHW cuda tensor -> padded-linear NHWC bf16 buffer (uint8 tensor)."""
    L = _lib.lib()
    N, C, H, W = x.shape
    c_pad = c_pad or C
    buf = torch.empty(L.stl_padded_bytes(N, c_pad, H, W), dtype=torch.uint8, device=x.device)
    _lib.check(L.stl_nchw_to_padded(_lib.ptr(x.contiguous()), _lib.ptr(buf), N, C, H, W, c_pad, _lib.current_stream()))
    return buf


def from_padded(buf, N, C, H, W, c_pad=None):
    L = _lib.lib()
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=buf.device)
    _lib.check(L.stl_padded_to_nchw(_lib.ptr(buf), _lib.ptr(out), N, C, H, W, c_pad or C, _lib.current_stream()))
    return out


def padded_border_is_zero(buf, N, C, H, W):
    t = buf.view(torch.bfloat16).view(N, H + 1, W + 1, C)
    return bool((t[:, H] == 0).all() and (t[:, :, W] == 0).all())


def pack(w, bn=None, bias=None, eps=1e-5):
    """OIHW fp32 (+BN dict) -> (packed bf16 weights buffer, fp32 bias, folded fp32 weights, folded bias)."""
    L = _lib.lib()
    cout, cin, k, _ = w.shape
    cout_pad = (cout + 15) // 16 * 16
    wp = torch.empty(k * k * cout_pad * cin * 2, dtype=torch.uint8, device=w.device)
    bp = torch.empty(cout_pad, dtype=torch.float32, device=w.device)
    g = b = m = v = None
    if bn is not None:
        g, b, m, v = (bn[n].contiguous() for n in ("weight", "bias", "running_mean", "running_var"))
    _lib.check(L.stl_pack_conv_weights(_lib.ptr(w.contiguous()), _lib.ptr(g), _lib.ptr(b), _lib.ptr(m), _lib.ptr(v),
                                       _lib.ptr(bias), eps if bn is not None else 0.0, cout, cin, k, cout_pad, cin,
                                       _lib.ptr(wp), _lib.ptr(bp), _lib.current_stream()))
    if bn is not None:
        scale = g / torch.sqrt(v + eps)
        wf = w * scale.view(-1, 1, 1, 1)
        bf = b - m * scale + (bias * scale if bias is not None else 0)
    else:
        wf = w
        bf = bias if bias is not None else torch.zeros(cout, device=w.device)
    return wp, bp, wf, bf, cout_pad


def run_conv(x, w, bn=None, bias=None, stride=1, relu=False, residual=None, ups=(), out_nchw=False, impl=0,
             force_mb=0, max_ctas=0):
    """x: fp32 NCHW (values should already be bf16-representable). Returns (y_ours fp32 NCHW, y_ref fp32 NCHW)."""
    L = _lib.lib()
    N, Cin, H, W = x.shape
    cout, _, k, _ = w.shape
    Ho, Wo = H // stride, W // stride
    wp, bp, wf, bfold, cout_pad = pack(w, bn, bias)
    xin = to_padded(x)
    d = _lib.ConvDesc()
    d.in_ = xin.data_ptr(); d.N, d.H, d.W, d.Cin = N, H, W, Cin
    if out_nchw:
        out = torch.full((N, cout, Ho, Wo), float("nan"), dtype=torch.float32, device=x.device)
    else:
        # zero cells as the engine's workspaces have them (stride-2 convs never touch them), every real cell
        # poisoned with a huge value so that an unwritten output is caught
        out = to_padded(torch.full((N, cout, Ho, Wo), 3.0e38, dtype=torch.float32, device=x.device))
    d.out = out.data_ptr(); d.Cout, d.Cout_pad = cout, cout_pad
    d.ksize, d.stride = k, stride
    d.w_packed = wp.data_ptr(); d.bias_packed = bp.data_ptr()
    keep = [xin, wp, bp]
    ref = F.conv2d(x, bf16_round(wf), None, stride, k // 2) + bfold.view(1, -1, 1, 1)
    if residual is not None:
        rb = to_padded(residual); keep.append(rb)
        d.residual = rb.data_ptr()
        ref = ref + residual
    d.n_up = len(ups)
    for i, (u, shift) in enumerate(ups):
        ub = to_padded(u); keep.append(ub)
        d.up_src[i] = ub.data_ptr(); d.up_shift[i] = shift
        ref = ref + F.interpolate(u, scale_factor=2 ** shift, mode="nearest")
    d.relu = int(relu); d.out_nchw = int(out_nchw); d.impl = impl; d.force_mb = force_mb; d.max_ctas = max_ctas
    if relu:
        ref = F.relu(ref)
    _lib.check(L.stl_conv2d(ctypes.byref(d), _lib.current_stream()))
    torch.cuda.synchronize()
    if out_nchw:
        return out, ref
    assert padded_border_is_zero(out, N, cout, Ho, Wo), "zero cells of the padded layout were overwritten"
    return from_padded(out, N, cout, Ho, Wo), ref


def rand_bn(c, gen, device):
    return {"weight": torch.empty(c, device=device).uniform_(0.5, 1.5, generator=gen),
            "bias": torch.empty(c, device=device).normal_(0, 0.1, generator=gen),
            "running_mean": torch.empty(c, device=device).normal_(0, 0.1, generator=gen),
            "running_var": torch.empty(c, device=device).uniform_(0.6, 1.4, generator=gen)}


def conv_case(N, cin, cout, H, W, k, stride, relu=True, with_res=False, n_up=0, out_nchw=False, with_bias=False,
              impl=0, seed=0, force_mb=0, max_ctas=0):
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(seed)
    x = bf16_round(torch.randn(N, cin, H, W, device=dev, generator=gen))
    w = torch.randn(cout, cin, k, k, device=dev, generator=gen) / (cin * k * k) ** 0.5
    bn = None if with_bias else rand_bn(cout, gen, dev)
    bias = torch.randn(cout, device=dev, generator=gen) * 0.1 if with_bias else None
    Ho, Wo = H // stride, W // stride
    res = bf16_round(torch.randn(N, cout, Ho, Wo, device=dev, generator=gen)) if with_res else None
    ups = [(bf16_round(torch.randn(N, cout, Ho >> (s + 1), Wo >> (s + 1), device=dev, generator=gen)), s + 1)
           for s in range(n_up)]
    y, ref = run_conv(x, w, bn, bias, stride, relu, res, ups, out_nchw, impl, force_mb, max_ctas)
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale
