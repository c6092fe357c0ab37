"""Fused BasicBlock kernel (stl_basic_block) vs the same block as two fused convolutions (stl_conv2d) and vs torch."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16_round, from_padded, pack, padded_border_is_zero, to_padded
from stlpose_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bn(c, g):
    return dict(weight=torch.rand(c, device=DEV, generator=g) + 0.5, bias=torch.randn(c, device=DEV, generator=g) * 0.1,
                running_mean=torch.randn(c, device=DEV, generator=g) * 0.1,
                running_var=torch.rand(c, device=DEV, generator=g) + 0.5)


def _conv(L, xin, out, n, h, w, c, wp, bp, residual=None):
    d = _lib.ConvDesc()
    d.in_ = xin.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, c
    d.out = out.data_ptr(); d.Cout, d.Cout_pad = c, c
    d.ksize, d.stride = 3, 1
    d.w_packed = wp.data_ptr(); d.bias_packed = bp.data_ptr()
    d.residual = residual.data_ptr() if residual is not None else None
    d.relu = 1
    _lib.check(L.stl_conv2d(ctypes.byref(d), _lib.current_stream()))


@pytest.mark.parametrize("n,h,w", [(1, 64, 48), (3, 64, 48), (70, 64, 48), (5, 16, 12), (2, 8, 6), (130, 32, 24)])
def test_fused_basic_block_equals_two_convs(n, h, w):
    L = _lib.lib()
    c = 32
    g = torch.Generator(device=DEV).manual_seed(n * 100 + h)
    x = bf16_round(torch.randn(n, c, h, w, device=DEV, generator=g))
    w1 = torch.randn(c, c, 3, 3, device=DEV, generator=g) / (c * 9) ** 0.5
    w2 = torch.randn(c, c, 3, 3, device=DEV, generator=g) / (c * 9) ** 0.5
    wp1, bp1, wf1, bf1, _ = pack(w1, _bn(c, g))
    wp2, bp2, wf2, bf2, _ = pack(w2, _bn(c, g))
    xin = to_padded(x)
    mid = torch.empty_like(xin)
    y2 = torch.empty_like(xin)
    _conv(L, xin, mid, n, h, w, c, wp1, bp1)
    _conv(L, mid, y2, n, h, w, c, wp2, bp2, residual=xin)
    yf = torch.full_like(xin, 0x7f)                      # poison: the fused kernel must write every cell
    _lib.check(L.stl_basic_block(_lib.ptr(xin), _lib.ptr(yf), _lib.ptr(wp1), _lib.ptr(bp1), _lib.ptr(wp2), _lib.ptr(bp2),
                                 n, h, w, c, _lib.current_stream()))
    torch.cuda.synchronize()
    assert padded_border_is_zero(yf, n, c, h, w)
    assert torch.equal(yf, y2), (yf != y2).float().mean().item()
    # and against torch fp32 on the folded weights (bf16 operands, bf16 intermediate)
    t1 = bf16_round(F.relu(F.conv2d(x, bf16_round(wf1), bf1, 1, 1)))
    ref = F.relu(F.conv2d(t1, bf16_round(wf2), bf2, 1, 1) + x)
    got = from_padded(yf, n, c, h, w)
    assert (got - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("n,h,w,max_ctas", [(1, 64, 48, 0), (3, 64, 48, 0), (70, 64, 48, 0), (9, 64, 48, 3), (5, 16, 12, 2),
                                            (2, 8, 6, 0), (40, 32, 24, 1)])
def test_bottleneck_link_equals_two_convs(n, h, w, max_ctas):
    """stl_bottleneck_link (conv3 + residual + ReLU of a Bottleneck and conv1 + ReLU of the next one in one kernel,
    models/HRnet.py:88-101 / :82-84) vs the same two 1x1 convolutions as stl_conv2d launches: bit-identical `out` and `a`,
    zero cells kept, and vs torch fp32 on the folded weights.  max_ctas > 0: every CTA walks many tiles (all rings wrap)."""
    L = _lib.lib()
    ct, co, ca = 64, 256, 64
    g = torch.Generator(device=DEV).manual_seed(n * 10 + h)
    t = bf16_round(torch.randn(n, ct, h, w, device=DEV, generator=g))
    x = bf16_round(torch.randn(n, co, h, w, device=DEV, generator=g))
    w3 = torch.randn(co, ct, 1, 1, device=DEV, generator=g) / ct ** 0.5
    w1 = torch.randn(ca, co, 1, 1, device=DEV, generator=g) / co ** 0.5
    wp3, bp3, wf3, bf3, _ = pack(w3, _bn(co, g))
    wp1, bp1, wf1, bf1, _ = pack(w1, _bn(ca, g))
    tin, xin = to_padded(t), to_padded(x)

    def conv1x1(src, cin, dst, cout, wp, bp, residual=None):
        d = _lib.ConvDesc()
        d.in_ = src.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, cin
        d.out = dst.data_ptr(); d.Cout, d.Cout_pad = cout, cout
        d.ksize, d.stride = 1, 1
        d.w_packed = wp.data_ptr(); d.bias_packed = bp.data_ptr()
        d.residual = residual.data_ptr() if residual is not None else None
        d.relu = 1
        _lib.check(L.stl_conv2d(ctypes.byref(d), _lib.current_stream()))

    out2 = torch.zeros_like(xin)
    a2 = torch.zeros_like(tin)
    conv1x1(tin, ct, out2, co, wp3, bp3, residual=xin)
    conv1x1(out2, co, a2, ca, wp1, bp1)
    out1 = torch.full_like(xin, 0x7f)                      # poison: the fused kernel must write every cell
    a1 = torch.full_like(tin, 0x7f)
    _lib.check(L.stl_bottleneck_link(_lib.ptr(tin), _lib.ptr(xin), _lib.ptr(out1), _lib.ptr(a1), _lib.ptr(wp3), _lib.ptr(bp3),
                                     _lib.ptr(wp1), _lib.ptr(bp1), n, h, w, max_ctas, _lib.current_stream()))
    torch.cuda.synchronize()
    assert padded_border_is_zero(out1, n, co, h, w) and padded_border_is_zero(a1, n, ca, h, w)
    assert torch.equal(out1, out2), (out1 != out2).float().mean().item()
    assert torch.equal(a1, a2), (a1 != a2).float().mean().item()
    o_ref = F.relu(F.conv2d(t, bf16_round(wf3), bf3) + x)
    a_ref = F.relu(F.conv2d(bf16_round(o_ref), bf16_round(wf1), bf1))
    assert (from_padded(out1, n, co, h, w) - o_ref).abs().max().item() < 1e-2 * max(1.0, o_ref.abs().max().item())
    assert (from_padded(a1, n, ca, h, w) - a_ref).abs().max().item() < 2e-2 * max(1.0, a_ref.abs().max().item())
    # the residual must not alias the output
    assert L.stl_bottleneck_link(_lib.ptr(tin), _lib.ptr(out1), _lib.ptr(out1), _lib.ptr(a1), _lib.ptr(wp3), _lib.ptr(bp3),
                                 _lib.ptr(wp1), _lib.ptr(bp1), n, h, w, 0, _lib.current_stream()) != 0


@pytest.mark.parametrize("n,h,w,max_ctas", [(1, 64, 48, 0), (70, 64, 48, 0), (9, 64, 48, 3), (5, 16, 12, 2), (2, 8, 6, 0),
                                            (40, 32, 24, 1)])
def test_bottleneck_link_two_inputs_equals_two_convs(n, h, w, max_ctas):
    """stl_bottleneck_link2 (layer1.0 -> layer1.1: conv3 + downsample over K = [t | t2] with concatenated weights, ReLU,
    then conv1 + ReLU of the next block, one kernel) vs the two-input stl_conv2d followed by the 256 -> 64 stl_conv2d:
    bit-identical, zero cells kept; and vs torch fp32 as two separate convolutions summed (models/HRnet.py:88-101)."""
    L = _lib.lib()
    ct, co, ca = 64, 256, 64
    g = torch.Generator(device=DEV).manual_seed(n * 10 + h + 1)
    t = bf16_round(torch.randn(n, ct, h, w, device=DEV, generator=g))
    t2 = bf16_round(torch.randn(n, ct, h, w, device=DEV, generator=g))
    w3 = torch.randn(co, ct, 1, 1, device=DEV, generator=g) / ct ** 0.5
    wd = torch.randn(co, ct, 1, 1, device=DEV, generator=g) / ct ** 0.5
    w1 = torch.randn(ca, co, 1, 1, device=DEV, generator=g) / co ** 0.5
    wp3, bp3, wf3, bf3, _ = pack(w3, _bn(co, g))
    wpd, bpd, wfd, bfd, _ = pack(wd, _bn(co, g))
    wp1, bp1, wf1, bf1, _ = pack(w1, _bn(ca, g))
    wcat = torch.cat([wp3.view(torch.bfloat16).view(co, ct), wpd.view(torch.bfloat16).view(co, ct)], 1).contiguous()
    bcat = (bp3 + bpd).contiguous()
    tin, t2in = to_padded(t), to_padded(t2)
    out2 = to_padded(torch.zeros(n, co, h, w, device=DEV))
    a2 = torch.zeros_like(tin)
    d = _lib.ConvDesc()
    d.in_ = tin.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, ct
    d.in2 = t2in.data_ptr(); d.Cin2 = ct
    d.out = out2.data_ptr(); d.Cout, d.Cout_pad = co, co
    d.ksize, d.stride = 1, 1
    d.w_packed = wcat.data_ptr(); d.bias_packed = bcat.data_ptr()
    d.relu = 1
    _lib.check(L.stl_conv2d(ctypes.byref(d), _lib.current_stream()))
    d2 = _lib.ConvDesc()
    d2.in_ = out2.data_ptr(); d2.N, d2.H, d2.W, d2.Cin = n, h, w, co
    d2.out = a2.data_ptr(); d2.Cout, d2.Cout_pad = ca, ca
    d2.ksize, d2.stride = 1, 1
    d2.w_packed = wp1.data_ptr(); d2.bias_packed = bp1.data_ptr()
    d2.relu = 1
    _lib.check(L.stl_conv2d(ctypes.byref(d2), _lib.current_stream()))
    out1 = torch.full_like(out2, 0x7f)
    a1 = torch.full_like(tin, 0x7f)
    _lib.check(L.stl_bottleneck_link2(_lib.ptr(tin), _lib.ptr(t2in), _lib.ptr(out1), _lib.ptr(a1), _lib.ptr(wcat),
                                      _lib.ptr(bcat), _lib.ptr(wp1), _lib.ptr(bp1), n, h, w, max_ctas, _lib.current_stream()))
    torch.cuda.synchronize()
    assert padded_border_is_zero(out1, n, co, h, w) and padded_border_is_zero(a1, n, ca, h, w)
    assert torch.equal(out1, out2), (out1 != out2).float().mean().item()
    assert torch.equal(a1, a2), (a1 != a2).float().mean().item()
    o_ref = F.relu(F.conv2d(t, bf16_round(wf3), bf3) + F.conv2d(t2, bf16_round(wfd), bfd))
    a_ref = F.relu(F.conv2d(bf16_round(o_ref), bf16_round(wf1), bf1))
    assert (from_padded(out1, n, co, h, w) - o_ref).abs().max().item() < 1e-2 * max(1.0, o_ref.abs().max().item())
    assert (from_padded(a1, n, ca, h, w) - a_ref).abs().max().item() < 2e-2 * max(1.0, a_ref.abs().max().item())
