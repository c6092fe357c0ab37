"""GPU parity: the tcgen05 implicit-GEMM convolution vs torch fp32 conv2d on the same bf16-rounded operands.

Tolerance: both sides multiply identical bf16 values and accumulate in fp32, so they differ only by summation
order and the final bf16 rounding of our output (2^-9 relative): |diff| <= 1e-2 * max|ref|.
"""
import pytest

from gpu_util import conv_case

pytestmark = pytest.mark.gpu

SHAPES = [
    # every conv family of SURVEY.md appendix A (W32) at small batch, plus W48 channel counts
    dict(N=2, cin=32, cout=32, H=64, W=48, k=3, stride=1),
    dict(N=3, cin=64, cout=64, H=32, W=24, k=3, stride=1, with_res=True),
    dict(N=5, cin=128, cout=128, H=16, W=12, k=3, stride=1, with_res=True),
    dict(N=9, cin=256, cout=256, H=8, W=6, k=3, stride=1),
    dict(N=2, cin=64, cout=64, H=64, W=48, k=3, stride=1),
    dict(N=2, cin=64, cout=256, H=64, W=48, k=1, stride=1, with_res=True),
    dict(N=1, cin=256, cout=32, H=64, W=48, k=3, stride=1),
    dict(N=2, cin=256, cout=64, H=64, W=48, k=1, stride=1),
    dict(N=2, cin=64, cout=64, H=128, W=96, k=3, stride=2),
    dict(N=1, cin=256, cout=64, H=64, W=48, k=3, stride=2),
    dict(N=3, cin=32, cout=64, H=64, W=48, k=3, stride=2, with_res=True, n_up=2),
    dict(N=3, cin=64, cout=128, H=32, W=24, k=3, stride=2, with_res=True, n_up=1),
    dict(N=3, cin=32, cout=32, H=64, W=48, k=3, stride=2),
    dict(N=5, cin=128, cout=256, H=16, W=12, k=3, stride=2, with_res=True),
    dict(N=3, cin=64, cout=32, H=32, W=24, k=1, stride=1, relu=False),
    dict(N=3, cin=256, cout=32, H=8, W=6, k=1, stride=1, relu=False),
    dict(N=3, cin=32, cout=17, H=64, W=48, k=1, stride=1, relu=False, out_nchw=True, with_bias=True),
    dict(N=4, cin=48, cout=48, H=24, W=18, k=3, stride=1),
    dict(N=2, cin=96, cout=96, H=48, W=36, k=3, stride=1),
    dict(N=3, cin=192, cout=192, H=24, W=18, k=3, stride=1),
    dict(N=3, cin=384, cout=384, H=12, W=9, k=3, stride=1),
    dict(N=2, cin=48, cout=96, H=96, W=72, k=3, stride=2),
]


@pytest.mark.parametrize("case", SHAPES, ids=lambda c: "c{cin}-{cout}_k{k}s{stride}_{H}x{W}_N{N}".format(**c))
@pytest.mark.parametrize("impl", [0, 1])
def test_conv_matches_torch(case, impl):
    err, scale = conv_case(impl=impl, **case)
    assert err <= 1e-2 * max(scale, 1.0), f"max err {err} vs ref max {scale}"


def test_reference_kernel_matches_torch():
    err, scale = conv_case(impl=2, N=2, cin=32, cout=32, H=16, W=12, k=3, stride=1, with_res=True)
    assert err <= 1e-2 * max(scale, 1.0)


@pytest.mark.parametrize("mb", [1, 2, 3])
def test_conv_tile_sizes_and_persistence(mb):
    # few CTAs -> every CTA walks many tiles: exercises ring wrap-around and accumulator double buffering
    err, scale = conv_case(N=6, cin=64, cout=64, H=32, W=24, k=3, stride=1, with_res=True, force_mb=mb, max_ctas=3)
    assert err <= 1e-2 * max(scale, 1.0)


KW_MERGED = [
    # layers the kw-merged MMA shape (impl 3: N = 3 x Cout, row-shifted sum in the epilogue) covers; many tiles per CTA
    # exercise the cross-warp row exchange, accumulator double buffering and the 126-row panels at tensor / image borders
    dict(N=3, cin=64, cout=64, H=32, W=24, k=3, stride=1, with_res=True),
    dict(N=2, cin=64, cout=64, H=64, W=48, k=3, stride=1),
    dict(N=2, cin=32, cout=32, H=64, W=48, k=3, stride=1, with_res=True),
    dict(N=7, cin=64, cout=64, H=32, W=24, k=3, stride=1, with_res=True, max_ctas=3),
    dict(N=1, cin=64, cout=64, H=8, W=6, k=3, stride=1, relu=False),
    dict(N=5, cin=128, cout=128, H=16, W=12, k=3, stride=1, with_res=True),   # not covered: must fall back to impl 0
]


@pytest.mark.parametrize("case", KW_MERGED, ids=lambda c: "c{cin}-{cout}_{H}x{W}_N{N}".format(**c))
def test_kw_merged_conv_matches_torch(case):
    err, scale = conv_case(impl=3, **case)
    assert err <= 1e-2 * max(scale, 1.0), f"max err {err} vs ref max {scale}"


@pytest.mark.parametrize("shape", [(2, 64, 64, 256, 64, 48), (3, 64, 64, 256, 16, 12), (2, 32, 64, 64, 32, 24),
                                   (5, 128, 64, 128, 8, 6), (2, 256, 64, 256, 16, 12)],
                         ids=lambda s: "c{1}+{2}-{3}_{4}x{5}_N{0}".format(*s))
@pytest.mark.parametrize("max_ctas", [0, 3])
def test_two_input_1x1_conv_equals_the_sum_of_two_convs(shape, max_ctas):
    """stl_conv_desc.in2: one 1x1 convolution over K = [in | in2] = relu(bn_a(conv_a(in)) + bn_b(conv_b(in2))), the
    Bottleneck tail of layer1.0 (conv3 + downsample, models/HRnet.py:88-101), against torch fp32 on the same
    bf16-rounded operands.  max_ctas = 3: every CTA walks many tiles (ring wrap-around with alternating tensor maps)."""
    import ctypes
    import torch
    import torch.nn.functional as F
    from stlpose_b200 import _lib
    from gpu_util import bf16_round, to_padded, from_padded, pack, rand_bn, padded_border_is_zero
    n, ca, cb, cout, h, w = shape
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(ca + 7 * cb + h)
    xa = bf16_round(torch.randn(n, ca, h, w, device=dev, generator=gen))
    xb = bf16_round(torch.randn(n, cb, h, w, device=dev, generator=gen))
    wa = torch.randn(cout, ca, 1, 1, device=dev, generator=gen) / ca ** 0.5
    wb = torch.randn(cout, cb, 1, 1, device=dev, generator=gen) / cb ** 0.5
    wpa, bpa, wfa, bfa, cout_pad = pack(wa, rand_bn(cout, gen, dev))
    wpb, bpb, wfb, bfb, _ = pack(wb, rand_bn(cout, gen, dev))
    # concatenated packed parameters, as Plan::pack_conv builds them: [cout_pad][ca | cb] bf16, summed bias
    wcat = torch.cat([wpa.view(torch.bfloat16).view(cout_pad, ca), wpb.view(torch.bfloat16).view(cout_pad, cb)], 1).contiguous()
    bcat = (bpa + bpb).contiguous()
    ina, inb = to_padded(xa), to_padded(xb)
    out = to_padded(torch.full((n, cout, h, w), 3.0e38, dtype=torch.float32, device=dev))
    d = _lib.ConvDesc()
    d.in_ = ina.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, ca
    d.in2 = inb.data_ptr(); d.Cin2 = cb
    d.out = out.data_ptr(); d.Cout, d.Cout_pad = cout, cout_pad
    d.ksize, d.stride = 1, 1
    d.w_packed = wcat.data_ptr(); d.bias_packed = bcat.data_ptr()
    d.relu = 1; d.max_ctas = max_ctas
    _lib.check(_lib.lib().stl_conv2d(ctypes.byref(d), _lib.current_stream()))
    torch.cuda.synchronize()
    assert padded_border_is_zero(out, n, cout, h, w)
    y = from_padded(out, n, cout, h, w)
    ref = F.relu(F.conv2d(xa, bf16_round(wfa)) + F.conv2d(xb, bf16_round(wfb)) + (bfa + bfb).view(1, -1, 1, 1))
    err, scale = (y - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 1e-2 * max(scale, 1.0), f"max err {err} vs ref max {scale}"
    # a second input on anything but a 1x1 / stride-1 convolution is refused
    d.ksize = 3
    assert _lib.lib().stl_conv2d(ctypes.byref(d), _lib.current_stream()) != 0
