"""GPU parity of the individual training kernels (through the C ABI) vs torch fp32 autograd on identical
bf16-rounded operands."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16_round, from_padded, to_padded
from stlpose_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _padded(t):                      # fp32 NCHW -> padded bf16 tensor view [N,H+1,W+1,C]
    n, c, h, w = t.shape
    return to_padded(t).view(torch.bfloat16).view(n, h + 1, w + 1, c)


def _unpadded(p):                    # padded bf16 [N,H+1,W+1,C] -> fp32 NCHW
    n, hp, wp, c = p.shape
    return from_padded(p.contiguous().view(torch.uint8).view(-1), n, c, hp - 1, wp - 1)


@pytest.mark.parametrize("shape", [(3, 32, 16, 12), (2, 64, 8, 6), (5, 48, 12, 9), (2, 256, 8, 6)])
@pytest.mark.parametrize("relu,with_res", [(True, True), (True, False), (False, False)])
def test_bn_train_forward_backward(shape, relu, with_res):
    L = _lib.lib()
    n, c, h, w = shape
    g = torch.Generator(device=DEV).manual_seed(c + h)
    z = bf16_round(torch.randn(shape, device=DEV, generator=g) * 1.5 + 0.3)
    res = bf16_round(torch.randn(shape, device=DEV, generator=g)) if with_res else None
    gamma = torch.rand(c, device=DEV, generator=g) + 0.5
    beta = torch.randn(c, device=DEV, generator=g) * 0.1
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    zp = _padded(z)
    y = torch.empty_like(zp)
    sums = torch.empty(L.stl_bn_workspace_floats(c), device=DEV); mean = torch.empty(c, device=DEV); rstd = torch.empty(c, device=DEV)
    resp = _padded(res) if with_res else None
    _lib.check(L.stl_bn_train_forward(_lib.ptr(zp), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(resp), int(relu), 1e-5, 0.1,
                                      n, h, w, c, _lib.ptr(y), _lib.ptr(sums), _lib.ptr(mean), _lib.ptr(rstd),
                                      _lib.ptr(rm), _lib.ptr(rv), _lib.current_stream()))
    # reference
    zr = z.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if with_res else None
    rm_ref, rv_ref = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    yr = F.batch_norm(zr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)
    if with_res:
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    got = _unpadded(y)
    assert (got - yr.detach()).abs().max().item() < 2e-2 * max(1.0, yr.abs().max().item())
    assert (rm - rm_ref).abs().max().item() < 1e-4 and (rv - rv_ref).abs().max().item() < 1e-3
    assert (y[:, h] == 0).all() and (y[:, :, w] == 0).all()
    # backward
    dy = bf16_round(torch.randn(shape, device=DEV, generator=g))
    yr.backward(dy)
    dyp = _padded(dy)
    dz = torch.empty_like(zp); dres = torch.empty_like(zp) if with_res else None
    sums2 = torch.empty(L.stl_bn_workspace_floats(c), device=DEV)
    _lib.check(L.stl_bn_train_backward(_lib.ptr(dyp), _lib.ptr(y), _lib.ptr(zp), _lib.ptr(mean), _lib.ptr(rstd),
                                       _lib.ptr(gamma), int(relu), n, h, w, c, _lib.ptr(dz), _lib.ptr(dres),
                                       _lib.ptr(sums2), _lib.current_stream()))
    scale = max(1.0, zr.grad.abs().max().item())
    assert (_unpadded(dz) - zr.grad).abs().max().item() < 3e-2 * scale
    assert (sums2[:c] - br.grad).abs().max().item() < 2e-2 * max(1.0, br.grad.abs().max().item())
    assert (sums2[c:2 * c] - gr.grad).abs().max().item() < 2e-2 * max(1.0, gr.grad.abs().max().item())
    if with_res:
        assert (_unpadded(dres) - rr.grad).abs().max().item() < 1e-2


@pytest.mark.parametrize("case", [
    dict(n=2, cin=32, cout=32, h=16, w=12, k=3, s=1), dict(n=3, cin=64, cout=32, h=16, w=12, k=1, s=1),
    dict(n=2, cin=32, cout=64, h=16, w=12, k=3, s=2), dict(n=2, cin=16, cout=64, h=32, w=24, k=3, s=2, cin_real=3),
    dict(n=2, cin=128, cout=256, h=16, w=12, k=3, s=2), dict(n=3, cin=32, cout=32, h=64, w=48, k=1, s=1, cout_real=17),
])
def test_conv_gradients(case):
    L = _lib.lib()
    n, cin, cout, h, w, k, s = (case[x] for x in ("n", "cin", "cout", "h", "w", "k", "s"))
    cin_real = case.get("cin_real", cin)
    g = torch.Generator(device=DEV).manual_seed(cin * cout + k)
    x = bf16_round(torch.randn(n, cin, h, w, device=DEV, generator=g))
    x[:, cin_real:] = 0
    wt = bf16_round(torch.randn(cout, cin_real, k, k, device=DEV, generator=g) / (cin_real * k * k) ** 0.5)
    ho, wo = h // s, w // s
    dz = bf16_round(torch.randn(n, cout, ho, wo, device=DEV, generator=g))
    if "cout_real" in case:
        dz[:, case["cout_real"]:] = 0
    # reference via autograd
    xr, wr = x[:, :cin_real].clone().requires_grad_(True), wt.clone().requires_grad_(True)
    F.conv2d(xr, wr, None, s, k // 2).backward(dz)
    # ours
    wp = torch.empty(k * k * cout * cin * 2, dtype=torch.uint8, device=DEV)
    bp = torch.empty(cout, device=DEV)
    _lib.check(L.stl_pack_conv_weights(_lib.ptr(wt.contiguous()), None, None, None, None, None, 0.0, cout, cin_real, k,
                                       cout, cin, _lib.ptr(wp), _lib.ptr(bp), _lib.current_stream()))
    xp, dzp = _padded(x), _padded(dz)
    dx = torch.empty_like(xp)
    _lib.check(L.stl_conv_dgrad(_lib.ptr(dzp), _lib.ptr(wp), _lib.ptr(dx), n, h, w, cin, cout, k, s, _lib.current_stream()))
    dw = torch.empty((cout, cin_real, k, k), device=DEV)
    wsb = L.stl_conv_wgrad_workspace_bytes(n, h, w, cin, cout, k, s, cin_real)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
    _lib.check(L.stl_conv_wgrad(_lib.ptr(xp), _lib.ptr(dzp), _lib.ptr(dw), n, h, w, cin, cout, k, s, cin_real,
                                _lib.ptr(ws), wsb, _lib.current_stream()))
    got_dx = _unpadded(dx)[:, :cin_real]
    assert (got_dx - xr.grad).abs().max().item() < 2e-2 * max(1.0, xr.grad.abs().max().item())
    assert (dx[:, h] == 0).all() and (dx[:, :, w] == 0).all()
    assert (dw - wr.grad).abs().max().item() < 1e-3 * max(1.0, wr.grad.abs().max().item())


def test_sum_relu_and_backward():
    L = _lib.lib()
    n, c, h, w = 3, 32, 16, 12
    g = torch.Generator(device=DEV).manual_seed(1)
    a = bf16_round(torch.randn(n, c, h, w, device=DEV, generator=g)).requires_grad_(True)
    b = bf16_round(torch.randn(n, c, h, w, device=DEV, generator=g)).requires_grad_(True)
    u1 = bf16_round(torch.randn(n, c, h // 2, w // 2, device=DEV, generator=g)).requires_grad_(True)
    u2 = bf16_round(torch.randn(n, c, h // 4, w // 4, device=DEV, generator=g)).requires_grad_(True)
    ref = F.relu(a + b + F.interpolate(u1, scale_factor=2, mode="nearest") + F.interpolate(u2, scale_factor=4, mode="nearest"))
    ap, bp_, u1p, u2p = _padded(a.detach()), _padded(b.detach()), _padded(u1.detach()), _padded(u2.detach())
    y = torch.empty_like(ap)
    same = (ctypes.c_void_p * 4)(ap.data_ptr(), bp_.data_ptr())
    ups = (ctypes.c_void_p * 3)(u1p.data_ptr(), u2p.data_ptr())
    sh = (ctypes.c_int * 3)(1, 2)
    _lib.check(L.stl_sum_relu_forward(same, 2, ups, sh, 2, _lib.ptr(y), n, h, w, c, _lib.current_stream()))
    assert (_unpadded(y) - ref.detach()).abs().max().item() < 2e-2 * ref.abs().max().item()
    dy = bf16_round(torch.randn(n, c, h, w, device=DEV, generator=g))
    ref.backward(dy)
    gm = torch.empty_like(ap)
    _lib.check(L.stl_relu_mask(_lib.ptr(_padded(dy)), _lib.ptr(y), _lib.ptr(gm), y.numel(), _lib.current_stream()))
    mask_ok = (ref.detach() > 1e-2) | (ref.detach() == 0)       # away from the bf16 rounding of the zero crossing
    assert ((_unpadded(gm) - a.grad).abs() * mask_ok).max().item() < 1e-6
    for up, shift in ((u1, 1), (u2, 2)):
        dlow = torch.empty((n, (h >> shift) + 1, (w >> shift) + 1, c), dtype=torch.bfloat16, device=DEV)
        _lib.check(L.stl_upsample_backward(_lib.ptr(gm), _lib.ptr(dlow), n, h, w, c, shift, _lib.current_stream()))
        want = F.avg_pool2d(_unpadded(gm), 1 << shift) * (1 << shift) ** 2
        assert (_unpadded(dlow) - want).abs().max().item() < 3e-2 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("cin,cout,k,hw", [(32, 32, 3, (64, 48)), (64, 64, 3, (32, 24)), (128, 128, 3, (16, 12)),
                                           (256, 256, 3, (8, 6)), (256, 64, 1, (64, 48)), (64, 256, 1, (64, 48)),
                                           (256, 32, 3, (64, 48)), (128, 32, 1, (16, 12)), (256, 128, 1, (8, 6))])
def test_stride1_dgrad_on_tensor_cores(cin, cout, k, hw):
    """training.conv_dgrad (stride 1 = tcgen05 convolution with the flipped, transposed filter) vs autograd."""
    from stlpose_b200 import training
    n, (h, w) = 3, hw
    g = torch.Generator(device=DEV).manual_seed(cin + cout + k)
    wt = bf16_round(torch.randn(cout, cin, k, k, device=DEV, generator=g) / (cin * k * k) ** 0.5)
    dz = bf16_round(torch.randn(n, cout, h, w, device=DEV, generator=g))
    x = torch.zeros(n, cin, h, w, device=DEV, requires_grad=True)
    F.conv2d(x, wt, None, 1, k // 2).backward(dz)
    dx = training.conv_dgrad(_padded(dz), wt, n, h, w, cin, 1)
    assert (_unpadded(dx) - x.grad).abs().max().item() < 2e-2 * max(1.0, x.grad.abs().max().item())
    assert (dx[:, h] == 0).all() and (dx[:, :, w] == 0).all()


@pytest.mark.parametrize("cin,cout,k,hw,n", [
    (32, 32, 3, (64, 48), 3), (64, 64, 3, (32, 24), 3), (128, 128, 3, (16, 12), 5), (256, 256, 3, (8, 6), 7),
    (64, 256, 1, (64, 48), 2), (256, 64, 1, (64, 48), 2), (256, 32, 3, (64, 48), 2), (128, 32, 1, (16, 12), 3),
    (256, 128, 1, (8, 6), 3), (32, 32, 1, (64, 48), 2), (64, 64, 1, (64, 48), 1), (32, 32, 3, (16, 12), 1),
    # HRNet-W48 channel counts (48 / 96 / 192 / 384): the next larger MMA shape, surplus channels zero-filled by TMA
    (48, 48, 3, (96, 72), 2), (96, 96, 3, (48, 36), 2), (192, 192, 3, (24, 18), 3), (384, 384, 3, (12, 9), 3),
    (256, 48, 3, (96, 72), 1), (96, 48, 1, (48, 36), 2), (384, 96, 1, (12, 9), 3), (192, 48, 1, (24, 18), 2),
    (48, 32, 1, (96, 72), 1), (384, 192, 1, (12, 9), 2)])
def test_stride1_wgrad_on_tensor_cores(cin, cout, k, hw, n):
    """stl_conv_wgrad (tcgen05, pixels as the reduction dimension) vs autograd and vs the CUDA-core kernel."""
    L = _lib.lib()
    h, w = hw
    g = torch.Generator(device=DEV).manual_seed(cin * 3 + cout + k)
    x = bf16_round(torch.randn(n, cin, h, w, device=DEV, generator=g))
    dz = bf16_round(torch.randn(n, cout, h, w, device=DEV, generator=g))
    wr = torch.zeros(cout, cin, k, k, device=DEV, requires_grad=True)
    F.conv2d(x, wr, None, 1, k // 2).backward(dz)
    xp, dzp = _padded(x), _padded(dz)
    dw = torch.full((cout, cin, k, k), 7.0, device=DEV)
    wsb = L.stl_conv_wgrad_workspace_bytes(n, h, w, cin, cout, k, 1, cin)
    assert wsb > 0                                       # tensor-core path
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    _lib.check(L.stl_conv_wgrad(_lib.ptr(xp), _lib.ptr(dzp), _lib.ptr(dw), n, h, w, cin, cout, k, 1, cin,
                                _lib.ptr(ws), wsb, _lib.current_stream()))
    dw2 = torch.empty_like(dw)
    _lib.check(L.stl_conv_wgrad(_lib.ptr(xp), _lib.ptr(dzp), _lib.ptr(dw2), n, h, w, cin, cout, k, 1, cin,
                                _lib.ptr(ws), wsb, _lib.current_stream()))
    assert torch.equal(dw, dw2)                          # fixed-order reduction: bit-reproducible
    dn, dn2 = torch.empty_like(dw), torch.empty_like(dw)
    nwb = L.stl_conv_wgrad_naive_workspace_bytes(n, h, w, cin, cout, k, 1, cin)
    nws = torch.empty(nwb, dtype=torch.uint8, device=DEV)
    for out in (dn, dn2):
        _lib.check(L.stl_conv_wgrad_naive(_lib.ptr(xp), _lib.ptr(dzp), _lib.ptr(out), n, h, w, cin, cout, k, 1, cin,
                                          _lib.ptr(nws), nwb, _lib.current_stream()))
    assert torch.equal(dn, dn2)                          # per-block slabs + fixed-order sum: bit-reproducible too
    scale = wr.grad.abs().max().item()
    assert (dn - wr.grad).abs().max().item() < 1e-3 * scale
    err = (dw - wr.grad).abs().max().item()
    assert err < 1e-3 * scale, (err, scale)


@pytest.mark.parametrize("cin,cout,hw,n", [(32, 64, (64, 48), 2), (64, 128, (32, 24), 3), (32, 32, (64, 48), 2),
                                           (128, 256, (16, 12), 3), (256, 64, (64, 48), 2), (64, 64, (128, 96), 2),
                                           (32, 128, (32, 24), 2), (64, 256, (16, 12), 2), (32, 256, (16, 12), 2),
                                           (48, 96, (96, 72), 1), (96, 192, (48, 36), 2), (192, 384, (24, 18), 2),
                                           (48, 48, (96, 72), 1), (96, 384, (48, 36), 1)])
def test_stride2_gradients_via_zero_stuffing(cin, cout, hw, n):
    """Stride-2 3x3 layers: zero-stuffed dz + the stride-1 tensor-core dgrad / wgrad kernels vs autograd."""
    from stlpose_b200 import training
    h, w = hw
    g = torch.Generator(device=DEV).manual_seed(cin + 7 * cout)
    wt = bf16_round(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (cin * 9) ** 0.5)
    x = bf16_round(torch.randn(n, cin, h, w, device=DEV, generator=g))
    dz = bf16_round(torch.randn(n, cout, h // 2, w // 2, device=DEV, generator=g))
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    F.conv2d(xr, wr, None, 2, 1).backward(dz)
    dzp = _padded(dz)
    u = training.zero_stuff(dzp, n, h, w)
    assert (u[:, h] == 0).all() and (u[:, :, w] == 0).all() and (u[:, 1::2] == 0).all() and (u[:, :, 1::2] == 0).all()
    assert torch.equal(u[:, 0:h:2, 0:w:2], dzp[:, :h // 2, :w // 2])
    dx = training.conv_dgrad(dzp, wt, n, h, w, cin, 2)
    assert (_unpadded(dx) - xr.grad).abs().max().item() < 2e-2 * max(1.0, xr.grad.abs().max().item())
    dw = training.conv_wgrad(_padded(x), dzp, wt.shape, n, h, w, 2)
    assert (dw - wr.grad).abs().max().item() < 1e-3 * wr.grad.abs().max().item()


@pytest.mark.parametrize("cin,cout,k,stride,hw,n", [
    (32, 32, 3, 1, (64, 48), 3), (64, 64, 3, 1, (32, 24), 3), (128, 128, 3, 1, (16, 12), 5), (256, 256, 3, 1, (8, 6), 7),
    (64, 256, 1, 1, (64, 48), 2), (256, 64, 1, 1, (64, 48), 2), (256, 32, 3, 1, (64, 48), 2), (16, 64, 3, 2, (128, 96), 2),
    (64, 64, 3, 2, (128, 96), 2), (32, 64, 3, 2, (64, 48), 3), (128, 256, 3, 2, (16, 12), 3), (48, 48, 3, 1, (96, 72), 2),
    (96, 192, 3, 2, (48, 36), 2), (384, 384, 3, 1, (12, 9), 3), (192, 48, 1, 1, (24, 18), 2), (32, 32, 3, 1, (8, 6), 1)])
def test_conv_epilogue_statistics(cin, cout, k, stride, hw, n):
    """stl_conv2d_stats / stl_conv2d_bn: per-channel sum / sum of squares of the stored conv output, accumulated in the
    epilogue, vs the sums of the output tensor itself; fixed-order reduction -> bit-reproducible; the normalisation that
    follows equals the stand-alone BatchNorm kernel's up to fp32 summation order.  The fused variants were measured
    slower than the separate statistics pass and are compiled only with -DSTL_CONV_STATS: in the
    default build every case skips; built with the flag (make NVCCFLAGS+=-DSTL_CONV_STATS) all 14 supported shapes pass."""
    from stlpose_b200 import training
    L = _lib.lib()
    h, w = hw
    g = torch.Generator(device=DEV).manual_seed(cin + cout + k + stride)
    cin_real = 3 if cin == 16 else cin
    x = bf16_round(torch.randn(n, cin_real, h, w, device=DEV, generator=g))
    wt = torch.randn(cout, cin_real, k, k, device=DEV, generator=g) / (cin_real * k * k) ** 0.5
    xp = to_padded(x, cin).view(torch.bfloat16).view(n, h + 1, w + 1, cin)
    wp, bp, cout_pad = training._pack_weights(wt, cin)
    part = torch.full((L.stl_conv2d_stats_floats(cout_pad),), float("nan"), device=DEV)
    z, rows = training._conv_raw(xp, wp, bp, cout, cout_pad, k, stride, stats=part)
    ho, wo = h // stride, w // stride
    if rows == 0:
        pytest.skip("no fused statistics for this shape (falls back to the reduction kernel)")
    tab = part[: rows * 2 * cout_pad].view(rows, 2, cout_pad).double().sum(0)
    zf = _unpadded(z).double()
    ref_s, ref_q = zf.sum(dim=(0, 2, 3)), (zf * zf).sum(dim=(0, 2, 3))
    assert (tab[0, :cout] - ref_s).abs().max().item() <= 1e-4 * max(1.0, ref_s.abs().max().item())
    assert (tab[1, :cout] - ref_q).abs().max().item() <= 1e-4 * ref_q.abs().max().item()
    part2 = torch.zeros_like(part)
    z2, rows2 = training._conv_raw(xp, wp, bp, cout, cout_pad, k, stride, stats=part2)
    assert rows2 == rows and torch.equal(part[: rows * 2 * cout_pad], part2[: rows * 2 * cout_pad]) and torch.equal(z, z2)
    assert (z[:, ho] == 0).all() and (z[:, :, wo] == 0).all()
    # finalize + apply vs the stand-alone statistics kernel
    gamma = torch.rand(cout, device=DEV, generator=g) + 0.5
    beta = torch.randn(cout, device=DEV, generator=g) * 0.1
    outs = []
    for fused in (True, False):
        y = torch.empty_like(z)
        mean, rstd = torch.empty(cout, device=DEV), torch.empty(cout, device=DEV)
        rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
        if fused:
            _lib.check(L.stl_bn_train_forward_fused(_lib.ptr(z), _lib.ptr(part), rows, cout_pad, _lib.ptr(gamma), _lib.ptr(beta),
                                                    None, 1, 1e-5, 0.1, n, ho, wo, cout, _lib.ptr(y), _lib.ptr(mean),
                                                    _lib.ptr(rstd), _lib.ptr(rm), _lib.ptr(rv), _lib.current_stream()))
        else:
            sums = torch.empty(L.stl_bn_workspace_floats(cout), device=DEV)
            _lib.check(L.stl_bn_train_forward(_lib.ptr(z), _lib.ptr(gamma), _lib.ptr(beta), None, 1, 1e-5, 0.1, n, ho, wo,
                                              cout, _lib.ptr(y), _lib.ptr(sums), _lib.ptr(mean), _lib.ptr(rstd),
                                              _lib.ptr(rm), _lib.ptr(rv), _lib.current_stream()))
        outs.append((y, mean, rstd, rm, rv))
    (ya, ma, ra, rma, rva), (yb, mb, rb, rmb, rvb) = outs
    assert (ma - mb).abs().max().item() < 1e-5 * max(1.0, mb.abs().max().item())
    assert ((ra - rb).abs() / rb).max().item() < 1e-4
    assert (rma - rmb).abs().max().item() < 1e-5 and ((rva - rvb).abs() / rvb).max().item() < 1e-4
    assert (_unpadded(ya) - _unpadded(yb)).abs().max().item() <= 2 ** -7 * max(1.0, _unpadded(yb).abs().max().item())
    # stl_conv2d_bn: the convolution's last CTA finalises the statistics itself (twice: the ticket must be left at zero)
    tickets = torch.zeros(8, dtype=torch.int32, device=DEV)
    for _ in range(2):
        mean, rstd = torch.full((cout,), float("nan"), device=DEV), torch.full((cout,), float("nan"), device=DEV)
        rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
        z3, done = training._conv_raw(xp, wp, bp, cout, cout_pad, k, stride,
                                      bn=(part2, tickets, 1e-5, 0.1, mean, rstd, rm, rv))
        if not done:
            break
        assert torch.equal(z3, z) and int(tickets[0]) == 0
        # (same rows, another grouping of the fixed-order sum than the stand-alone finalize kernel)
        assert (mean - ma).abs().max().item() < 1e-5 * max(1.0, ma.abs().max().item())
        assert ((rstd - ra).abs() / ra).max().item() < 1e-4 and ((rv - rva).abs() / rva).max().item() < 1e-4
        assert (rm - rma).abs().max().item() < 1e-5


@pytest.mark.parametrize("shape", [(3, 32, 16, 12), (2, 64, 32, 24), (5, 48, 12, 9), (2, 256, 8, 6)])
def test_bn_backward_mask_recomputed_from_z_is_bit_identical(shape):
    """ReLU units without residual: stl_bn_train_backward_ticket_z derives the mask (y > 0) from z with the forward's
    exact operations instead of reading y; dz and dbeta | dgamma must equal the y-reading kernel bit for bit."""
    L = _lib.lib()
    n, c, h, w = shape
    g = torch.Generator(device=DEV).manual_seed(c * 7 + h)
    z = bf16_round(torch.randn(shape, device=DEV, generator=g) * 1.3 + 0.2)
    gamma = torch.rand(c, device=DEV, generator=g) + 0.5
    beta = torch.randn(c, device=DEV, generator=g) * 0.3
    zp = _padded(z)
    y = torch.empty_like(zp)
    sums = torch.empty(L.stl_bn_workspace_floats(c), device=DEV); mean = torch.empty(c, device=DEV); rstd = torch.empty(c, device=DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    _lib.check(L.stl_bn_train_forward(_lib.ptr(zp), _lib.ptr(gamma), _lib.ptr(beta), None, 1, 1e-5, 0.1, n, h, w, c,
                                      _lib.ptr(y), _lib.ptr(sums), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(rm), _lib.ptr(rv),
                                      _lib.current_stream()))
    assert 0.2 < (y[:, :h, :w] > 0).float().mean().item() < 0.8           # the mask is non-trivial
    dyp = _padded(bf16_round(torch.randn(shape, device=DEV, generator=g)))
    outs = []
    for from_z in (False, True):
        dz = torch.empty_like(zp)
        dbg = torch.empty(2 * c, device=DEV)
        ws = torch.empty(L.stl_bn_workspace_floats(c), device=DEV)
        ticket = torch.zeros(2, dtype=torch.int32, device=DEV)
        if from_z:
            _lib.check(L.stl_bn_train_backward_ticket_z(_lib.ptr(dyp), _lib.ptr(zp), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
                                                        _lib.ptr(beta), n, h, w, c, _lib.ptr(dz), _lib.ptr(dbg), _lib.ptr(ws),
                                                        ticket.data_ptr(), _lib.current_stream()))
        else:
            _lib.check(L.stl_bn_train_backward_ticket(_lib.ptr(dyp), _lib.ptr(y), _lib.ptr(zp), _lib.ptr(mean), _lib.ptr(rstd),
                                                      _lib.ptr(gamma), 1, n, h, w, c, _lib.ptr(dz), None, _lib.ptr(dbg),
                                                      _lib.ptr(ws), ticket.data_ptr(), _lib.current_stream()))
        outs.append((dz, dbg))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("shape", [(3, 32, 16, 12), (2, 64, 32, 24), (5, 48, 12, 9), (2, 256, 8, 6), (6, 32, 64, 48)])
@pytest.mark.parametrize("relu,with_res", [(True, True), (True, False), (False, False)])
def test_cooperative_bn_kernels_equal_the_separate_launches(shape, relu, with_res):
    """stl_bn_train_forward_coop / _backward_coop (reduction, in-kernel hand-over, normalisation in one cooperative
    launch) vs the two-launch entry points: identical outputs, statistics and gradients, bit for bit; the ticket and
    hand-over words are left at zero (second call on the same words)."""
    L = _lib.lib()
    n, c, h, w = shape
    g = torch.Generator(device=DEV).manual_seed(c * 3 + h)
    z = bf16_round(torch.randn(shape, device=DEV, generator=g) * 1.2 + 0.1)
    res = _padded(bf16_round(torch.randn(shape, device=DEV, generator=g))) if with_res else None
    gamma = torch.rand(c, device=DEV, generator=g) + 0.5
    beta = torch.randn(c, device=DEV, generator=g) * 0.2
    zp = _padded(z)
    dyp = _padded(bf16_round(torch.randn(shape, device=DEV, generator=g)))
    words = torch.zeros(8, dtype=torch.int32, device=DEV)

    def run(coop):
        y = torch.empty_like(zp)
        sums = torch.empty(L.stl_bn_workspace_floats(c), device=DEV)
        mean, rstd = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        if coop:
            _lib.check(L.stl_bn_train_forward_coop(_lib.ptr(zp), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(res), int(relu),
                                                   1e-5, 0.1, n, h, w, c, _lib.ptr(y), _lib.ptr(sums), _lib.ptr(mean),
                                                   _lib.ptr(rstd), _lib.ptr(rm), _lib.ptr(rv), words.data_ptr(),
                                                   words.data_ptr() + 8, _lib.current_stream()))
        else:
            _lib.check(L.stl_bn_train_forward_ticket(_lib.ptr(zp), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(res), int(relu),
                                                     1e-5, 0.1, n, h, w, c, _lib.ptr(y), _lib.ptr(sums), _lib.ptr(mean),
                                                     _lib.ptr(rstd), _lib.ptr(rm), _lib.ptr(rv), words.data_ptr(),
                                                     _lib.current_stream()))
        dz = torch.empty_like(zp)
        dres = torch.empty_like(zp) if with_res else None
        dbg = torch.empty(2 * c, device=DEV)
        ws = torch.empty(L.stl_bn_workspace_floats(c), device=DEV)
        mode = 0 if not relu else (1 if with_res else 2)
        if coop:
            _lib.check(L.stl_bn_train_backward_coop(_lib.ptr(dyp), _lib.ptr(y) if mode == 1 else None, _lib.ptr(zp),
                                                    _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma), _lib.ptr(beta), mode, n, h,
                                                    w, c, _lib.ptr(dz), _lib.ptr(dres), _lib.ptr(dbg), _lib.ptr(ws),
                                                    words.data_ptr() + 4, words.data_ptr() + 16, _lib.current_stream()))
        elif mode == 2:
            _lib.check(L.stl_bn_train_backward_ticket_z(_lib.ptr(dyp), _lib.ptr(zp), _lib.ptr(mean), _lib.ptr(rstd),
                                                        _lib.ptr(gamma), _lib.ptr(beta), n, h, w, c, _lib.ptr(dz), _lib.ptr(dbg),
                                                        _lib.ptr(ws), words.data_ptr() + 4, _lib.current_stream()))
        else:
            _lib.check(L.stl_bn_train_backward_ticket(_lib.ptr(dyp), _lib.ptr(y), _lib.ptr(zp), _lib.ptr(mean), _lib.ptr(rstd),
                                                      _lib.ptr(gamma), int(relu), n, h, w, c, _lib.ptr(dz), _lib.ptr(dres),
                                                      _lib.ptr(dbg), _lib.ptr(ws), words.data_ptr() + 4, _lib.current_stream()))
        torch.cuda.synchronize()
        assert int(words.abs().sum()) == 0
        return y, mean, rstd, rm, rv, dz, dres, dbg

    ref = run(False)
    for _ in range(2):
        got = run(True)
        for a, b in zip(got, ref):
            assert (a is None and b is None) or torch.equal(a, b)
