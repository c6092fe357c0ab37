"""Train-mode forward and backward of the HRNet pose network (the fine-tuning step of the reference,
/root/reference/src/02_train.py:203-218: ``model.train()``, ``forward_pass``, ``loss.backward()``, ``optimizer.step()``).

BatchNorm cannot be folded in training mode (batch statistics, HRnet.py:38-59 under ``model.train()``), so every
conv + BN [+ residual] [+ ReLU] unit is an autograd node of three device steps:

    z = conv(x, W)                      tcgen05 implicit-GEMM kernel, raw bf16 output
    mean, var over N*H*W of z           per-channel reduction; running statistics updated like nn.BatchNorm2d
    y = [relu](gamma*(z-mean)*rstd + beta [+ residual])

and its backward (BN backward, conv dgrad, conv wgrad).  torch.autograd provides the graph (fan-out accumulation,
ordering); every tensor operation on an activation is one of this library's kernels.  Activations and their gradients
stay in the engine's padded-linear NHWC bf16 layout; parameter gradients are fp32 tensors shaped like the parameters,
so ``torch.optim`` works unchanged.
"""
import ctypes
import os

import torch

from . import _lib

BN_EPS = 1e-5


def _stream():
    return _lib.current_stream()


def _padded_zeros(n, h, w, c, device):
    return torch.zeros((n, h + 1, w + 1, c), dtype=torch.bfloat16, device=device)


_ZERO_BIAS = {}
_CONV_CALLS = 0        # number of _conv_raw calls so far (see _zero_bias)


def _zero_bias(n, device):
    """Shared all-zero fp32 bias (BatchNorm follows every conv of the training graph; only the head has a bias)."""
    key = (device, n)
    if key not in _ZERO_BIAS:
        t = _ZERO_BIAS[key] = torch.zeros(n, dtype=torch.float32, device=device)
        t._stl_fresh_call = _CONV_CALLS     # its fill kernel was just enqueued: the next convolution must not pre-load it
    return _ZERO_BIAS[key]


# Weights packed ahead of time by one batched launch per step (see _Prepack): (data_ptr, cols_pad, dgrad) ->
# (packed view, rows_pad, weight version at pack time, weakref to the weight).
_PREPACKED = {}


def _prepacked(w, cols_pad, dgrad):
    hit = _PREPACKED.get((w.data_ptr(), cols_pad, dgrad))
    if hit is not None and hit[2] == w._version and hit[3]() is w:
        return hit[0], hit[1]
    return None


class _Prepack:
    """All forward / input-gradient weight layouts of a model's BatchNorm-followed convolutions, re-packed from the fp32
    masters by ONE launch at the start of every training forward (stl_pack_conv_weights_batched) instead of one launch
    per layer and direction (585 per step).  Entries are only used while the weight's version counter is unchanged."""

    def __init__(self, model, device):
        import weakref
        import numpy as np
        convs = [m for m in model.modules() if isinstance(m, torch.nn.Conv2d) and m is not model.final_layer]
        self.weights = [c.weight for c in convs]
        self.ptrs = tuple(w.data_ptr() for w in self.weights)
        self.device = device
        specs, off = [], 0
        for w in self.weights:
            cout, cin, k, _ = w.shape
            cin_pad, cout_pad = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
            layouts = [(0, cout_pad, cin_pad)]
            if cin == cin_pad:
                layouts.append((1, cin_pad, cout_pad))       # stride-1 dgrad as a convolution (conv_dgrad)
            for dgrad, rows, cols in layouts:
                nbytes = k * k * rows * cols * 2
                specs.append((w, dgrad, rows, cols, off, nbytes))
                off += (nbytes + 255) // 256 * 256
        self.arena = torch.empty(off, dtype=torch.uint8, device=device)
        item = np.dtype([("w", "<u8"), ("wp", "<u8"), ("Cout", "<i4"), ("Cin", "<i4"), ("k", "<i4"), ("rows", "<i4"),
                         ("cols", "<i4"), ("dgrad", "<i4")])
        items = np.zeros(len(specs), dtype=item)
        offsets = np.zeros(len(specs) + 1, dtype=np.int32)
        self.entries = []
        for i, (w, dgrad, rows, cols, o, nbytes) in enumerate(specs):
            cout, cin, k, _ = w.shape
            items[i] = (w.data_ptr(), self.arena.data_ptr() + o, cout, cin, k, rows, cols, dgrad)
            offsets[i + 1] = offsets[i] + (k * k * rows * cols + 1023) // 1024
            self.entries.append(((w.data_ptr(), cols, dgrad), self.arena[o:o + nbytes], rows, weakref.ref(w)))
        self.items = torch.from_numpy(items.view(np.uint8).copy()).to(device)
        self.offsets = torch.from_numpy(offsets).to(device)
        self.n, self.blocks = len(specs), int(offsets[-1])

    def valid_for(self, device):
        return self.device == device and self.ptrs == tuple(w.data_ptr() for w in self.weights) and \
            all(w.dtype == torch.float32 and w.is_contiguous() for w in self.weights)

    def run(self):
        _lib.check(_lib.lib().stl_pack_conv_weights_batched(_lib.ptr(self.items), _lib.ptr(self.offsets), self.n,
                                                            self.blocks, _stream()))
        dead = [k for k, v in _PREPACKED.items() if v[3]() is None]      # weights of models that no longer exist
        for k in dead:
            del _PREPACKED[k]
        for key, view, rows, ref in self.entries:
            _PREPACKED[key] = (view, rows, ref()._version, ref)


def _pack_weights(w, cin_pad, own_bias=False):
    """fp32 OIHW -> ([k*k][cout_pad][cin_pad] bf16 buffer, zero fp32 bias)."""
    L = _lib.lib()
    cout, cin, k, _ = w.shape
    cout_pad = (cout + 15) // 16 * 16
    if not own_bias:
        hit = _prepacked(w, cin_pad, 0)
        if hit is not None and hit[1] == cout_pad:
            return hit[0], _zero_bias(cout_pad, w.device), cout_pad
    wp = torch.empty(k * k * cout_pad * cin_pad * 2, dtype=torch.uint8, device=w.device)
    bp = torch.empty(cout_pad, dtype=torch.float32, device=w.device) if own_bias else None
    w32 = w.detach().float().contiguous()
    scratch = bp if own_bias else torch.empty(cout_pad, dtype=torch.float32, device=w.device)
    _lib.check(L.stl_pack_conv_weights(_lib.ptr(w32), None, None, None, None, None, 0.0, cout, cin, k, cout_pad,
                                       cin_pad, _lib.ptr(wp), _lib.ptr(scratch), _stream()))
    wp._stl_fresh = True      # written by the kernel just enqueued: the consuming convolution must not pre-load it (no PDL)
    return wp, (bp if own_bias else _zero_bias(cout_pad, w.device)), cout_pad


def _pack_weights_dgrad(w, k_pad):
    """fp32 OIHW of a forward layer -> weights of the convolution computing its stride-1 input gradient."""
    L = _lib.lib()
    cout, cin, k, _ = w.shape
    rows_pad = (cin + 15) // 16 * 16
    hit = _prepacked(w, k_pad, 1)
    if hit is not None and hit[1] == rows_pad:
        return hit[0], _zero_bias(rows_pad, w.device), rows_pad
    wp = torch.empty(k * k * rows_pad * k_pad * 2, dtype=torch.uint8, device=w.device)
    w32 = w.detach().float().contiguous()
    _lib.check(L.stl_pack_conv_weights_dgrad(_lib.ptr(w32), cout, cin, k, rows_pad, k_pad, _lib.ptr(wp), None, _stream()))
    wp._stl_fresh = True
    return wp, _zero_bias(rows_pad, w.device), rows_pad


# BatchNorm statistics accumulated by the convolution's epilogue and finalised by its last CTA (stl_conv2d_bn).  Measured
# slower than the separate statistics pass on B200 in both forms: off, and the kernel variants are
# only in -DSTL_CONV_STATS builds (without them stl_conv2d_bn reports "not done" and the separate pass runs anyway).
FUSED_BN_STATS = os.environ.get("STLPOSE_FUSED_BN_STATS", "0") == "1"
STEM_IM2COL = os.environ.get("STLPOSE_TRAIN_STEM_IM2COL", "1") != "0"
MASK_FROM_Z = os.environ.get("STLPOSE_TRAIN_MASK_FROM_Z", "1") != "0"
# statistics + normalisation / both backward passes of a BatchNorm layer as ONE cooperative launch per direction.  The
# backward pair re-reads dy and z right after the reduction read them, which hits L2 inside one launch (104 vs 120 us at
# 52 MB, 24 vs 35 us at 13 MB); the forward pair gains nothing but the saved graph node and its in-kernel hand-over makes
# it slower than the two launches (15 vs 13 us at 6 MB, 69 vs 65 us at 52 MB).  Measured per step (batch 32 / 128):
# neither 17.6 / 42.8 ms, both for tensors <= 16 MB 17.2-17.5 / 42.9-43.1, backward only at every size 16.75 / 41.3.
# "auto" = that last policy, and only in a single-GPU process - a cooperative grid that spins for its last blocks next to
# in-flight NCCL kernels is not something to rely on.  "1" / "0" force both directions on / off.
COOP_BN = os.environ.get("STLPOSE_TRAIN_COOP_BN", "auto")
# size limits of "auto" per direction, in MB (measurement knobs)
_COOP_BN_MAX_BYTES = (int(os.environ.get("STLPOSE_TRAIN_COOP_BN_FWD_MB", "0")) << 20,
                      int(os.environ.get("STLPOSE_TRAIN_COOP_BN_BWD_MB", "1000000")) << 20)
# programmatic dependent launch for the convolutions of the training path (stl_conv_desc.pdl): their packed weights are
# written once at the start of a step, never by the kernel in front of them.  (The BatchNorm kernels read the same switch
# in the library.)  Measured on B200: no gain (17.6-17.8 vs 17.3-17.6 ms per step at batch 32) - off by default.
TRAIN_PDL = os.environ.get("STLPOSE_TRAIN_PDL", "0") == "1"


def _use_coop_bn(z, backward=False):
    if COOP_BN in ("0", "1"):
        return COOP_BN == "1"
    if z.numel() * z.element_size() > _COOP_BN_MAX_BYTES[1 if backward else 0]:
        return False
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)


def _conv_raw(x, wp, bp, cout, cout_pad, k, stride, out=None, out_nchw=False, bias=None, stats=None, bn=None):
    """Tensor-core convolution on a padded bf16 tensor; returns padded bf16 [N,Ho+1,Wo+1,cout] (or fp32 NCHW).
    stats: an fp32 buffer of stl_conv2d_stats_floats(cout_pad) elements -> returns (out, rows): the epilogue also left
    `rows` rows of per-channel sum / sum of squares there (rows == 0: not for this shape).
    bn: (stats buffer, tickets, eps, momentum, mean, rstd, running_mean, running_var) -> returns (out, done): the last
    CTA also finalised the BatchNorm statistics (done False: not for this shape, nothing written)."""
    L = _lib.lib()
    n, hp, wpd, cin = x.shape
    h, w = hp - 1, wpd - 1
    ho, wo = h // stride, w // stride
    d = _lib.ConvDesc()
    d.in_ = x.data_ptr(); d.N, d.H, d.W, d.Cin = n, h, w, cin
    if out is None:
        if out_nchw:
            out = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=x.device)
        elif stride == 1:      # flat-pixel tiles write every cell of the output, zero cells included
            out = torch.empty((n, ho + 1, wo + 1, cout), dtype=torch.bfloat16, device=x.device)
        else:                  # structured stride-2 tiles write the valid pixels only
            out = _padded_zeros(n, ho, wo, cout, x.device)
    d.out = out.data_ptr(); d.Cout, d.Cout_pad = cout, cout_pad
    d.ksize, d.stride = k, stride
    d.w_packed = wp.data_ptr(); d.bias_packed = (bias if bias is not None else bp).data_ptr()
    d.relu = 0; d.out_nchw = int(out_nchw)
    global _CONV_CALLS
    d.pdl = int(TRAIN_PDL and not getattr(wp, "_stl_fresh", False) and bias is None and
                getattr(bp, "_stl_fresh_call", -1) != _CONV_CALLS)
    _CONV_CALLS += 1
    if stats is not None:
        rows = ctypes.c_int(0)
        _lib.check(L.stl_conv2d_stats(ctypes.byref(d), _lib.ptr(stats), ctypes.byref(rows), _stream()))
        return out, rows.value
    if bn is not None:
        part, tickets, eps, momentum, mean, rstd, run_mean, run_var = bn
        done = ctypes.c_int(0)
        _lib.check(L.stl_conv2d_bn(ctypes.byref(d), _lib.ptr(part), tickets.data_ptr(), eps, momentum, _lib.ptr(mean),
                                   _lib.ptr(rstd), _lib.ptr(run_mean), _lib.ptr(run_var), ctypes.byref(done), _stream()))
        return out, bool(done.value)
    _lib.check(L.stl_conv2d(ctypes.byref(d), _stream()))
    return out


def zero_stuff(dz, n, h, w):
    """dz of a stride-2 convolution -> full-resolution tensor with dz at the even positions and zeros elsewhere; the
    stride-2 gradients are then the stride-1 gradients (tensor-core kernels) of the stuffed tensor."""
    u = torch.empty((n, h + 1, w + 1, dz.shape[3]), dtype=torch.bfloat16, device=dz.device)
    _lib.check(_lib.lib().stl_zero_stuff(_lib.ptr(dz), _lib.ptr(u), n, h, w, dz.shape[3], _stream()))
    return u


_SIDE_STREAMS = {}
SIDE_WGRAD = os.environ.get("STLPOSE_SIDE_WGRAD", "1") != "0"


def _side_stream(device):
    """One extra stream per (device, current stream): the weight-gradient kernels of a layer run there, next to its
    input-gradient convolution on the current stream (both only need dz); forked and joined inside the layer's backward,
    so a captured step records it as two parallel branches."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return s


_BRANCH_STREAMS = {}
BRANCH_STREAMS = os.environ.get("STLPOSE_TRAIN_BRANCH_STREAMS", "1") != "0"


def _branch_stream(device, b):
    """Stream of branch b >= 1 of a HighResolutionModule: the branches of a module are independent chains of small
    kernels (at fine-tuning batch sizes the low-resolution ones fill a fraction of the SMs), so they run side by side;
    autograd replays each backward node on its forward stream, so the backward is parallel in the same way."""
    s = _BRANCH_STREAMS.get((device, b))
    if s is None:
        s = _BRANCH_STREAMS[(device, b)] = torch.cuda.Stream(device=device)
    return s


def conv_wgrad(x, dz, weight_shape, n, h, w, stride):
    """Gradient wrt the OIHW weights (fp32).  `dz` may already be zero-stuffed for a stride-2 layer."""
    dw, _ = _conv_wgrad(x, dz, weight_shape, n, h, w, stride)
    return dw if dw.shape[0] == weight_shape[0] else dw[:weight_shape[0]].contiguous()


def _conv_wgrad(x, dz, weight_shape, n, h, w, stride, side=None, out=None):
    """-> (dw with cout_pad rows, workspace to keep alive).  side: a stream to launch on; the buffers are still allocated
    on the current stream and the caller joins the side stream before dw is used or the workspace released.
    out: fp32 [cout_pad, cin_real, k, k] destination (a view into a gradient bucket) instead of a fresh tensor."""
    L = _lib.lib()
    cout, cin_real, k, _ = weight_shape
    cin_pad, cout_pad = x.shape[3], dz.shape[3]
    dw = out if out is not None else torch.empty((cout_pad, cin_real, k, k), dtype=torch.float32, device=x.device)
    stuffed = dz.shape[1] == h + 1
    if stride == 2 and not stuffed and cin_pad >= 32 and cout_pad >= 32:    # (the 3-channel stem stays on its own kernel)
        dz, stuffed = zero_stuff(dz, n, h, w), True
    s_eff = 1 if stuffed else stride
    ws_bytes = L.stl_conv_wgrad_workspace_bytes(n, h, w, cin_pad, cout_pad, k, s_eff, cin_real)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    if side is not None:
        side.wait_stream(torch.cuda.current_stream(x.device))          # dz (and everything allocated above) is ready
    _lib.check(L.stl_conv_wgrad(_lib.ptr(x), _lib.ptr(dz), _lib.ptr(dw), n, h, w, cin_pad, cout_pad, k, s_eff, cin_real,
                                _lib.ptr(ws), ws_bytes, side.cuda_stream if side is not None else _stream()))
    return dw, ws


def conv_dgrad(dz, weight, n, h, w, cin_pad, stride):
    """Gradient of a pad-k//2 convolution wrt its input, on padded bf16 tensors: dz [n,h/s+1,w/s+1,cout_pad] ->
    dx [n,h+1,w+1,cin_pad].  Stride 1 is itself a convolution of dz with the spatially flipped, channel-transposed
    filter and runs on the tcgen05 kernel; stride 2 (30 of the 293 layers) uses the CUDA-core gather kernel."""
    L = _lib.lib()
    cout, cin_real, k, _ = weight.shape
    if stride == 2 and cin_real == cin_pad and dz.shape[1] != h + 1:
        dz = zero_stuff(dz, n, h, w)
    if cin_real == cin_pad and dz.shape[1] == h + 1:
        wp, bp, cpad = _pack_weights_dgrad(weight, dz.shape[3])                     # flipped, transposed filter
        return _conv_raw(dz, wp, bp, cin_real, cpad, k, 1)
    wp, _, cout_pad = _pack_weights(weight, cin_pad)
    dx = torch.empty((n, h + 1, w + 1, cin_pad), dtype=torch.bfloat16, device=dz.device)
    _lib.check(L.stl_conv_dgrad(_lib.ptr(dz), _lib.ptr(wp), _lib.ptr(dx), n, h, w, cin_pad, cout_pad, k, stride, _stream()))
    return dx


class _ConvBN(torch.autograd.Function):
    """conv (no bias) + train-mode BatchNorm [+ residual] [+ ReLU] on padded bf16 activations."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, run_mean, run_var, stride, relu, momentum, tickets, sinks=None):
        L = _lib.lib()
        n, hp, wpd, cin_pad = x.shape
        h, w = hp - 1, wpd - 1
        cout, cin_real, k, _ = weight.shape
        wp, bp, cout_pad = _pack_weights(weight, cin_pad)
        ho, wo = h // stride, w // stride
        mean = torch.empty(cout, dtype=torch.float32, device=x.device)
        rstd = torch.empty(cout, dtype=torch.float32, device=x.device)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        done = False
        if FUSED_BN_STATS and cout == cout_pad:
            # the convolution's epilogue accumulates the batch statistics of z and its last CTA finalises mean / rstd /
            # running statistics (forward ticket of this layer): conv -> normalise, nothing in between
            part = torch.empty(L.stl_conv2d_stats_floats(cout_pad), dtype=torch.float32, device=x.device)
            z, done = _conv_raw(x, wp, bp, cout, cout_pad, k, stride,
                                bn=(part, tickets, BN_EPS, float(momentum), mean, rstd, run_mean, run_var))
        else:
            z = _conv_raw(x, wp, bp, cout, cout_pad, k, stride)
        y = torch.empty_like(z)
        if done:
            _lib.check(L.stl_bn_apply(_lib.ptr(z), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(g32), _lib.ptr(b32),
                                      _lib.ptr(residual), int(relu), n, ho, wo, cout, _lib.ptr(y), _stream()))
        else:
            sums = torch.empty(L.stl_bn_workspace_floats(cout), dtype=torch.float32, device=x.device)
            if _use_coop_bn(z):
                _lib.check(L.stl_bn_train_forward_coop(_lib.ptr(z), _lib.ptr(g32), _lib.ptr(b32), _lib.ptr(residual),
                                                       int(relu), BN_EPS, float(momentum), n, ho, wo, cout, _lib.ptr(y),
                                                       _lib.ptr(sums), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(run_mean),
                                                       _lib.ptr(run_var), tickets.data_ptr(), tickets.data_ptr() + 8,
                                                       _stream()))
            else:
                _lib.check(L.stl_bn_train_forward_ticket(_lib.ptr(z), _lib.ptr(g32), _lib.ptr(b32), _lib.ptr(residual),
                                                         int(relu), BN_EPS, float(momentum), n, ho, wo, cout, _lib.ptr(y),
                                                         _lib.ptr(sums), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(run_mean),
                                                         _lib.ptr(run_var), tickets.data_ptr(), _stream()))
        ctx.tickets = tickets
        ctx.sinks = sinks
        ctx.save_for_backward(x, weight, z, y, mean, rstd, g32, b32)
        ctx.meta = (n, h, w, cin_pad, cin_real, cout, k, stride, bool(relu), residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        x, weight, z, y, mean, rstd, g32, b32 = ctx.saved_tensors
        n, h, w, cin_pad, cin_real, cout, k, stride, relu, has_res = ctx.meta
        ho, wo = h // stride, w // stride
        dy = dy.contiguous()
        dz = torch.empty_like(z)
        dres = torch.empty_like(z) if has_res else None
        ws = torch.empty(L.stl_bn_workspace_floats(cout), dtype=torch.float32, device=x.device)
        # parameter gradients: fresh tensors handed to autograd, or - when the parameters are bound to a gradient bucket
        # (parallel.GradientReducer.bind) - written straight into the bucket by the kernels, nothing returned
        wsink, bsink = ctx.sinks if ctx.sinks is not None else (None, None)
        sums = bsink.view if bsink is not None else torch.empty(2 * cout, dtype=torch.float32, device=x.device)   # dbeta | dgamma
        if _use_coop_bn(z, backward=True):
            mode = 0 if not relu else (2 if (not has_res and MASK_FROM_Z) else 1)
            _lib.check(L.stl_bn_train_backward_coop(_lib.ptr(dy), _lib.ptr(y) if mode == 1 else None, _lib.ptr(z), _lib.ptr(mean),
                                                    _lib.ptr(rstd), _lib.ptr(g32), _lib.ptr(b32), mode, n, ho, wo, cout,
                                                    _lib.ptr(dz), _lib.ptr(dres), _lib.ptr(sums), _lib.ptr(ws),
                                                    ctx.tickets.data_ptr() + 4, ctx.tickets.data_ptr() + 16, _stream()))
        elif relu and not has_res and MASK_FROM_Z:
            # ReLU unit without residual: y > 0 <=> gamma * (z - mean) * rstd + beta > 0, recomputed from z with the
            # forward's exact operations - one tensor read less in each of the two backward passes
            _lib.check(L.stl_bn_train_backward_ticket_z(_lib.ptr(dy), _lib.ptr(z), _lib.ptr(mean), _lib.ptr(rstd),
                                                        _lib.ptr(g32), _lib.ptr(b32), n, ho, wo, cout, _lib.ptr(dz),
                                                        _lib.ptr(sums), _lib.ptr(ws), ctx.tickets.data_ptr() + 4, _stream()))
        else:
            _lib.check(L.stl_bn_train_backward_ticket(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(z), _lib.ptr(mean), _lib.ptr(rstd),
                                                      _lib.ptr(g32), int(relu), n, ho, wo, cout, _lib.ptr(dz), _lib.ptr(dres),
                                                      _lib.ptr(sums), _lib.ptr(ws), ctx.tickets.data_ptr() + 4, _stream()))
        dbeta, dgamma = sums[:cout], sums[cout:]
        if bsink is not None:
            bsink.done()
        if stride == 2 and cin_pad >= 32:
            dz = zero_stuff(dz, n, h, w)               # shared by dgrad and wgrad
        # dgrad (main stream) and wgrad (side stream) both only need dz: two parallel branches of the step
        side = _side_stream(x.device) if (SIDE_WGRAD and ctx.needs_input_grad[0]) else None
        direct = wsink is not None and dz.shape[3] == cout
        dw, keep = _conv_wgrad(x, dz, weight.shape, n, h, w, stride, side=side, out=wsink.view if direct else None)
        dx = conv_dgrad(dz, weight, n, h, w, cin_pad, stride) if ctx.needs_input_grad[0] else None
        if side is not None:
            torch.cuda.current_stream(x.device).wait_stream(side)       # join before dw / the workspace are touched again
        del keep
        if wsink is not None:
            if not direct:
                wsink.view.copy_(dw[:cout])
            wsink.done()
            return dx, None, None if bsink is not None else dgamma, None if bsink is not None else dbeta, dres, \
                None, None, None, None, None, None, None
        if dw.shape[0] != cout:
            dw = dw[:cout].contiguous()
        return dx, dw, None if bsink is not None else dgamma, None if bsink is not None else dbeta, dres, \
            None, None, None, None, None, None, None


class _Head(torch.autograd.Function):
    """final_layer: 1x1 conv with bias, no activation, fp32 NCHW out (HRnet.py:331-337, 466)."""

    @staticmethod
    def forward(ctx, x, weight, bias, sinks=None):
        n, hp, wpd, cin = x.shape
        cout = weight.shape[0]
        ctx.sinks = sinks
        wp, bp, cout_pad = _pack_weights(weight, cin, own_bias=True)
        bp[:cout] = bias.detach().float()
        heat = _conv_raw(x, wp, bp, cout, cout_pad, 1, 1, out_nchw=True)
        ctx.save_for_backward(x, weight)
        ctx.meta = (n, hp - 1, wpd - 1, cin, cout, cout_pad)
        return heat

    @staticmethod
    def backward(ctx, dheat):
        L = _lib.lib()
        x, weight = ctx.saved_tensors
        n, h, w, cin, cout, cout_pad = ctx.meta
        dheat = dheat.contiguous().float()
        dz = torch.empty((n, h + 1, w + 1, cout_pad), dtype=torch.bfloat16, device=x.device)
        _lib.check(L.stl_nchw_to_padded(_lib.ptr(dheat), _lib.ptr(dz), n, cout, h, w, cout_pad, _stream()))
        dx = conv_dgrad(dz, weight, n, h, w, cin, 1)
        dw = conv_wgrad(x, dz, weight.shape, n, h, w, 1)
        db = dheat.sum(dim=(0, 2, 3))
        if ctx.sinks is not None:                      # parameters bound to a gradient bucket (see _ConvBN.backward)
            wsink, bsink = ctx.sinks
            wsink.view.copy_(dw)
            bsink.view.copy_(db)
            wsink.done()
            bsink.done()
            return dx, None, None, None
        return dx, dw, db, None


class _FuseSum(torch.autograd.Function):
    """One output row of HighResolutionModule.forward (HRnet.py:255-264): relu(sum of same-resolution terms +
    nearest-upsampled low-resolution terms)."""

    @staticmethod
    def forward(ctx, n_same, shifts, *tensors):
        L = _lib.lib()
        same, ups = tensors[:n_same], tensors[n_same:]
        n, hp, wpd, c = same[0].shape
        y = torch.empty_like(same[0])
        sp = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in same])
        up = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in ups])
        sh = (ctypes.c_int * 3)(*shifts)
        _lib.check(L.stl_sum_relu_forward(sp, len(same), up, sh, len(ups), _lib.ptr(y), n, hp - 1, wpd - 1, c,
                                          _stream()))
        ctx.save_for_backward(y)
        ctx.meta = (n_same, tuple(shifts), [tuple(t.shape) for t in ups])
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        (y,) = ctx.saved_tensors
        n_same, shifts, up_shapes = ctx.meta
        n, hp, wpd, c = y.shape
        g = torch.empty_like(y)
        _lib.check(L.stl_relu_mask(_lib.ptr(dy.contiguous()), _lib.ptr(y), _lib.ptr(g), y.numel(), _stream()))
        grads = [g] * n_same
        for s, shp in zip(shifts, up_shapes):
            dlow = torch.empty(shp, dtype=torch.bfloat16, device=y.device)
            _lib.check(L.stl_upsample_backward(_lib.ptr(g), _lib.ptr(dlow), n, hp - 1, wpd - 1, c, s, _stream()))
            grads.append(dlow)
        return (None, None, *grads)


def _tickets(bn, device):
    """Zero-initialised device words per BatchNorm layer (forward / backward reduction tickets and the hand-over words of
    the cooperative kernels, see stl_bn_train_*_ticket / _coop): the kernels leave them zero, so no memset per call."""
    t = getattr(bn, "_stl_tickets", None)
    if t is None or t.device != device:
        t = torch.zeros(8, dtype=torch.int32, device=device)   # [0] fwd ticket, [1] bwd ticket, [2:4] / [4:6] fwd / bwd hand-over
        bn._stl_tickets = t
    return t


def _sinks(weight, bias):
    """Gradient destinations of a (conv weight, BatchNorm bias [+ weight]) or (head weight, head bias) pair when the
    parameters are bound to a gradient bucket (parallel.GradientReducer.bind), else None."""
    ws, bs = getattr(weight, "_stl_sink", None), getattr(bias, "_stl_sink", None)
    if ws is None and bs is None:
        return None
    if ws is None or bs is None:
        raise _lib.StlError("gradient buckets are bound to only part of a layer's parameters (GradientReducer.bind)")
    return ws, bs


def _convbn(x, conv, bn, stride, relu, residual=None):
    return _ConvBN.apply(x, conv.weight, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var, stride, relu,
                         bn.momentum, _tickets(bn, x.device), _sinks(conv.weight, bn.bias))


def train_forward(model, x):
    """PoseHighResolutionNet.forward in training mode (HRnet.py:433-468) -> fp32 heatmaps [B,J,H/4,W/4] with grad_fn."""
    L = _lib.lib()
    if not x.is_cuda:
        raise _lib.StlError("input must be a CUDA tensor (there is no CPU fallback)")
    x = x.detach().float().contiguous()
    B, _, H, W = x.shape
    counters = getattr(model, "_bn_counters", None)
    if counters is None or counters[0].device != x.device:
        counters = [b for n, b in model.named_buffers() if n.endswith("num_batches_tracked")]
        model._bn_counters = counters
    with torch.cuda.device(x.device):
        pre = getattr(model, "_stl_prepack", None)
        if pre is None or not pre.valid_for(x.device):
            pre = model._stl_prepack = _Prepack(model, x.device) if all(
                p.dtype == torch.float32 and p.is_contiguous() for p in model.parameters()) else None
        if pre is not None:
            pre.run()                                                  # every weight layout of this step, one launch
        torch._foreach_add_(counters, 1)                               # every BatchNorm runs once per forward
        c1 = model.conv1.weight
        if STEM_IM2COL and tuple(c1.shape[1:]) == (3, 3, 3) and c1.is_contiguous():
            # conv1 (3 -> 64, 3x3, stride 2) as a 1x1 convolution over its im2col rows (27 -> 32 values per output pixel, K
            # order = the OIHW flatten order, so the weight is just viewed as [64, 27, 1, 1]): forward and weight gradient
            # both run on the tensor cores instead of the stride-2 kernel with 3 of 16 channels used and the CUDA-core
            # stem weight-gradient kernel (1.06 ms per step at batch 128)
            t = torch.zeros((B, H // 2 + 1, W // 2 + 1, 32), dtype=torch.bfloat16, device=x.device)
            _lib.check(L.stl_stem_im2col(_lib.ptr(x), _lib.ptr(t), B, H, W, _stream()))
            bn1 = model.bn1
            t = _ConvBN.apply(t, c1.view(c1.shape[0], 27, 1, 1), bn1.weight, bn1.bias, None, bn1.running_mean,
                              bn1.running_var, 1, True, bn1.momentum, _tickets(bn1, x.device), _sinks(c1, bn1.bias))
        else:
            t = torch.empty((B, H + 1, W + 1, 16), dtype=torch.bfloat16, device=x.device)
            _lib.check(L.stl_nchw_to_padded(_lib.ptr(x), _lib.ptr(t), B, 3, H, W, 16, _stream()))
            t = _convbn(t, model.conv1, model.bn1, 2, True)
        t = _convbn(t, model.conv2, model.bn2, 2, True)
        for blk in model.layer1:                                     # Bottleneck.forward, HRnet.py:82-102
            res = t
            o = _convbn(t, blk.conv1, blk.bn1, 1, True)
            o = _convbn(o, blk.conv2, blk.bn2, 1, True)
            if hasattr(blk, "downsample"):
                res = _convbn(t, blk.downsample[0], blk.downsample[1], 1, False)
            t = _convbn(o, blk.conv3, blk.bn3, 1, True, residual=res)
        tr = model.transition1
        xs = [_convbn(t, tr[0][0], tr[0][1], 1, True), _convbn(t, tr[1][0][0], tr[1][0][1], 2, True)]
        for stage in (2, 3, 4):
            if stage > 2:                                            # HRnet.py:450-463: new branch from y_list[-1]
                seq = getattr(model, f"transition{stage - 1}")[stage - 1][0]
                xs.append(_convbn(xs[-1], seq[0], seq[1], 2, True))
            for mod in getattr(model, f"stage{stage}"):
                xs = _hr_module(mod, xs)
        return _Head.apply(xs[0], model.final_layer.weight, model.final_layer.bias,
                           _sinks(model.final_layer.weight, model.final_layer.bias))


def _hr_module(mod, xs):
    """HighResolutionModule.forward, HRnet.py:248-266."""
    xs = list(xs)
    dev = xs[0].device
    main = torch.cuda.current_stream(dev)
    forked = []
    for b, branch in enumerate(mod.branches):
        st = main
        if BRANCH_STREAMS and b > 0:
            st = _branch_stream(dev, b)
            st.wait_stream(main)                                     # fork: the branch input is ready
            forked.append(st)
        with torch.cuda.stream(st):
            for blk in branch:                                       # BasicBlock.forward, HRnet.py:45-61
                o = _convbn(xs[b], blk.conv1, blk.bn1, 1, True)
                xs[b] = _convbn(o, blk.conv2, blk.bn2, 1, True, residual=xs[b])
    # exchange (fuse) layers: output row i needs every branch, so the streams first wait for each other; then row i runs
    # on stream i (the rows are independent of each other), and everything joins the main stream at the end
    streams = [main] + forked
    for si in streams:
        for sj in streams:
            if sj is not si:
                si.wait_stream(sj)
    if forked:
        # Branch outputs are allocated on their branch's stream and read by the fuse rows on every other stream; tell the
        # caching allocator, so that a block is never handed out again on its home stream while a kernel queued on
        # another stream still reads it (the event waits above order the kernels, record_stream orders the memory reuse).
        for j, t in enumerate(xs):
            for si in streams:
                if si is not (streams[j] if j < len(streams) else main):
                    t.record_stream(si)
    outs = []
    for i, row in enumerate(mod.fuse_layers):
        with torch.cuda.stream(streams[i] if i < len(streams) else main):
            same, ups, shifts = [xs[i]], [], []
            for j in range(len(xs)):
                if j > i:
                    ups.append(_convbn(xs[j], row[j][0], row[j][1], 1, False))
                    shifts.append(j - i)
                elif j < i:
                    t = xs[j]
                    hops = len(row[j])
                    for k, seq in enumerate(row[j]):
                        t = _convbn(t, seq[0], seq[1], 2, k != hops - 1)
                    same.append(t)
            outs.append(_FuseSum.apply(len(same), shifts, *same, *ups))
    for st in forked:
        main.wait_stream(st)
    for i, t in enumerate(outs):                                     # rows computed on a branch stream are consumed on main
        if 0 < i < len(streams):
            t.record_stream(main)
    return outs
