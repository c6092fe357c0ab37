"""CPU restatement of the reference HRNet forward pass (test infrastructure; see oracle/__init__.py).

A functional fp32 forward over a plain ``state_dict`` whose keys and shapes are those of the
reference ``PoseHighResolutionNet`` (/root/reference/src/models/HRnet.py:275-339), so the same
checkpoint drives the reference module, this oracle and the sm_100a implementation.

The dense arithmetic of the reference lives in PyTorch (``nn.Conv2d`` / ``nn.BatchNorm2d`` /
``nn.ReLU`` / ``nn.Upsample``), which is a third-party dependency of the reference (pinned 1.2.0,
environment.yml:245; 2.11.0 in this image).  The restatement therefore calls the same
``torch.nn.functional`` primitives on the CPU; what is restated is the *wiring* of HRnet.py.

Pinned against the reference module itself by tests/test_oracle_vs_reference.py (build container)
and by the frozen fixture tests/golden/hrnet_w32_fwd.npz (everywhere).
"""
import zlib

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, HRnet.py:23,38
STAGE_MODULES = {2: 1, 3: 4, 4: 3}  # NUM_MODULES of STAGE2/3/4
BLOCKS_PER_BRANCH = 4
NUM_JOINTS = 17


def branch_channels(width):
    return [width, 2 * width, 4 * width, 8 * width]


# ----------------------------------------------------------------------------------------------
# state_dict schema
# ----------------------------------------------------------------------------------------------
def _bn(prefix, c):
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,)),
            (prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)),
            (prefix + ".num_batches_tracked", ())]


def hrnet_schema(width=32, joints=NUM_JOINTS):
    """Ordered list of (key, shape) for the reference module's state_dict (HRnet.py:288-337)."""
    ch = branch_channels(width)
    s = []
    s += [("conv1.weight", (64, 3, 3, 3))] + _bn("bn1", 64)          # HRnet.py:290-292
    s += [("conv2.weight", (64, 64, 3, 3))] + _bn("bn2", 64)         # HRnet.py:293-295
    for b in range(4):                                                # layer1, HRnet.py:297, 64-102
        p = f"layer1.{b}"
        cin = 64 if b == 0 else 256
        s += [(p + ".conv1.weight", (64, cin, 1, 1))] + _bn(p + ".bn1", 64)
        s += [(p + ".conv2.weight", (64, 64, 3, 3))] + _bn(p + ".bn2", 64)
        s += [(p + ".conv3.weight", (256, 64, 1, 1))] + _bn(p + ".bn3", 256)
        if b == 0:
            s += [(p + ".downsample.0.weight", (256, 64, 1, 1))] + _bn(p + ".downsample.1", 256)
    pre = [256]
    for stage in (2, 3, 4):
        nb = stage
        cur = ch[:nb]
        # transition, HRnet.py:341-380
        for i in range(nb):
            p = f"transition{stage - 1}.{i}"
            if i < len(pre):
                if cur[i] != pre[i]:
                    s += [(p + ".0.weight", (cur[i], pre[i], 3, 3))] + _bn(p + ".1", cur[i])
            else:
                for j in range(i + 1 - len(pre)):
                    cout = cur[i] if j == i - len(pre) else pre[-1]
                    s += [(f"{p}.{j}.0.weight", (cout, pre[-1], 3, 3))] + _bn(f"{p}.{j}.1", cout)
        # stage modules, HRnet.py:401-431, 105-266
        for m in range(STAGE_MODULES[stage]):
            mp = f"stage{stage}.{m}"
            for b in range(nb):
                for k in range(BLOCKS_PER_BRANCH):
                    bp = f"{mp}.branches.{b}.{k}"
                    s += [(bp + ".conv1.weight", (cur[b], cur[b], 3, 3))] + _bn(bp + ".bn1", cur[b])
                    s += [(bp + ".conv2.weight", (cur[b], cur[b], 3, 3))] + _bn(bp + ".bn2", cur[b])
            multi = not (stage == 4 and m == STAGE_MODULES[stage] - 1)  # HRnet.py:413-416
            for i in range(nb if multi else 1):
                for j in range(nb):
                    fp = f"{mp}.fuse_layers.{i}.{j}"
                    if j > i:                                           # HRnet.py:198-209
                        s += [(fp + ".0.weight", (cur[i], cur[j], 1, 1))] + _bn(fp + ".1", cur[i])
                    elif j < i:                                         # HRnet.py:212-240
                        for k in range(i - j):
                            cout = cur[i] if k == i - j - 1 else cur[j]
                            s += [(f"{fp}.{k}.0.weight", (cout, cur[j], 3, 3))] + _bn(f"{fp}.{k}.1", cout)
        pre = cur
    s += [("final_layer.weight", (joints, ch[0], 1, 1)), ("final_layer.bias", (joints,))]  # :331-337
    return s


def synth_state_dict(width=32, seed=0, joints=NUM_JOINTS):
    """Deterministic synthetic checkpoint: every tensor drawn from a generator seeded by its key.

    Independent of module construction order and of torch's RNG stream, so the reference module,
    the oracle and the CUDA implementation can all be handed bit-identical weights on any box.
    Conv weights ~ U(+-sqrt(3/fan_in)) (variance-preserving), BatchNorm affine and running statistics
    non-trivial so that BN folding is actually exercised.
    """
    sd = {}
    for key, shape in hrnet_schema(width, joints):
        rng = np.random.default_rng([seed, zlib.crc32(key.encode())])
        if key.endswith("num_batches_tracked"):
            t = torch.tensor(0, dtype=torch.int64)
        elif key.endswith("running_var"):
            t = torch.from_numpy(rng.uniform(0.6, 1.4, shape).astype(np.float32))
        elif key.endswith("running_mean"):
            t = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            bound = np.sqrt(3.0 / fan_in)
            if key == "final_layer.weight":
                bound /= 48.0  # keeps heatmaps at the reference's random-init scale (std ~0.16)
            t = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
        elif key == "final_layer.bias":
            t = torch.from_numpy(rng.normal(0.0, 0.05, shape).astype(np.float32))
        elif key.endswith(".bias"):
            t = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        else:  # BN weight (gamma)
            t = torch.from_numpy(rng.uniform(0.4, 0.9, shape).astype(np.float32))
        sd[key] = t
    return sd


def default_init_state_dict(width=32, seed=0, joints=NUM_JOINTS):
    """Checkpoint with torch's default init statistics (kaiming-uniform a=sqrt(5) convs; BN (1,0,0,1)).

    Same distribution as ``PoseHighResolutionNet()`` under ``torch.manual_seed``
    but keyed per tensor like synth_state_dict so it is reproducible without the reference.
    """
    sd = {}
    for key, shape in hrnet_schema(width, joints):
        rng = np.random.default_rng([seed, 1, zlib.crc32(key.encode())])
        if key.endswith("num_batches_tracked"):
            t = torch.tensor(0, dtype=torch.int64)
        elif key.endswith("running_var") or (key.endswith(".weight") and len(shape) == 1):
            t = torch.ones(shape)
        elif key.endswith("running_mean") or (key.endswith(".bias") and "final_layer" not in key):
            t = torch.zeros(shape)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / np.sqrt(fan_in)
            t = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
        else:  # final_layer.bias
            bound = 1.0 / np.sqrt(width)
            t = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
        sd[key] = t
    return sd


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
_TRAIN = False  # set by hrnet_forward_train: BatchNorm uses batch statistics and updates the running ones in place
_BF16 = False   # set by hrnet_forward_train(bf16_storage=True): tensors are rounded to bf16 where the device path stores them
_TRACE = None   # set by hrnet_forward_train(trace=[...]): one record per conv + BatchNorm unit (inputs, output)


def _q(t):
    """Round to bf16 and back (identity unless bf16 storage is emulated).  Autograd rounds the gradient that flows
    back through this point to bf16 as well, which is where the device path stores activation gradients."""
    return t.bfloat16().float() if _BF16 else t


def _qw(w):
    """bf16-rounded weights with a straight-through fp32 gradient (the device path keeps parameter gradients fp32)."""
    return w + (w.detach().bfloat16().float() - w.detach()) if _BF16 else w


def _conv_bn(sd, x, conv, bn, stride=1, relu=False, res=None):
    w = _qw(sd[conv + ".weight"])
    y = _q(F.conv2d(x, w, None, stride, w.shape[-1] // 2))
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"],
                     sd[bn + ".bias"], _TRAIN, 0.1, BN_EPS)
    if res is not None:
        y = y + res
    y = _q(F.relu(y) if relu else y)
    if _TRACE is not None:
        y.retain_grad()                 # after loss.backward(): y.grad = the gradient this unit receives
        _TRACE.append(dict(conv=conv, bn=bn, stride=stride, relu=relu, x=x.detach(),
                           res=None if res is None else res.detach(), y=y))
    return y


def _bottleneck(sd, x, p):
    """Bottleneck.forward, HRnet.py:82-102."""
    out = _conv_bn(sd, x, p + ".conv1", p + ".bn1", relu=True)
    out = _conv_bn(sd, out, p + ".conv2", p + ".bn2", relu=True)
    res = x
    if (p + ".downsample.0.weight") in sd:
        res = _conv_bn(sd, x, p + ".downsample.0", p + ".downsample.1")
    return _conv_bn(sd, out, p + ".conv3", p + ".bn3", relu=True, res=res)   # relu(bn3(conv3) + residual)


def _basic_block(sd, x, p):
    """BasicBlock.forward, HRnet.py:45-61."""
    out = _conv_bn(sd, x, p + ".conv1", p + ".bn1", relu=True)
    return _conv_bn(sd, out, p + ".conv2", p + ".bn2", relu=True, res=x)     # relu(bn2(conv2) + x)


def _hr_module(sd, xs, mp, n_out):
    """HighResolutionModule.forward, HRnet.py:248-266."""
    nb = len(xs)
    xs = list(xs)
    for b in range(nb):
        for k in range(BLOCKS_PER_BRANCH):
            xs[b] = _basic_block(sd, xs[b], f"{mp}.branches.{b}.{k}")
    outs = []
    for i in range(n_out):
        y = None
        for j in range(nb):
            fp = f"{mp}.fuse_layers.{i}.{j}"
            if j == i:
                t = xs[j]
            elif j > i:
                t = _conv_bn(sd, xs[j], fp + ".0", fp + ".1")
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _conv_bn(sd, t, f"{fp}.{k}.0", f"{fp}.{k}.1", stride=2, relu=(k != i - j - 1))
            y = t if y is None else y + t
        outs.append(_q(F.relu(y)))
    return outs


def _transition(sd, stage, ys, width):
    """_make_transition_layer + its use in forward, HRnet.py:341-380, 442-463."""
    nb = stage
    xs = []
    for i in range(nb):
        p = f"transition{stage - 1}.{i}"
        if i < len(ys):
            if (p + ".0.weight") in sd:
                xs.append(_conv_bn(sd, ys[i], p + ".0", p + ".1", relu=True))
            else:
                xs.append(ys[i])
        else:
            t = ys[-1]
            j = 0
            while f"{p}.{j}.0.weight" in sd:
                t = _conv_bn(sd, t, f"{p}.{j}.0", f"{p}.{j}.1", stride=2, relu=True)
                j += 1
            xs.append(t)
    return xs


def conv_bn_unit(sd, x, conv, bn, stride=1, relu=False, res=None, bf16_storage=True):
    """ONE conv + train-mode BatchNorm [+ residual] [+ ReLU] unit of the training graph (HRnet.py:48-59 under
    model.train()) on caller-supplied inputs, with live autograd.  Updates sd's running statistics of `bn` in place."""
    global _TRAIN, _BF16
    _TRAIN, _BF16 = True, bool(bf16_storage)
    try:
        return _conv_bn(sd, x, conv, bn, stride, relu, res)
    finally:
        _TRAIN, _BF16 = False, False


def hrnet_forward_train(sd, x, width=32, bf16_storage=False, trace=None):
    """PoseHighResolutionNet.forward under model.train() (02_train.py:153, 208): BatchNorm normalises with batch
    statistics (per replica, momentum 0.1 running-stat update in place on `sd`) and autograd is live, so
    ``loss.backward()`` fills ``.grad`` of every tensor of `sd` that requires grad.

    bf16_storage=True restates the SAME graph with every tensor rounded to bf16 at the points where the device path
    stores it (input, conv weights, raw conv outputs, block outputs and their gradients; arithmetic stays fp32).
    Batch-statistics BatchNorm over a few crops amplifies rounding differences chaotically with depth, so the fp32
    result is only a loose anchor for a bf16 pipeline in train mode; this variant is the tight one.

    trace: a list that receives one record per conv + BatchNorm unit in execution order (keys conv, bn, stride, relu,
    x, res, y; ``y.grad`` holds the unit's incoming gradient after ``loss.backward()``) - the inputs of the
    layer-by-layer device parity test."""
    global _TRAIN, _BF16, _TRACE
    _TRAIN, _BF16, _TRACE = True, bool(bf16_storage), trace
    try:
        return _forward(sd, x, width)
    finally:
        _TRAIN, _BF16, _TRACE = False, False, None


@torch.no_grad()
def hrnet_forward(sd, x, width=32):
    """PoseHighResolutionNet.forward in eval mode, HRnet.py:433-468.  x: f32 [B,3,H,W] -> [B,J,H/4,W/4]."""
    return _forward(sd, x, width)


def _forward(sd, x, width):
    x = _conv_bn(sd, _q(x), "conv1", "bn1", stride=2, relu=True)
    x = _conv_bn(sd, x, "conv2", "bn2", stride=2, relu=True)
    for b in range(4):
        x = _bottleneck(sd, x, f"layer1.{b}")
    ys = [x]
    for stage in (2, 3, 4):
        xs = _transition(sd, stage, ys, width)  # stage 2: both transition1 entries read x (HRnet.py:442-447)
        n_mod = STAGE_MODULES[stage]
        for m in range(n_mod):
            n_out = 1 if (stage == 4 and m == n_mod - 1) else stage
            xs = _hr_module(sd, xs, f"stage{stage}.{m}", n_out)
        ys = xs
    return F.conv2d(ys[0], _qw(sd["final_layer.weight"]), sd["final_layer.bias"])


def conv_flops_per_crop(width=32, image_hw=(256, 192), joints=NUM_JOINTS):
    """2 * MACs over the 293 executed convs, from the schema alone."""
    H, W = image_hw
    total = 0
    h4, w4 = H // 4, W // 4
    for key, shape in hrnet_schema(width, joints):
        if len(shape) != 4:
            continue
        cout, cin, k, _ = shape
        if key == "conv1.weight":
            oh, ow = H // 2, W // 2
        elif key == "conv2.weight" or key.startswith("layer1") or key == "final_layer.weight":
            oh, ow = h4, w4
        elif key.startswith("transition"):
            i = int(key.split(".")[1])
            oh, ow = h4 >> i, w4 >> i
        else:
            parts = key.split(".")
            if parts[2] == "branches":
                b = int(parts[3])
                oh, ow = h4 >> b, w4 >> b
            else:  # fuse_layers.i.j[.k]
                i, j = int(parts[3]), int(parts[4])
                if j > i:
                    oh, ow = h4 >> j, w4 >> j
                else:
                    kk = int(parts[5])
                    oh, ow = h4 >> (j + kk + 1), w4 >> (j + kk + 1)
        total += 2 * cout * cin * k * k * oh * ow
    return total
