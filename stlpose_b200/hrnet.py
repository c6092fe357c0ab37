"""Drop-in for the reference ``models.PoseHighResolutionNet`` (/root/reference/src/models/HRnet.py:275-499).

Same constructor call (``PoseHighResolutionNet(**kwargs)``), same ``forward(x)`` contract
(f32 ``[B,3,H,W]`` -> f32 ``[B,J,H/4,W/4]`` on the input's device) and a ``state_dict`` whose keys and shapes
are identical to the reference module (1 754 entries for W32), so upstream ``pose_hrnet_w32_256x192.pth``
checkpoints and the reference's own ``save_checkpoint`` files load with ``strict=True``.

The torch sub-modules below are parameter containers only; the arithmetic runs in libstlpose_b200.so
(tcgen05 implicit-GEMM convolutions with BatchNorm folded in).  There is no eager fallback.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib

BN_MOMENTUM = 0.1  # HRnet.py:23

DEFAULT_EXTRA = {  # upstream HRNet-W32 MODEL.EXTRA; W48 = width 48
    "STAGE2": {"NUM_MODULES": 1, "NUM_BRANCHES": 2, "NUM_BLOCKS": [4, 4]},
    "STAGE3": {"NUM_MODULES": 4, "NUM_BRANCHES": 3, "NUM_BLOCKS": [4, 4, 4]},
    "STAGE4": {"NUM_MODULES": 3, "NUM_BRANCHES": 4, "NUM_BLOCKS": [4, 4, 4, 4]},
}


def _conv_bn(cin, cout, k, stride, relu):
    mods = [nn.Conv2d(cin, cout, k, stride, k // 2, bias=False), nn.BatchNorm2d(cout, momentum=BN_MOMENTUM)]
    if relu:
        mods.append(nn.ReLU(inplace=True))  # parameter-free; keeps Sequential indices equal to the reference
    return nn.Sequential(*mods)


class _Params(nn.Module):
    """Named parameter holder (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; computation runs in libstlpose_b200.so")


def _basic_block(c):
    m = _Params()
    m.conv1 = nn.Conv2d(c, c, 3, 1, 1, bias=False)
    m.bn1 = nn.BatchNorm2d(c, momentum=BN_MOMENTUM)
    m.conv2 = nn.Conv2d(c, c, 3, 1, 1, bias=False)
    m.bn2 = nn.BatchNorm2d(c, momentum=BN_MOMENTUM)
    return m


def _bottleneck(cin, planes, with_downsample):
    m = _Params()
    m.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
    m.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
    m.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
    m.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
    m.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
    m.bn3 = nn.BatchNorm2d(planes * 4, momentum=BN_MOMENTUM)
    if with_downsample:
        m.downsample = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, bias=False),
                                     nn.BatchNorm2d(planes * 4, momentum=BN_MOMENTUM))
    return m


def _hr_module(channels, n_blocks, multi_scale_output):
    """Parameters of one HighResolutionModule (HRnet.py:105-243)."""
    nb = len(channels)
    m = _Params()
    m.branches = nn.ModuleList(
        [nn.Sequential(*[_basic_block(channels[b]) for _ in range(n_blocks[b])]) for b in range(nb)])
    rows = []
    for i in range(nb if multi_scale_output else 1):
        row = []
        for j in range(nb):
            if j > i:    # 1x1 conv + BN (+ nearest upsample, parameter-free)
                row.append(nn.Sequential(nn.Conv2d(channels[j], channels[i], 1, bias=False),
                                         nn.BatchNorm2d(channels[i])))
            elif j == i:
                row.append(None)
            else:        # chain of stride-2 3x3 convs; channels change only on the last hop
                hops = i - j
                row.append(nn.Sequential(*[
                    _conv_bn(channels[j], channels[i] if k == hops - 1 else channels[j], 3, 2, relu=k != hops - 1)
                    for k in range(hops)]))
        rows.append(nn.ModuleList(row))
    m.fuse_layers = nn.ModuleList(rows)
    return m


def _load_yaml_cfg():
    """Honour the reference's cwd-relative architecture YAML when it exists (HRnet.py:280-283)."""
    path = os.path.join("..", "resources", "HRnet", "cfg_hrnet_w32_256x192.yaml")
    if not os.path.isfile(path):
        return None
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


class PoseHighResolutionNet(nn.Module):
    """HRNet-W32/W48 pose network.  ``PoseHighResolutionNet()`` reads the reference's YAML if present, else W32.

    Extra keyword arguments (all optional, ignored by the reference): ``width`` (32/48), ``num_joints``,
    ``image_size=(H, W)`` used to size the first plan (the network stays fully convolutional: a plan is
    built per input resolution on demand).
    """

    def __init__(self, **kwargs):
        super().__init__()
        cfg = kwargs.get("cfg") or _load_yaml_cfg()
        extra = dict(DEFAULT_EXTRA)
        width = kwargs.get("width")
        joints = kwargs.get("num_joints", 17)
        if cfg is not None:
            model_cfg = cfg.get("MODEL", cfg)
            extra = model_cfg.get("EXTRA", extra)
            joints = model_cfg.get("NUM_JOINTS", joints)
            if width is None:
                width = extra["STAGE2"]["NUM_CHANNELS"][0]
        width = 32 if width is None else int(width)
        self.width = width
        self.num_joints = int(joints)
        self.stage_modules = [int(extra[f"STAGE{s}"]["NUM_MODULES"]) for s in (2, 3, 4)]
        blocks = {b for s in (2, 3, 4) for b in extra[f"STAGE{s}"]["NUM_BLOCKS"]}
        for s in (2, 3, 4):
            st = extra[f"STAGE{s}"]
            if st["NUM_BRANCHES"] != s or len(st["NUM_BLOCKS"]) != s:   # HRnet.py:125-138
                raise ValueError(f"NUM_BRANCHES({st['NUM_BRANCHES']}) <> NUM_BLOCKS({len(st['NUM_BLOCKS'])})")
            if "NUM_CHANNELS" in st and list(st["NUM_CHANNELS"]) != [width * 2 ** b for b in range(s)]:
                raise ValueError(f"NUM_CHANNELS of STAGE{s} must be width*(1,2,4,8)")
            if st.get("BLOCK", "BASIC") != "BASIC" or st.get("FUSE_METHOD", "SUM") != "SUM":
                raise ValueError("only BASIC blocks with SUM fusion are supported (upstream HRNet-W32/W48)")
        if len(blocks) != 1:
            raise ValueError("all branches must use the same NUM_BLOCKS")
        self.blocks = blocks.pop()
        if extra.get("FINAL_CONV_KERNEL", 1) != 1:
            raise ValueError("FINAL_CONV_KERNEL must be 1")
        ch = [width, 2 * width, 4 * width, 8 * width]

        # ---- parameter containers, named exactly as in the reference module ----------------------
        self.conv1 = nn.Conv2d(3, 64, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.layer1 = nn.Sequential(*[_bottleneck(64 if b == 0 else 256, 64, b == 0) for b in range(4)])
        transitions = {
            2: nn.ModuleList([_conv_bn(256, ch[0], 3, 1, True), nn.Sequential(_conv_bn(256, ch[1], 3, 2, True))]),
            3: nn.ModuleList([None, None, nn.Sequential(_conv_bn(ch[1], ch[2], 3, 2, True))]),
            4: nn.ModuleList([None, None, None, nn.Sequential(_conv_bn(ch[2], ch[3], 3, 2, True))]),
        }
        for s in (2, 3, 4):  # registration order = the reference's state_dict order
            setattr(self, f"transition{s - 1}", transitions[s])
            n_mod = self.stage_modules[s - 2]
            mods = [_hr_module(ch[:s], [self.blocks] * s, not (s == 4 and m == n_mod - 1)) for m in range(n_mod)]
            setattr(self, f"stage{s}", nn.Sequential(*mods))
        self.final_layer = nn.Conv2d(ch[0], self.num_joints, 1)
        self.pretrained_layers = extra.get("PRETRAINED_LAYERS", ["*"])

        # ---- device-side state ---------------------------------------------------------------------
        self._plans = {}          # (H, W) -> plan handle
        self._arena = None        # packed bf16 weights + fp32 biases (layout shared by all plans)
        self._buffer_generation = 0   # bumped whenever the arena or a workspace is (re)allocated: captured graphs hold
                                      # the old addresses and must be re-captured (KeypointPipeline checks it)
        self._arena_sig = None
        self._workspaces = {}     # (H, W) -> activation workspace (one per plan: each plan owns its zero cells)
        self._image_size = tuple(kwargs.get("image_size", (256, 192)))

    # -------------------------------------------------------------------------------------------------
    def load_pretrained(self, pretrained=""):
        """HRnet.py:470-499: re-initialise, then optionally load a checkpoint filtered by PRETRAINED_LAYERS."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if os.path.isfile(pretrained):
            sd = torch.load(pretrained, map_location="cpu")
            keep = {k: v for k, v in sd.items()
                    if k.split(".")[0] in self.pretrained_layers or self.pretrained_layers[0] == "*"}
            self.load_state_dict(keep, strict=False)
        elif pretrained:
            raise ValueError("{} is not exist!".format(pretrained))
        self.invalidate_packed_weights()
        return self

    def invalidate_packed_weights(self):
        """Call after mutating parameters in place (e.g. an optimizer step) so the next forward re-packs them."""
        self._arena_sig = None

    def load_state_dict(self, *args, **kwargs):
        r = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed_weights()
        return r

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        self._arena = None
        self._workspaces = {}
        self._buffer_generation += 1
        self.invalidate_packed_weights()
        return r

    # -------------------------------------------------------------------------------------------------
    def _plan(self, h, w):
        key = (h, w)
        if key not in self._plans:
            if h % 32 or w % 32:
                raise ValueError(f"input size must be a multiple of 32, got {h}x{w}")
            cfg = _lib.HrnetCfg(self.width, self.num_joints, (ctypes.c_int * 3)(*self.stage_modules), self.blocks, h, w)
            handle = _lib.lib().stl_plan_create(ctypes.byref(cfg))
            if not handle:
                raise _lib.StlError(_lib.lib().stl_last_error().decode())
            self._plans[key] = ctypes.c_void_p(handle)
        return self._plans[key]

    def _signature(self):
        w = self.conv1.weight
        return (w.device, w.data_ptr(), sum(p._version for p in self.parameters()))

    def _ensure_packed(self, plan):
        L = _lib.lib()
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise _lib.StlError("PoseHighResolutionNet runs on CUDA only: move the module with .to('cuda') "
                                "(there is no CPU fallback)")
        sig = self._signature()
        if self._arena is not None and self._arena_sig == sig:
            return
        nbytes = L.stl_plan_weight_bytes(plan)
        if self._arena is None or self._arena.numel() != nbytes or self._arena.device != dev:
            self._arena = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self._buffer_generation += 1
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        info = _lib.ConvInfo()
        stream = _lib.current_stream()

        def f32(t):
            t = t.detach()
            return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()

        keep = []
        for i in range(L.stl_plan_num_convs(plan)):
            _lib.check(L.stl_plan_conv_info(plan, i, ctypes.byref(info)))
            ck, bk = info.conv_key.decode(), info.bn_key.decode()
            w = f32(sd[ck + ".weight"])
            if bk:
                g, b = f32(sd[bk + ".weight"]), f32(sd[bk + ".bias"])
                m, v = f32(sd[bk + ".running_mean"]), f32(sd[bk + ".running_var"])
                eps = 1e-5
                cb = None
            else:
                g = b = m = v = None
                eps = 0.0
                cb = f32(sd[ck + ".bias"])
            keep += [w, g, b, m, v, cb]
            _lib.check(L.stl_plan_pack_conv(plan, i, _lib.ptr(w), _lib.ptr(g), _lib.ptr(b), _lib.ptr(m), _lib.ptr(v),
                                            _lib.ptr(cb), eps, _lib.ptr(self._arena), stream))
        self._arena_sig = sig

    def _run(self, x, flip_pair):
        if self.training:
            if flip_pair:
                raise NotImplementedError("the flip test is an evaluation feature; call .eval() first")
            from .training import train_forward
            out = train_forward(self, x)
            self.invalidate_packed_weights()     # running statistics moved: the folded eval weights are stale
            return out
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected input [B,3,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise _lib.StlError("input must be a CUDA tensor (there is no CPU fallback)")
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        L = _lib.lib()
        B, _, H, W = x.shape
        n_img = 2 * B if flip_pair else B
        heat = torch.empty((n_img, self.num_joints, H // 4, W // 4), dtype=torch.float32, device=x.device)
        if B == 0:
            return heat
        with torch.cuda.device(x.device):
            plan = self._plan(H, W)
            self._ensure_packed(plan)
            need = L.stl_plan_workspace_bytes(plan, n_img)
            ws = self._workspaces.get((H, W))
            if ws is None or ws.numel() < need or ws.device != x.device:
                self._workspaces.pop((H, W), None)
                ws = None
                ws = self._workspaces[(H, W)] = torch.empty(need, dtype=torch.uint8, device=x.device)
                self._buffer_generation += 1
            _lib.check(L.stl_plan_forward(plan, _lib.ptr(x), B, int(flip_pair), _lib.ptr(heat), _lib.ptr(self._arena),
                                          _lib.ptr(ws), ws.numel(), _lib.current_stream()))
        return heat

    def profile_ops(self, x, flip_pair=False):
        """Measurement aid: one forward with every launch bracketed by CUDA events.

        Returns a list of dicts (one per launch): kind, conv key, shapes, algorithmic flops/bytes for this batch,
        launch shape and the measured duration in ms."""
        L = _lib.lib()
        x = x.detach().float().contiguous()
        B, _, H, W = x.shape
        n_img = 2 * B if flip_pair else B
        heat = torch.empty((n_img, self.num_joints, H // 4, W // 4), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            plan = self._plan(H, W)
            self._ensure_packed(plan)
            need = L.stl_plan_workspace_bytes(plan, n_img)
            ws = self._workspaces.get((H, W))
            if ws is None or ws.numel() < need:
                ws = self._workspaces[(H, W)] = torch.empty(need, dtype=torch.uint8, device=x.device)
                self._buffer_generation += 1
            n_ops = L.stl_plan_launches_per_forward(plan)
            ms = (ctypes.c_float * n_ops)()
            _lib.check(L.stl_plan_forward_timed(plan, _lib.ptr(x), B, int(flip_pair), _lib.ptr(heat),
                                                _lib.ptr(self._arena), _lib.ptr(ws), ws.numel(),
                                                _lib.current_stream(), ms))
        out = []
        info, cinfo = _lib.OpInfo(), _lib.ConvInfo()
        for i in range(n_ops):
            _lib.check(L.stl_plan_op_info(plan, i, ctypes.byref(info)))
            d = {n: getattr(info, n) for n, _ in _lib.OpInfo._fields_}
            d["kind"] = ("stem", "conv_tc", "fuse_sum", "block_tc", "link_tc")[info.kind]
            d["key"] = ""
            if info.layer >= 0:
                _lib.check(L.stl_plan_conv_info(plan, info.layer, ctypes.byref(cinfo)))
                d["key"] = cinfo.conv_key.decode()
            d["flops"] = info.flops_per_image * n_img
            d["bytes"] = info.bytes_per_image * n_img
            d["ms"] = ms[i]
            out.append(d)
        return out

    def forward(self, x):
        """HRnet.py:433-468 (eval mode)."""
        return self._run(x, flip_pair=False)

    def forward_flip_pair(self, x):
        """Both flip-test passes as one 2B batch: rows [0,B) = model(x), rows [B,2B) = model(x.flip(3)) raw."""
        return self._run(x, flip_pair=True)

    def launches_per_forward(self, h=None, w=None):
        """Kernels one forward enqueues (for the plan's current batch binding; ops of a sub-batched branch group
        count once per sub-batch)."""
        h, w = (h, w) if h else self._image_size
        return _lib.lib().stl_plan_kernel_launches(self._plan(h, w))

    def __del__(self):
        try:
            for p in self._plans.values():
                _lib.lib().stl_plan_destroy(p)
        except Exception:
            pass
