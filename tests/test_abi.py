"""CPU-side checks: the C-ABI library loads, exports every declared symbol, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "stlpose_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from stlpose_b200 import _lib
    handle = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/stlpose_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == declared, "ctypes prototypes out of sync with the header"
    assert handle.stl_abi_version() == 1


def test_plan_topology_without_gpu():
    from oracle import hrnet_oracle
    from stlpose_b200 import _lib
    import stlpose_b200 as S
    L = _lib.lib()
    for width, hw in ((32, (256, 192)), (48, (384, 288))):
        m = S.PoseHighResolutionNet(width=width, image_size=hw)
        plan = m._plan(*hw)
        n = L.stl_plan_num_convs(plan)
        assert n == 293                                   # 293 convs
        # + fuse-sum per HRModule + input packing; the 32 BasicBlocks of W32's 32-channel branch are one launch each; in
        # layer1, downsample + conv3 of block 0 + conv1 of block 1 are one launch and so are conv3 + the next conv1 of
        # blocks 1 and 2; W32's last fuse row and the heatmap head are one launch
        assert L.stl_plan_launches_per_forward(plan) == 293 + 8 + 1 - 4 - (33 if width == 32 else 0)
        schema = dict(hrnet_oracle.hrnet_schema(width))
        info = _lib.ConvInfo()
        seen = set()
        flops = 0
        for i in range(n):
            _lib.check(L.stl_plan_conv_info(plan, i, ctypes.byref(info)))
            ck, bk = info.conv_key.decode(), info.bn_key.decode()
            assert schema[ck + ".weight"] == (info.cout, info.cin, info.ksize, info.ksize)
            assert (bk + ".running_var") in schema if bk else (ck + ".bias") in schema
            seen.add(ck)
        assert len(seen) == n                              # every conv of the reference appears exactly once
        assert L.stl_plan_workspace_bytes(plan, 2) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    import numpy as np
    import stlpose_b200 as S
    from stlpose_b200 import _lib
    L = _lib.lib()
    assert L.stl_flip_avg(None, None, None, 1, 17, 64, 48, None, 0, None) != 0
    assert b"no CUDA device" in L.stl_last_error()
    with pytest.raises(Exception):
        S.get_max_preds_hrnet(np.zeros((1, 17, 64, 48), np.float32))


@pytest.mark.parametrize("env,extra", [({}, 0), ({"STLPOSE_FUSE_LINK": "0"}, 3), ({"STLPOSE_FUSE_DOWNSAMPLE": "0"}, 2),
                                       ({"STLPOSE_FUSE_HEAD": "0"}, 1), ({"STLPOSE_FUSE_BLOCK": "0"}, 32),
                                       ({"STLPOSE_FUSE_LINK": "0", "STLPOSE_FUSE_DOWNSAMPLE": "0", "STLPOSE_FUSE_HEAD": "0",
                                         "STLPOSE_FUSE_BLOCK": "0"}, 37)])
def test_plan_fusion_switches_keep_every_conv(env, extra, monkeypatch):
    """Every fusion of the plan (BasicBlock kernel, conv3 + downsample as one two-input convolution, Bottleneck junctions,
    last fuse row + head) can be switched off by its environment variable; whatever the combination, each of the 293
    convolutions of the reference's state_dict is packed exactly once, and the launch count moves by the documented
    amount (no GPU needed: the plan is host-side)."""
    from stlpose_b200 import _lib
    import stlpose_b200 as S
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    L = _lib.lib()
    m = S.PoseHighResolutionNet(width=32, image_size=(256, 192))
    plan = m._plan(256, 192)
    n = L.stl_plan_num_convs(plan)
    info = _lib.ConvInfo()
    keys = set()
    for i in range(n):
        _lib.check(L.stl_plan_conv_info(plan, i, ctypes.byref(info)))
        keys.add(info.conv_key.decode())
    assert n == 293 and len(keys) == 293
    assert L.stl_plan_launches_per_forward(plan) == 293 + 8 + 1 - 4 - 33 + extra


def test_cooperative_batchnorm_policy():
    """training._use_coop_bn ("auto"): the cooperative single-launch BatchNorm is used for the backward direction at
    every size and never for the forward direction, and is off in multi-process jobs (checked by the
    world-size-2 tests of test_parallel_cpu.py through the same function)."""
    from stlpose_b200 import training
    if training.COOP_BN != "auto":
        pytest.skip("STLPOSE_TRAIN_COOP_BN is forced")
    small = torch.empty(4, 9, 7, 32, dtype=torch.bfloat16)
    big = torch.empty(0, dtype=torch.bfloat16).new_empty((64, 65, 49, 256))
    assert training._use_coop_bn(small, backward=True) and training._use_coop_bn(big, backward=True)
    assert not training._use_coop_bn(small) and not training._use_coop_bn(big)
