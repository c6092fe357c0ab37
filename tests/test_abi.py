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
