"""PoseHighResolutionNet.load_pretrained against the reference method
(/root/reference/src/models/HRnet.py:470-499): same re-initialisation draw for draw, same PRETRAINED_LAYERS filter,
same error for a missing checkpoint.  Host-side logic only: runs on the CPU."""
import os

import pytest
import torch

from oracle import hrnet_oracle, ref_shim


def _ours(width=32):
    import stlpose_b200 as S
    return S.PoseHighResolutionNet(width=width)


def test_reinitialisation_statistics_and_error():
    m = _ours()
    torch.manual_seed(3)
    assert m.load_pretrained() is m                                  # HRnet.py:499 returns self
    sd = m.state_dict()
    w = sd["stage3.2.branches.1.2.conv1.weight"]
    assert abs(w.std().item() - 1e-3) < 1e-4 and abs(w.mean().item()) < 1e-4        # normal_(std=0.001), HRnet.py:474
    assert torch.all(sd["final_layer.bias"] == 0)
    assert torch.all(sd["bn1.weight"] == 1) and torch.all(sd["bn1.bias"] == 0)       # HRnet.py:478-480
    with pytest.raises(ValueError, match="is not exist!"):                            # HRnet.py:496-497
        m.load_pretrained("/nonexistent/checkpoint.pth")


def test_checkpoint_filter(tmp_path):
    """Keys whose first component is in PRETRAINED_LAYERS are loaded (strict=False), everything else keeps the fresh
    initialisation; '*' loads everything (HRnet.py:486-494)."""
    ck = hrnet_oracle.synth_state_dict(32, seed=5)
    path = os.path.join(tmp_path, "ck.pth")
    torch.save(ck, path)
    m = _ours()
    m.pretrained_layers = ["conv1", "bn1", "layer1"]
    torch.manual_seed(0)
    m.load_pretrained(path)
    sd = m.state_dict()
    for k, v in sd.items():
        if k.split(".")[0] in ("conv1", "bn1", "layer1"):
            assert torch.equal(v, ck[k]), k
    assert not torch.equal(sd["conv2.weight"], ck["conv2.weight"])
    assert abs(sd["stage2.0.branches.0.0.conv1.weight"].std().item() - 1e-3) < 2e-4
    m.pretrained_layers = ["*"]
    m.load_pretrained(path)
    assert all(torch.equal(v, ck[k]) for k, v in m.state_dict().items())


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not available")
@pytest.mark.parametrize("layers", [["*"], ["conv1", "bn1", "conv2", "bn2", "layer1", "transition1", "stage2"]])
def test_matches_reference_method(tmp_path, layers):
    """Same seed -> the reference's load_pretrained and ours leave identical state_dicts (the module iteration order, and
    with it the order of the random draws, is part of the drop-in contract), with and without a checkpoint."""
    ref = ref_shim.build_reference_hrnet(32, (256, 192))
    m = _ours()
    ck = hrnet_oracle.synth_state_dict(32, seed=7)
    path = os.path.join(tmp_path, "ck.pth")
    torch.save(ck, path)
    for pretrained in ("", path):
        ref.pretrained_layers = list(layers)
        m.pretrained_layers = list(layers)
        torch.manual_seed(11)
        ref.load_pretrained(pretrained)
        torch.manual_seed(11)
        m.load_pretrained(pretrained)
        a, b = ref.state_dict(), m.state_dict()
        assert list(a.keys()) == list(b.keys())
        bad = [k for k in a if not torch.equal(a[k], b[k])]
        assert not bad, bad[:3]
