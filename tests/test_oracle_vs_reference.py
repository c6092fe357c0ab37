"""Oracle vs the live reference code (only where /root/reference exists, i.e. the build container)."""
import numpy as np
import pytest
import torch

from oracle import hrnet_oracle, pose_oracle, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted")


def test_schema_matches_reference_state_dict():
    for width, hw in ((32, (256, 192)), (48, (384, 288))):
        m = ref_shim.build_reference_hrnet(width, hw)
        sd = m.state_dict()
        schema = hrnet_oracle.hrnet_schema(width)
        assert [k for k, _ in schema] == list(sd.keys())
        assert all(tuple(sd[k].shape) == s for k, s in schema)
        m.load_state_dict(hrnet_oracle.synth_state_dict(width), strict=True)


def test_forward_matches_reference_module():
    m = ref_shim.build_reference_hrnet(32).eval()
    sd = hrnet_oracle.default_init_state_dict(32, seed=3)
    m.load_state_dict(sd, strict=True)
    x = torch.randn(1, 3, 256, 192, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        y_ref = m(x)
    assert (hrnet_oracle.hrnet_forward(sd, x, 32) - y_ref).abs().max().item() < 1e-5


def test_decode_matches_reference_random():
    L = ref_shim.lib()
    rng = np.random.default_rng(0)
    for (h, w) in ((64, 48), (96, 72), (16, 12)):
        hm = rng.standard_normal((5, 17, h, w)).astype(np.float32)
        hm[0, 0] = -np.abs(hm[0, 0])      # all-negative map
        hm[1, 1] = 0.0                     # all ties at zero
        c, s = pose_oracle.synth_boxes(5, seed=h)
        ref = L.pose_parsing.get_final_preds_hrnet(hm.copy(), c, s)
        got = pose_oracle.get_final_preds(hm, c, s)
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[1], ref[1])
        assert np.abs(got[0] - ref[0]).max() < 1e-3


def test_train_step_matches_reference_module():
    """model.train() forward, PersonMSELoss-style loss, backward: heatmaps, every parameter gradient and the updated
    running statistics of the oracle equal the reference module's (02_train.py:203-218)."""
    m = ref_shim.build_reference_hrnet(32).train()
    sd0 = hrnet_oracle.synth_state_dict(32, seed=0)
    m.load_state_dict(sd0, strict=True)
    x = torch.randn(2, 3, 128, 96, generator=torch.Generator().manual_seed(7))
    tgt = torch.from_numpy(pose_oracle.blob_heatmaps(2, 17, 32, 24, seed=1, noise=0.0))
    y_ref = m(x)
    (0.5 * ((y_ref - tgt) ** 2).mean()).backward()
    sd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone())
          for k, v in sd0.items()}
    y = hrnet_oracle.hrnet_forward_train(sd, x, 32)
    (0.5 * ((y - tgt) ** 2).mean()).backward()
    assert (y - y_ref).abs().max().item() < 1e-5
    ref_sd = m.state_dict()
    for name, p in m.named_parameters():
        assert (sd[name].grad - p.grad).abs().max().item() <= 1e-5 * max(1.0, p.grad.abs().max().item()), name
    for k, v in ref_sd.items():
        if "running" in k:
            assert (sd[k] - v).abs().max().item() < 1e-5, k


def test_pck_helpers_match_reference_source():
    M = ref_shim.metrics_functions()
    rng = np.random.default_rng(3)
    for _ in range(5):
        pred = rng.integers(0, 48, size=(7, 17, 2)).astype(np.float32)
        tgt = rng.integers(0, 48, size=(7, 17, 2)).astype(np.float32)
        tgt[rng.random((7, 17)) < 0.3] = 0
        norm = np.ones((7, 2)) * np.array([64, 48]) / 10
        d_ref, d = M.calc_dists(pred, tgt, norm), pose_oracle.calc_dists(pred, tgt, norm)
        assert np.allclose(d, d_ref, atol=1e-12)
        for j in range(17):
            assert pose_oracle.dist_acc(d[j]) == M.dist_acc(d_ref[j])


def test_affine_solve_and_warp_match_cv2_bit_exact():
    import cv2
    L = ref_shim.lib()
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(120, 160, 3), dtype=np.uint8)
    for t in range(40):
        c = rng.uniform(10, 150, 2).astype(np.float32)
        s = rng.uniform(0.1, 1.2, 2).astype(np.float32)
        rot = float(rng.uniform(-80, 80)) if t % 2 else 0
        m_ref = L.transforms.get_affine_transform(c, s, rot, np.array([48, 64]))
        m = pose_oracle.forward_affine(c, s, rot, (48, 64))
        assert np.array_equal(m, m_ref)
        assert np.array_equal(pose_oracle.warp_affine_u8(img, m, (48, 64)),
                              cv2.warpAffine(img, m_ref, (48, 64), flags=cv2.INTER_LINEAR))
        if t < 12:                                          # float32 images: float interpolation, same source coordinates
            imgf = rng.standard_normal(img.shape).astype(np.float32)
            assert np.array_equal(pose_oracle.warp_affine_f32(imgf, m, (48, 64)),
                                  cv2.warpAffine(imgf, m_ref, (48, 64), flags=cv2.INTER_LINEAR))
    td = L.transforms.TransformDetection()
    for box in ([5, 5, 60, 100], [-30, 20, 90, 80], [100, 60, 200, 140]):
        rc, rs = td._coords2cs(box)
        oc, os_ = pose_oracle.coords2cs(box)
        assert np.array_equal(rc, oc) and np.array_equal(rs, os_)
