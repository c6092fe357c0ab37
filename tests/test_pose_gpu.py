"""GPU parity: decode / flip-average / loss kernels (through the C ABI) vs the CPU oracle and the golden fixtures."""
import numpy as np
import pytest
import torch

from oracle import pose_oracle

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_decode_matches_reference_goldens(golden, golden_inputs):
    import stlpose_b200 as S
    g = golden("decode.npz")
    cases = {
        "rand64": (golden_inputs["hm_rand_64x48"], golden_inputs["center_a"], golden_inputs["scale_a"]),
        "rand96": (golden_inputs["hm_rand_96x72"], golden_inputs["center_b"], golden_inputs["scale_b"]),
        "blobs": (g["hm_blobs_f16"].astype(np.float32),) + pose_oracle.synth_boxes(2, seed=9),
    }
    for tag, (hm, c, s) in cases.items():
        p0, m0 = S.get_max_preds_hrnet(hm)
        assert np.array_equal(p0, g[f"{tag}_max_preds"]) and np.array_equal(m0, g[f"{tag}_max_vals"])
        preds, maxvals, coords = S.get_final_preds_hrnet(hm, c, s)
        assert preds.dtype == np.float32 and preds.shape == (hm.shape[0], 17, 2)
        assert np.array_equal(coords, g[f"{tag}_coords"])       # bit-exact indices and quarter-pixel offsets
        assert np.array_equal(maxvals, g[f"{tag}_maxvals"])     # bit-exact max values
        assert np.abs(preds - g[f"{tag}_preds"]).max() < 1e-3   # px; closed-form affine vs cv2
        # CUDA-tensor input, tensor output
        p2, m2, c2 = S.get_final_preds_hrnet(_cuda(hm), _cuda(c), _cuda(s), as_tensor=True)
        assert p2.is_cuda and np.array_equal(c2.cpu().numpy(), coords)


def test_decode_edge_cases():
    import stlpose_b200 as S
    rng = np.random.default_rng(5)
    for (h, w) in ((64, 48), (96, 72), (16, 12), (9, 7)):     # 9x7: width not a multiple of 4 (scalar path)
        hm = rng.standard_normal((6, 17, h, w)).astype(np.float32)
        hm[0, 0] = -np.abs(hm[0, 0])          # all negative -> coords (0,0)
        hm[1, 1] = 0.0                        # all ties at 0 -> idx 0, masked
        hm[2, 2] = 1.0                        # all ties positive -> first index
        hm[3, 3, 0, 0] = 50.0                 # corner peaks: no refinement
        hm[3, 4, h - 1, w - 1] = 50.0
        hm[3, 5, 1, 1] = 50.0                 # px = 1 is outside the refinement window (1 < px)
        hm[3, 6, 2, 2] = 50.0                 # first refined position
        hm[3, 7, h - 2, w - 2] = 50.0         # last refined position
        hm[4, 8, 3, 3] = 7.0; hm[4, 8, 3, 4] = 7.0   # tie between neighbours -> lower index
        c, s = pose_oracle.synth_boxes(6, seed=h)
        ref = pose_oracle.get_final_preds(hm, c, s)
        got = S.get_final_preds_hrnet(hm, c, s)
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[1], ref[1])
        assert np.abs(got[0] - ref[0]).max() < 1e-3
    assert S.get_max_preds_hrnet(np.zeros((0, 17, 64, 48), np.float32)) == ([], [])


def test_flip_average_bit_exact(golden, golden_inputs):
    import stlpose_b200 as S
    from stlpose_b200 import _lib
    from stlpose_b200.transforms import _pairs_array
    g = golden("flip.npz")
    fb = S.flip_back(golden_inputs["flip_out_f"], S.FLIP_PAIRS)
    assert not fb.is_cuda and np.array_equal(fb.numpy(), g["flip_back"])
    a, f = _cuda(golden_inputs["flip_out"]), _cuda(golden_inputs["flip_out_f"])
    out = torch.empty_like(a)
    pairs, n = _pairs_array(S.FLIP_PAIRS)
    B, J, h, w = a.shape
    _lib.check(_lib.lib().stl_flip_avg(_lib.ptr(a), _lib.ptr(f), _lib.ptr(out), B, J, h, w, pairs, n,
                                       _lib.current_stream()))
    assert np.array_equal(out.cpu().numpy(), g["avg"])


def test_fused_flip_decode_equals_two_step(golden_inputs):
    from stlpose_b200.pose_parsing import _decode
    import stlpose_b200 as S
    a, f = golden_inputs["flip_out"], golden_inputs["flip_out_f"]
    c, s = pose_oracle.synth_boxes(2, seed=1)
    avg_ref = pose_oracle.flip_average(a, f)
    ref = pose_oracle.get_final_preds(avg_ref, c, s)
    preds, maxvals, coords, avg = _decode(a, c, s, True, heat_flipped=f, pairs=S.FLIP_PAIRS, want_avg=True)
    assert np.array_equal(avg.cpu().numpy(), avg_ref)
    assert np.array_equal(coords.cpu().numpy(), ref[2]) and np.array_equal(maxvals.cpu().numpy(), ref[1])
    assert np.abs(preds.cpu().numpy() - ref[0]).max() < 1e-3


def test_loss_forward_backward(golden, golden_inputs):
    import stlpose_b200 as S
    g = golden("loss.npz")
    o = _cuda(golden_inputs["loss_out"]).requires_grad_(True)
    loss = S.PersonMSELoss()(o, _cuda(golden_inputs["loss_tgt"]), _cuda(golden_inputs["loss_tw"]))
    assert loss.dim() == 0
    (loss * 3.0).backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))
    assert np.abs(o.grad.cpu().numpy() / 3.0 - g["grad"]).max() < 1e-9 + 1e-6 * np.abs(g["grad"]).max()
    with pytest.raises(TypeError):
        S.PersonMSELoss()(o, o)


@pytest.mark.parametrize("batch", [1024, 16384, 65536])
def test_decode_full_size_properties(batch):
    """BASELINE config 5 sizes: size-independent properties instead of a CPU oracle pass."""
    import stlpose_b200 as S
    gen = torch.Generator(device="cuda").manual_seed(batch)
    hm = torch.randn(batch, 17, 64, 48, device="cuda", generator=gen)
    coords, maxvals = S.get_max_preds_hrnet(hm, as_tensor=True)
    flat = hm.view(batch, 17, -1)
    mx, idx = flat.max(dim=2)          # torch's own reduction as an independent check of value and location
    assert torch.equal(maxvals[..., 0], mx)
    gathered = flat.gather(2, (coords[..., 1] * 48 + coords[..., 0]).long().unsqueeze(-1))[..., 0]
    assert torch.equal(gathered, mx)
    # scaling by a power of two is exact in fp32 (no new ties): the argmax must not move and the max scales with it
    c2, m2 = S.get_max_preds_hrnet(hm * 4.0, as_tensor=True)
    pos = maxvals[..., 0] > 0                      # coords are zeroed where the max is <= 0, on both sides alike
    assert torch.equal(c2, coords) and torch.equal(m2, maxvals * 4.0) and bool(pos.any())


def test_decode_96x72_full_size_properties():
    """BASELINE config 5, 96x72 maps: flip-fused decode == decode of the separately averaged maps (bit-exact)."""
    import stlpose_b200 as S
    from stlpose_b200 import _lib
    from stlpose_b200.pose_parsing import _decode
    from stlpose_b200.transforms import _pairs_array
    B = 4096
    gen = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(B, 17, 96, 72, device="cuda", generator=gen)
    f = torch.randn(B, 17, 96, 72, device="cuda", generator=gen)
    avg = torch.empty_like(a)
    pairs, n = _pairs_array(S.FLIP_PAIRS)
    _lib.check(_lib.lib().stl_flip_avg(_lib.ptr(a), _lib.ptr(f), _lib.ptr(avg), B, 17, 96, 72, pairs, n,
                                       _lib.current_stream()))
    _, m1, c1, _ = _decode(avg, None, None, True)
    _, m2, c2, avg2 = _decode(a, None, None, True, heat_flipped=f, pairs=S.FLIP_PAIRS, want_avg=True)
    assert torch.equal(avg, avg2) and torch.equal(m1, m2) and torch.equal(c1, c2)
    # linearity of the loss gradient in (out - tgt) and its closed form at full size
    tw = torch.ones(B, 17, 1, device="cuda")
    o = a.clone().requires_grad_(True)
    loss = S.PersonMSELoss()(o, f, tw)
    loss.backward()
    denom = 17 * B * 96 * 72
    assert torch.allclose(o.grad, (a - f) / denom, rtol=1e-5, atol=1e-12)
    assert abs(loss.item() - 0.5 * ((a - f).double() ** 2).sum().item() / denom) < 1e-6 * loss.item()


def test_pck_accuracy_matches_oracle_and_fixture(golden):
    """stlpose_b200.metrics.accuracy (device decode x2 + PCK kernel) vs the oracle and the reference fixture."""
    from oracle.make_golden import pck_inputs
    from stlpose_b200 import metrics
    out, tgt = pck_inputs()
    acc, avg, cnt, pred = metrics.accuracy(out, tgt)
    o_acc, o_avg, o_cnt, o_pred = pose_oracle.accuracy(out, tgt)
    assert np.array_equal(pred, o_pred) and cnt == o_cnt
    assert np.abs(acc - o_acc).max() < 1e-6 and abs(avg - o_avg) < 1e-6
    assert np.abs(acc[1:] - golden("pck.npz")["per_joint"]).max() < 1e-6
    # CUDA tensors in, tensors out (no host copy), other threshold, larger random batch
    rng = np.random.default_rng(5)
    big_o = rng.standard_normal((300, 17, 64, 48)).astype(np.float32)
    big_t = np.where(rng.random((300, 17, 1, 1)) < 0.25, 0, rng.standard_normal((300, 17, 64, 48))).astype(np.float32)
    a_t, avg_t, cnt_t, _ = metrics.accuracy(torch.from_numpy(big_o).cuda(), torch.from_numpy(big_t).cuda(), thr=3.0,
                                            as_tensor=True)
    r_acc, r_avg, r_cnt, _ = pose_oracle.accuracy(big_o, big_t, thr=3.0)
    assert a_t.is_cuda and int(cnt_t) == r_cnt and np.abs(a_t.cpu().numpy() - r_acc).max() < 1e-6
    # nothing labeled: every joint -1, average 0, cnt 0 (metrics.py:359-362)
    acc0, avg0, cnt0, _ = metrics.accuracy(out, np.zeros_like(tgt))
    assert cnt0 == 0 and avg0 == 0 and (acc0[1:] == -1).all() and acc0[0] == 0


def test_crop_extraction_bit_exact(golden):
    """stlpose_b200.transforms.TransformDetection / crop (device warp) vs the reference fixture and the oracle."""
    from oracle.make_golden import crop_inputs
    from stlpose_b200 import transforms as T
    g = golden("crops.npz")
    img, boxes = crop_inputs()
    td = T.TransformDetection()
    dets, centers, scales = td(img, boxes)
    assert dets.dtype == np.uint8 and dets.shape == (5, 3, 256, 192)
    assert np.array_equal(centers, g["centers"]) and np.array_equal(scales, g["scales"])
    assert np.array_equal(dets, g["dets"])
    assert np.array_equal(T.crop(img, g["centers"][1], g["scales"][1], np.array([192, 256]), rot=30), g["rot30"])
    # fused ToTensor + Normalize: same float32 operation order as torchvision on the reference's uint8 crops
    x, _, _ = td.extract_normalized(img, boxes)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    want = (torch.from_numpy(g["dets"]).float().div(255) - mean) / std
    assert x.is_cuda and x.dtype == torch.float32 and torch.equal(x.cpu(), want)
    # other geometry, random rotations, CUDA image input, vs the oracle
    rng = np.random.default_rng(3)
    img2 = rng.integers(0, 256, size=(97, 131, 3), dtype=np.uint8)
    mats = [pose_oracle.forward_affine(rng.uniform(0, 130, 2).astype(np.float32), rng.uniform(0.1, 1.0, 2).astype(np.float32),
                                       float(rng.uniform(-90, 90)), (72, 96)) for _ in range(7)]
    got = T.warp_affine_crops(torch.from_numpy(img2).cuda(), mats, (72, 96), as_tensor=True).cpu().numpy()
    for i, m in enumerate(mats):
        assert np.array_equal(got[i].transpose(1, 2, 0), pose_oracle.warp_affine_u8(img2, m, (72, 96))), i
    e_d, e_c, e_s = td(img, [])
    assert len(e_d) == 0 and len(e_c) == 0 and len(e_s) == 0
    # float32 images (04_evaluate_vases_qualitatively.py:209-213): bit-identical to cv2's float interpolation
    gf = golden("crops_f32.npz")
    imgf = (img.astype(np.float32) / np.float32(255)).astype(np.float16).astype(np.float32)
    dets_f, _, _ = td(imgf, boxes[1:2])
    assert dets_f.dtype == np.float32 and np.array_equal(dets_f, gf["dets"])
    img3 = rng.standard_normal((97, 131, 3)).astype(np.float32)
    got_f = T.warp_affine_crops(torch.from_numpy(img3).cuda(), mats, (72, 96), as_tensor=True).cpu().numpy()
    for i, m in enumerate(mats):
        assert np.array_equal(got_f[i].transpose(1, 2, 0), pose_oracle.warp_affine_f32(img3, m, (72, 96))), i
    # end to end: boxes -> network input -> keypoints runs without touching the host with the crops
    assert x.shape == (5, 3, 256, 192)


def test_upsampled_decode_and_pose_entries(golden):
    """Fused bilinear-upsample + arg-max (stl_upsampled_argmax) and create_pose_from_outputs vs the reference fixture."""
    from oracle.make_golden import pose_entry_inputs
    from stlpose_b200 import pose_parsing as PP
    g = golden("pose_entries.npz")
    hm = pose_entry_inputs()
    coords, maxv = PP.get_max_preds_upsampled(hm, (256, 192))
    assert coords.shape == (3, 17, 2) and maxv.shape == (3, 17, 1)
    assert np.allclose(maxv, g["maxvals"], rtol=2e-6, atol=1e-7)
    assert np.array_equal(coords, g["coords"])              # blob maxima: top-2 margin far above the 1-ulp arithmetic slack
    entries, allk = PP.create_pose_from_outputs(torch.from_numpy(hm).cuda(), keypoint_thr=0.1)
    assert np.array_equal(np.array(entries), g["entries"]) and np.array_equal(allk, g["all_keypoints"])
    # other sizes, random maps: compare against torch's own upsample (the reference's library call) with a margin rule
    rng = np.random.default_rng(9)
    for (h, w, oh, ow) in ((96, 72, 384, 288), (16, 12, 50, 37), (8, 6, 8, 6)):
        x = rng.standard_normal((5, 17, h, w)).astype(np.float32)
        c, m = PP.get_max_preds_upsampled(x, (oh, ow))
        up = torch.nn.functional.interpolate(torch.from_numpy(x), (oh, ow), mode="bilinear", align_corners=True).numpy()
        rc, rm = pose_oracle.get_max_preds(up)
        assert np.allclose(m, rm, rtol=2e-6, atol=1e-6)
        flat = np.sort(up.reshape(5, 17, -1), axis=2)
        clear = (flat[..., -1] - flat[..., -2]) > 1e-5        # ties within float rounding may resolve either way
        assert clear.mean() > 0.9 and np.array_equal(c[clear], rc[clear])
    e0, k0 = PP.create_pose_entries([], None)
    assert e0 == [] and k0 == []


def test_generate_target_matches_reference_fixture(golden):
    """stl_generate_target (batched, device) vs targets produced by the reference's JointsDataset.generate_target."""
    from oracle.make_golden import target_inputs
    from stlpose_b200 import targets
    g = golden("targets.npz")
    joints, vis = target_inputs()
    t, w = targets.generate_target(joints, vis)
    assert t.shape == (6, 17, 64, 48) and w.shape == (6, 17, 1) and t.is_cuda
    assert np.array_equal(w.cpu().numpy(), g["weight_plain"])
    assert np.abs(t.cpu().numpy() - g["target_plain"]).max() < 3e-7           # expf vs libm: <= 2 ulp of values <= 1
    assert np.array_equal(t.cpu().numpy() > 0, g["target_plain"] > 0)         # identical support
    jw = [1., 1., 1., 1., 1., 1., 1., 1.2, 1.2, 1.5, 1.5, 1., 1., 1.2, 1.2, 1.5, 1.5]
    _, ww = targets.generate_target(joints, vis, joints_weight=jw)
    assert np.array_equal(ww.cpu().numpy(), g["weight_weighted"])
    t1, w1 = targets.generate_target(joints[1], vis[1])                       # single sample, like the reference method
    assert t1.shape == (17, 64, 48) and np.abs(t1.cpu().numpy() - g["target_plain"][1]).max() < 3e-7
    # other geometry (W48: 288x384 crops -> 72x96 maps, sigma 3) vs the oracle
    rng = np.random.default_rng(4)
    j2 = np.zeros((3, 17, 3)); j2[..., 0] = rng.uniform(-20, 300, (3, 17)); j2[..., 1] = rng.uniform(-20, 400, (3, 17))
    v2 = np.ones((3, 17, 3))
    t2, w2 = targets.generate_target(j2, v2, image_size=(288, 384), heatmap_size=(72, 96), sigma=3)
    for b in range(3):
        ot, ow = pose_oracle.generate_target(j2[b], v2[b], (288, 384), (72, 96), 3)
        assert np.abs(t2[b].cpu().numpy() - ot).max() < 3e-7 and np.array_equal(w2[b].cpu().numpy(), ow)


def test_oks_rescoring_nms_matches_reference_fixture(golden, tmp_path):
    """stl_oks_nms (rescoring + greedy OKS-NMS for all images in one launch) through the reference-shaped drop-ins vs
    the JSON written by the unmodified reference (tests/golden/submission.npz) and vs the oracle on larger random sets."""
    import json
    from oracle.make_golden import submission_inputs
    from stlpose_b200 import nms
    g = golden("submission.npz")
    preds, boxes, ids = submission_inputs()
    path = tmp_path / "preds.json"
    nms.generate_submission_hrnet([preds[:20].copy(), preds[20:].copy()], [boxes[:20].copy(), boxes[20:].copy()], list(ids),
                                  str(path))
    res = json.load(open(path))
    assert [r["image_id"] for r in res] == g["image_id"].tolist()
    assert np.array_equal(np.array([r["score"] for r in res]), g["score"])          # bit-exact (float32 sum, fp64 product)
    assert np.array_equal(np.array([r["keypoints"] for r in res]), g["keypoints"])
    assert np.array_equal(np.array([r["center"] for r in res]), g["center"])
    assert np.array_equal(np.array([r["scale"] for r in res]), g["scale"])
    big = [m for m, i in enumerate(ids) if i == 1000 + 7 * 3]
    db = [{"keypoints": preds[m], "area": boxes[m, 4], "score": boxes[m, 5]} for m in big]
    assert nms.oks_nms(db, 0.9) == g["keep_t09"].tolist()
    assert nms.oks_nms(db, 0.5) == g["keep_t05"].tolist()
    assert nms.oks_nms(db, 0.7, in_vis_thre=0.4) == g["keep_t07_vis"].tolist()
    assert nms.oks_nms([], 0.9) == []
    # larger evaluation: 300 images x up to 40 persons vs the oracle
    rng = np.random.default_rng(5)
    P, B, I = [], [], []
    for img in range(300):
        n = int(rng.integers(1, 41))
        c = rng.uniform(50, 600, (max(1, n // 3), 1, 2))
        k = c[rng.integers(len(c), size=n)] + rng.normal(0, 30, (1, 17, 2)) + rng.normal(0, 3, (n, 17, 2))
        P.append(np.concatenate([k, rng.uniform(0, 1, (n, 17, 1))], axis=2).astype(np.float32))
        sc = rng.uniform(0.3, 2.0, (n, 1)) * np.array([0.75, 1.0])
        B.append(np.concatenate([k.mean(1), sc, np.prod(sc * 200, 1, keepdims=True), rng.uniform(0.2, 1, (n, 1))], axis=1))
        I += [img] * n
    P, B = np.concatenate(P), np.concatenate(B)
    got = nms.rescore_and_nms(P, B, I)
    want = pose_oracle.rescore_and_nms(P, B, I)
    assert len(got) == len(want) == 300
    kept = 0
    for gi, wi in zip(got, want):
        assert [p["score"] for p in gi] == [s for _, s in wi]
        assert all(np.array_equal(p["keypoints"], P[m]) for p, (m, _) in zip(gi, wi))
        kept += len(wi)
    assert kept < len(I) * 0.8                                                        # the NMS really suppressed
    # crowded images: lib/nms.py has no limit on the persons of one image (the kernel's shared arrays are sized per call)
    P2, B2, I2 = [], [], []
    for img, n in enumerate((700, 3, 129)):
        c = rng.uniform(50, 600, (n // 4 + 1, 1, 2))
        k = c[rng.integers(len(c), size=n)] + rng.normal(0, 30, (1, 17, 2)) + rng.normal(0, 3, (n, 17, 2))
        P2.append(np.concatenate([k, rng.uniform(0, 1, (n, 17, 1))], axis=2).astype(np.float32))
        sc = rng.uniform(0.3, 2.0, (n, 1)) * np.array([0.75, 1.0])
        B2.append(np.concatenate([k.mean(1), sc, np.prod(sc * 200, 1, keepdims=True), rng.uniform(0.2, 1, (n, 1))], axis=1))
        I2 += [img] * n
    P2, B2 = np.concatenate(P2), np.concatenate(B2)
    got2, want2 = nms.rescore_and_nms(P2, B2, I2), pose_oracle.rescore_and_nms(P2, B2, I2)
    for gi, wi in zip(got2, want2):
        assert [p["score"] for p in gi] == [s for _, s in wi]
        assert all(np.array_equal(p["keypoints"], P2[m]) for p, (m, _) in zip(gi, wi))
