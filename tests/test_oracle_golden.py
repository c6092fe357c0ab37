"""Pin the CPU oracle against fixtures produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import numpy as np
import torch

from oracle import hrnet_oracle, pose_oracle


def _decode_case(golden, golden_inputs, tag):
    g = golden("decode.npz")
    if tag == "rand64":
        return g, golden_inputs["hm_rand_64x48"], golden_inputs["center_a"], golden_inputs["scale_a"]
    if tag == "rand96":
        return g, golden_inputs["hm_rand_96x72"], golden_inputs["center_b"], golden_inputs["scale_b"]
    c, s = pose_oracle.synth_boxes(2, seed=9)
    return g, g["hm_blobs_f16"].astype(np.float32), c, s


def test_get_max_preds_bit_exact(golden, golden_inputs):
    for tag in ("rand64", "rand96", "blobs"):
        g, hm, _, _ = _decode_case(golden, golden_inputs, tag)
        p, m = pose_oracle.get_max_preds(hm)
        assert np.array_equal(p, g[f"{tag}_max_preds"])
        assert np.array_equal(m, g[f"{tag}_max_vals"])
        assert p.dtype == np.float32 and m.dtype == np.float32


def test_get_final_preds(golden, golden_inputs):
    for tag in ("rand64", "rand96", "blobs"):
        g, hm, c, s = _decode_case(golden, golden_inputs, tag)
        preds, maxvals, coords = pose_oracle.get_final_preds(hm, c, s)
        assert np.array_equal(coords, g[f"{tag}_coords"])          # quarter-pixel offsets are exact
        assert np.array_equal(maxvals, g[f"{tag}_maxvals"])
        # float64 3-point solve vs cv2.getAffineTransform: 1e-3 px bound
        assert np.abs(preds - g[f"{tag}_preds"]).max() < 1e-3
        frac = np.mod(coords, 1.0)
        assert np.all(np.isin(frac, [0.0, 0.25, 0.75]))


def test_flip_average_bit_exact(golden, golden_inputs):
    g = golden("flip.npz")
    fb = pose_oracle.flip_back(golden_inputs["flip_out_f"])
    assert np.array_equal(fb, g["flip_back"])
    avg = pose_oracle.flip_average(golden_inputs["flip_out"], golden_inputs["flip_out_f"])
    assert np.array_equal(avg, g["avg"])


def test_person_mse_loss(golden, golden_inputs):
    g = golden("loss.npz")
    loss, grad = pose_oracle.person_mse_loss(golden_inputs["loss_out"], golden_inputs["loss_tgt"],
                                             golden_inputs["loss_tw"])
    assert abs(loss - float(g["loss"])) < 1e-6 * max(1.0, abs(loss))
    assert np.abs(grad - g["grad"]).max() < 1e-9


def test_hrnet_forward_w32(golden, golden_inputs):
    y_ref = golden("hrnet_w32_fwd.npz")["y"]
    sd = hrnet_oracle.synth_state_dict(32, seed=0)
    y = hrnet_oracle.hrnet_forward(sd, torch.from_numpy(golden_inputs["x_w32"]), 32).numpy()
    assert y.shape == (2, 17, 64, 48)
    # same fp32 library kernels; allow for thread-count dependent summation order
    assert np.abs(y - y_ref).max() < 1e-4


def test_hrnet_forward_w48(golden, golden_inputs):
    y_ref = golden("hrnet_w48_fwd.npz")["y"]
    sd = hrnet_oracle.synth_state_dict(48, seed=0)
    y = hrnet_oracle.hrnet_forward(sd, torch.from_numpy(golden_inputs["x_w48"]), 48).numpy()
    assert y.shape == (1, 17, 96, 72)
    assert np.abs(y - y_ref).max() < 1e-4


def test_flop_count_matches_survey():
    assert abs(hrnet_oracle.conv_flops_per_crop(32, (256, 192)) / 1e9 - 15.290) < 1e-3
    assert abs(hrnet_oracle.conv_flops_per_crop(48, (384, 288)) / 1e9 - 70.613) < 1e-3


def test_known_answers():
    """Hand-made cases: ties -> first index; all<=0 -> (0,0); border peaks unrefined."""
    hm = np.full((1, 4, 8, 6), -1.0, np.float32)
    hm[0, 0, 3, 2] = 5.0
    hm[0, 0, 5, 4] = 5.0                     # tie: first (row-major) wins
    hm[0, 2, 0, 5] = 2.0                     # border peak: no refinement
    hm[0, 3, 4, 3] = 1.0
    hm[0, 3, 4, 4] = 0.5                     # dx>0 -> +0.25 ; dy: hm[5,3]-hm[3,3] = 0 -> 0
    c = np.array([[100.0, 100.0]])
    s = np.array([[0.75, 1.0]])
    preds, maxvals, coords = pose_oracle.get_final_preds(hm, c, s)
    assert coords[0, 0].tolist() == [2.0, 3.0]   # flat neighbourhood: sign(0) = 0, no shift
    p0, m0 = pose_oracle.get_max_preds(hm)
    assert p0[0, 0].tolist() == [2.0, 3.0]
    assert p0[0, 1].tolist() == [0.0, 0.0] and m0[0, 1, 0] == -1.0
    assert coords[0, 2].tolist() == [5.0, 0.0]
    assert coords[0, 3].tolist() == [3.25, 4.0]
    # closed form of the inverse affine with rot=0: k = scale[0]*200/w
    k = s[0, 0] * 200.0 / 6
    expect = c[0] + (coords[0, 3] - np.array([3.0, 4.0])) * k
    assert np.abs(preds[0, 3] - expect).max() < 1e-3
    assert pose_oracle.get_max_preds(np.zeros((0, 17, 8, 6), np.float32)) == ([], [])


def test_pck_accuracy_oracle_matches_reference_fixture(golden):
    """oracle calc_dists / dist_acc / accuracy vs the fixture produced by the reference's own function bodies."""
    from oracle.make_golden import pck_inputs
    g = golden("pck.npz")
    out, tgt = pck_inputs()
    acc, avg, cnt, pred = pose_oracle.accuracy(out, tgt)
    assert np.array_equal(pred, g["pred"])
    assert np.allclose(acc[1:], g["per_joint"], atol=0, rtol=0)
    valid = g["per_joint"] >= 0
    assert cnt == int(valid.sum()) and abs(avg - g["per_joint"][valid].mean()) < 1e-12 and acc[0] == avg
    p2, _ = pose_oracle.get_max_preds(out)
    t2, _ = pose_oracle.get_max_preds(tgt)
    d = pose_oracle.calc_dists(p2, t2, np.ones((out.shape[0], 2)) * np.array([64, 48]) / 10)
    assert np.allclose(d, g["dists"], atol=1e-12)
    assert g["per_joint"][5] == -1                          # never-labeled joint


def test_crop_extraction_oracle_matches_reference_fixture(golden):
    """TransformDetection / crop restatement (cv2.getAffineTransform LU + cv2.warpAffine fixed-point bilinear) is
    bit-exact against crops produced by the reference's own code."""
    from oracle.make_golden import crop_inputs
    g = golden("crops.npz")
    img, boxes = crop_inputs()
    dets, centers, scales = pose_oracle.transform_detection(img, boxes)
    assert np.array_equal(centers, g["centers"]) and np.array_equal(scales, g["scales"])
    assert dets.dtype == np.uint8 and np.array_equal(dets, g["dets"])
    m = pose_oracle.forward_affine(g["centers"][1], g["scales"][1], 30, (192, 256))
    assert np.array_equal(pose_oracle.warp_affine_u8(img, m, (192, 256)), g["rot30"])
    d0, c0, s0 = pose_oracle.transform_detection(img, [])
    assert len(d0) == 0 and len(c0) == 0
    # float32 image (04_evaluate_vases_qualitatively.py:209-213): cv2 interpolates in float32 with its weight table
    gf = golden("crops_f32.npz")
    imgf = (img.astype(np.float32) / np.float32(255)).astype(np.float16).astype(np.float32)
    dets_f, _, _ = pose_oracle.transform_detection(imgf, boxes[1:2])
    assert dets_f.dtype == np.float32 and np.array_equal(dets_f, gf["dets"])


def test_upsampled_decode_and_pose_entries_match_reference_fixture(golden):
    from oracle.make_golden import pose_entry_inputs
    g = golden("pose_entries.npz")
    hm = pose_entry_inputs()
    coords, maxv = pose_oracle.upsampled_max_preds(hm)
    assert np.array_equal(coords, g["coords"]) and np.allclose(maxv, g["maxvals"], rtol=1e-6, atol=1e-7)
    entries, allk = pose_oracle.create_pose_from_outputs(hm, keypoint_thr=0.1)
    assert np.array_equal(np.array(entries), g["entries"]) and np.array_equal(allk, g["all_keypoints"])
    assert allk[0 * 17 + 3, 3] == 0 and allk[1 * 17 + 7, 3] == 0 and allk[2 * 17 + 16, 3] == 0


def test_generate_target_oracle_matches_reference_fixture(golden):
    from oracle.make_golden import target_inputs
    g = golden("targets.npz")
    joints, vis = target_inputs()
    jw = np.array([1., 1., 1., 1., 1., 1., 1., 1.2, 1.2, 1.5, 1.5, 1., 1., 1.2, 1.2, 1.5, 1.5], np.float32).reshape(17, 1)
    for b in range(joints.shape[0]):
        t, w = pose_oracle.generate_target(joints[b], vis[b])
        assert np.array_equal(t, g["target_plain"][b]) and np.array_equal(w, g["weight_plain"][b])
        t, w = pose_oracle.generate_target(joints[b], vis[b], joints_weight=jw)
        assert np.array_equal(w, g["weight_weighted"][b])
    assert g["weight_plain"][0, 3, 0] == 0                  # patch entirely below the map: weight forced to 0
    assert not g["target_plain"][0, 2].any()                # patch touching the left border from outside: empty map


def test_oks_rescoring_nms_oracle_matches_reference_fixture(golden):
    """Oracle restatement of generate_submission_hrnet's rescoring + lib.nms.oks_nms / oks_iou + the COCO result
    packing vs the JSON the unmodified reference wrote (tests/golden/submission.npz)."""
    from oracle.make_golden import submission_inputs
    g = golden("submission.npz")
    preds, boxes, ids = submission_inputs()
    kept = pose_oracle.rescore_and_nms(preds, boxes, ids)
    res = pose_oracle.coco_results(preds, boxes, ids, kept)
    assert len(res) == len(g["score"]) and len(res) < len(ids)             # something was suppressed
    assert [r["image_id"] for r in res] == g["image_id"].tolist()
    assert np.array_equal(np.array([r["score"] for r in res]), g["score"])  # bit-exact rescoring
    assert np.array_equal(np.array([r["keypoints"] for r in res]), g["keypoints"])
    assert np.array_equal(np.array([r["center"] for r in res]), g["center"])
    assert np.array_equal(np.array([r["scale"] for r in res]), g["scale"])
    assert (g["score"] == 0).sum() == 1                                     # the person without a visible joint
    big = [m for m, i in enumerate(ids) if i == 1000 + 7 * 3]
    kp, ar, sc = preds[big], boxes[big, 4], boxes[big, 5]
    assert pose_oracle.oks_nms(kp, sc, ar, 0.9) == g["keep_t09"].tolist()
    assert pose_oracle.oks_nms(kp, sc, ar, 0.5) == g["keep_t05"].tolist()
    assert pose_oracle.oks_nms(kp, sc, ar, 0.7, in_vis_thre=0.4) == g["keep_t07_vis"].tolist()
    flat = kp.reshape(len(big), -1)
    assert np.array_equal(pose_oracle.oks_iou(flat[0], flat[1:], ar[0], ar[1:]), g["iou_plain"])
    assert np.array_equal(pose_oracle.oks_iou(flat[0], flat[1:], ar[0], ar[1:], in_vis_thre=0.4), g["iou_vis"])
    assert pose_oracle.oks_nms(kp[:0], sc[:0], ar[:0], 0.9) == []
