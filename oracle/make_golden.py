"""Generate tests/golden/*.npz by running the UNMODIFIED reference code (build container only).

Test infrastructure.  Run:  python -m oracle.make_golden
Needs /root/reference; the fixtures it writes are committed so that the GPU box (which has no
/root/reference) can still check the oracle and the CUDA path against the reference's outputs.

Inputs are regenerated from seeds by tests (``golden_inputs``) except where platform libm could
perturb them (Gaussian blobs), which are stored as float16-exact values.
"""
import os
import sys

import numpy as np
import torch

from oracle import hrnet_oracle, pose_oracle, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_inputs():
    """Seeded inputs shared by the generator and the tests."""
    rng = np.random.default_rng(1234)
    d = {}
    d["hm_rand_64x48"] = rng.standard_normal((2, 17, 64, 48)).astype(np.float32)
    d["hm_rand_96x72"] = rng.standard_normal((1, 17, 96, 72)).astype(np.float32)
    d["center_a"], d["scale_a"] = pose_oracle.synth_boxes(2, seed=7)
    d["center_b"], d["scale_b"] = pose_oracle.synth_boxes(1, seed=8)
    d["flip_out"] = rng.standard_normal((2, 17, 64, 48)).astype(np.float32)
    d["flip_out_f"] = rng.standard_normal((2, 17, 64, 48)).astype(np.float32)
    d["loss_out"] = rng.standard_normal((3, 17, 64, 48)).astype(np.float32)
    d["loss_tgt"] = rng.random((3, 17, 64, 48)).astype(np.float32)
    d["loss_tw"] = rng.choice(np.array([0, 1, 1.2, 1.5], np.float32), size=(3, 17, 1))
    d["x_w32"] = rng.standard_normal((2, 3, 256, 192)).astype(np.float32)
    d["x_w48"] = rng.standard_normal((1, 3, 384, 288)).astype(np.float32)
    return d


class _TwoShotModel:
    """Stand-in 'model' for lib.inference.forward_pass: returns fixed heatmaps for the plain and the
    flipped call, so the reference's flip_back / shift / average code runs on known inputs."""

    def __init__(self, out, out_f):
        self.outs = [torch.from_numpy(out.copy()), torch.from_numpy(out_f.copy())]
        self.calls = 0

    def __call__(self, img):
        r = self.outs[self.calls]
        self.calls += 1
        return r


def pck_inputs():
    """Predicted / ground-truth heatmaps for the PCK fixture: blobs whose predicted peak is displaced by 0-4 pixels,
    some joints unlabeled (all-zero target -> arg-max (0,0) -> not counted), one joint never labeled."""
    rng = np.random.default_rng(77)
    B, J, h, w = 6, 17, 64, 48
    tgt = np.zeros((B, J, h, w), np.float32)
    out = (rng.standard_normal((B, J, h, w)) * 0.05).astype(np.float32)
    for n in range(B):
        for j in range(J):
            if j == 5 or rng.random() < 0.2:
                continue
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            tgt[n, j, y, x] = 1.0
            dx, dy = rng.integers(-4, 5, size=2)
            out[n, j, min(max(y + dy, 0), h - 1), min(max(x + dx, 0), w - 1)] += 1.0
    return out, tgt


def make_pck():
    """PCK fixture: the reference's own calc_dists / dist_acc (source text, see ref_shim.metrics_functions) on the
    arg-max coordinates produced by the reference's get_max_preds_hrnet."""
    L, M = ref_shim.lib(), ref_shim.metrics_functions()
    out, tgt = pck_inputs()
    pred, _ = L.pose_parsing.get_max_preds_hrnet(out)
    tcoord, _ = L.pose_parsing.get_max_preds_hrnet(tgt)
    norm = np.ones((pred.shape[0], 2)) * np.array([out.shape[2], out.shape[3]]) / 10      # metrics.py:347
    dists = M.calc_dists(pred, tcoord, norm)
    per_joint = np.array([M.dist_acc(dists[j]) for j in range(out.shape[1])], np.float64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pck.npz"), dists=dists, per_joint=per_joint, pred=pred)
    print("pck per-joint", np.round(per_joint, 3))


def crop_inputs():
    """A 427x640 uint8 image (smooth ramps + blobs + sparse texture, so the fixture compresses) and person boxes that
    cover: a plain box, fractional coordinates, boxes leaving the image on every side, a wide box (aspect fix-up)."""
    rng = np.random.default_rng(2024)
    H, W = 427, 640
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([(xx * 255 // W), (yy * 255 // H), ((xx + yy) % 256)], axis=2).astype(np.float64)
    for _ in range(40):
        cx, cy, r = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(5, 40)
        img += rng.uniform(-120, 120, size=3) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))[..., None]
    tex = rng.integers(-60, 60, size=(H // 4 + 1, W // 4 + 1, 3)).repeat(4, 0).repeat(4, 1)[:H, :W]
    img = np.clip(img + tex, 0, 255).astype(np.uint8)
    boxes = [[50, 30, 210, 400], [300.5, 100.2, 420.7, 380.1], [-20, -10, 120, 200], [500, 300, 700, 500],
             [10, 10, 630, 100]]
    return img, boxes


def make_crops():
    """Crop-extraction fixture: the reference's TransformDetection.__call__ (cv2.warpAffine inside) and crop() with a
    rotation, run here with OpenCV 4.13."""
    L = ref_shim.lib()
    img, boxes = crop_inputs()
    dets, centers, scales = L.transforms.TransformDetection()(img, boxes)
    rot = L.transforms.crop(img, centers[1], scales[1], np.array([192, 256]), rot=30)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "crops.npz"), dets=dets, centers=centers, scales=scales, rot30=rot)
    print("crops", dets.shape, dets.dtype, "mean", dets.mean())
    # float32 image, as 04_evaluate_vases_qualitatively.py:209-210 passes it (a [0, 1] CHW tensor turned HWC): one box
    # is enough to pin the float interpolation path (stored as float16-exact inputs -> exact float32 outputs)
    imgf = (img.astype(np.float32) / np.float32(255)).astype(np.float16).astype(np.float32)
    dets_f, _, _ = L.transforms.TransformDetection()(imgf, boxes[1:2])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "crops_f32.npz"), dets=dets_f)
    print("crops f32", dets_f.shape, dets_f.dtype, "mean", dets_f.mean())


def pose_entry_inputs():
    """Blob heatmaps [3,17,64,48] (float16-exact values so platform libm cannot perturb them) with some weak joints."""
    hm = pose_oracle.blob_heatmaps(3, 17, 64, 48, seed=21, noise=0.01).astype(np.float16).astype(np.float32)
    hm[0, 3] *= 0.05            # below the 0.1 confidence threshold
    hm[1, 7] *= 0.02
    hm[2, 16] = -np.abs(hm[2, 16])   # nothing positive: coordinates zeroed, visibility 0
    return hm


def make_pose_entries():
    import torch
    L = ref_shim.lib()
    hm = pose_entry_inputs()
    entries, allk = L.pose_parsing.create_pose_from_outputs(torch.from_numpy(hm), keypoint_thr=0.1)
    import torch.nn.functional as F
    scaled = F.interpolate(torch.from_numpy(hm).clone(), (256, 192), mode="bilinear", align_corners=True)
    coords, maxv = L.pose_parsing.get_max_preds_hrnet(scaled.numpy())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pose_entries.npz"), entries=np.array(entries), all_keypoints=allk,
                        coords=coords, maxvals=maxv)
    print("pose entries", np.array(entries).shape, allk.shape)


def target_inputs():
    """Joints [B,17,3] in crop pixels (256x192) and visibilities: inside, on every border, far outside, invisible."""
    rng = np.random.default_rng(99)
    B, J = 6, 17
    joints = np.zeros((B, J, 3))
    joints[..., 0] = rng.uniform(-40, 232, size=(B, J))
    joints[..., 1] = rng.uniform(-40, 296, size=(B, J))
    joints[0, 0, :2] = (0.0, 0.0)
    joints[0, 1, :2] = (191.9, 255.9)
    joints[0, 2, :2] = (-30.0, 100.0)       # patch entirely left of the map
    joints[0, 3, :2] = (100.0, 290.0)       # patch entirely below
    joints[0, 4, :2] = (-22.1, -22.1)       # patch just touching the corner
    vis = (rng.random((B, J)) > 0.25).astype(np.float64)
    joints_vis = np.stack([vis, vis, np.zeros_like(vis)], axis=2)
    return joints, joints_vis


def make_targets():
    import types
    f = ref_shim.generate_target_function()
    joints, joints_vis = target_inputs()
    jw = np.array([1., 1., 1., 1., 1., 1., 1., 1.2, 1.2, 1.5, 1.5, 1., 1., 1.2, 1.2, 1.5, 1.5], np.float32).reshape(17, 1)
    outs = {}
    for tag, diff in (("plain", False), ("weighted", True)):
        me = types.SimpleNamespace(num_joints=17, target_type="gaussian", heatmap_size=np.array([48, 64]),
                                   image_size=np.array([192, 256]), sigma=2, use_different_joints_weight=diff,
                                   joints_weight=jw)
        res = [f(me, joints[b], joints_vis[b]) for b in range(joints.shape[0])]
        outs[f"target_{tag}"] = np.stack([r[0] for r in res])
        outs[f"weight_{tag}"] = np.stack([r[1] for r in res])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "targets.npz"), **outs)
    print("targets", outs["target_plain"].shape, float(outs["target_plain"].sum()), outs["weight_weighted"][0, :6, 0])


def submission_inputs():
    """Persons of an evaluation: all_preds [M,17,3] float32 (x, y, max value), all_bboxes [M,6] float64 (center, scale,
    area, box score), image ids interleaved across images.  Images hold 1..24 persons; many are jittered duplicates of
    another person of the same image (so the OKS-NMS suppresses), one person has no joint above the visibility threshold."""
    rng = np.random.default_rng(2024)
    preds, boxes, ids = [], [], []
    for img, n in enumerate((1, 2, 5, 24, 9, 3)):
        base = []
        for p in range(n):
            if base and rng.random() < 0.55:
                src = base[rng.integers(len(base))]
                k = src + rng.normal(0, rng.choice([0.3, 2.0, 8.0]), size=src.shape)
            else:
                c = rng.uniform(60, 500, 2)
                k = c + rng.normal(0, 40, size=(17, 2))
                base.append(k)
            sc = rng.uniform(0.05, 0.98, size=(17, 1))
            scale = rng.uniform(0.4, 2.5) * np.array([0.75, 1.0])
            preds.append(np.concatenate([k, sc], axis=1).astype(np.float32))
            boxes.append(np.concatenate([k.mean(0), scale, [np.prod(scale * 200)], [rng.uniform(0.3, 1.0)]]))
            ids.append(1000 + 7 * img)
    preds, boxes = np.stack(preds), np.stack(boxes)
    preds[0, :, 2] = rng.uniform(0.01, 0.19, 17).astype(np.float32)      # nothing visible: rescored to 0 (kept: alone)
    perm = rng.permutation(len(ids))                                     # images interleaved, as batches arrive
    return preds[perm], boxes[perm], [ids[i] for i in perm]


def make_submission():
    import json
    import tempfile
    F = ref_shim.submission_functions()
    preds, boxes, ids = submission_inputs()
    path = os.path.join(tempfile.mkdtemp(prefix="stlpose_sub_"), "preds.json")
    # the reference takes lists of per-batch arrays (03_evaluate.py:185-186) and concatenates them
    F.generate_submission_hrnet([preds[:20].copy(), preds[20:].copy()], [boxes[:20].copy(), boxes[20:].copy()], list(ids), path)
    results = json.load(open(path))
    out = {"image_id": np.array([r["image_id"] for r in results]),
           "keypoints": np.array([r["keypoints"] for r in results]),
           "score": np.array([r["score"] for r in results]),
           "center": np.array([r["center"] for r in results]), "scale": np.array([r["scale"] for r in results])}
    # lib.nms.oks_nms directly on the biggest image: default call, other thresholds, the visibility mask
    big = [m for m, i in enumerate(ids) if i == 1000 + 7 * 3]
    db = [{"keypoints": preds[m], "area": boxes[m, 4], "score": boxes[m, 5]} for m in big]
    for tag, kw in (("t09", dict(thresh=0.9)), ("t05", dict(thresh=0.5)), ("t07_vis", dict(thresh=0.7, in_vis_thre=0.4))):
        out[f"keep_{tag}"] = np.array(F.nms.oks_nms(db, **kw))
    g = preds[big[0]].reshape(-1)
    d = np.stack([preds[m].reshape(-1) for m in big[1:]])
    out["iou_plain"] = F.nms.oks_iou(g, d, boxes[big[0], 4], boxes[big[1:], 4])
    out["iou_vis"] = F.nms.oks_iou(g, d, boxes[big[0], 4], boxes[big[1:], 4], in_vis_thre=0.4)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "submission.npz"), **out)
    print("submission", len(results), "of", len(ids), "persons kept;", {k: v.tolist() for k, v in out.items() if k.startswith("keep")})


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "submission":
        return make_submission()
    if len(sys.argv) > 1 and sys.argv[1] == "targets":
        return make_targets()
    if len(sys.argv) > 1 and sys.argv[1] == "pose_entries":
        return make_pose_entries()
    if len(sys.argv) > 1 and sys.argv[1] == "pck":
        return make_pck()
    if len(sys.argv) > 1 and sys.argv[1] == "crops":
        return make_crops()
    if not ref_shim.available():
        sys.exit("reference not available; goldens can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    L = ref_shim.lib()
    g = golden_inputs()

    # ---- decode ------------------------------------------------------------------------------
    blobs = pose_oracle.blob_heatmaps(2, 17, 64, 48, seed=3).astype(np.float16).astype(np.float32)
    center_c, scale_c = pose_oracle.synth_boxes(2, seed=9)
    out = {"hm_blobs_f16": blobs.astype(np.float16)}
    for tag, hm, c, s in (("rand64", g["hm_rand_64x48"], g["center_a"], g["scale_a"]),
                          ("rand96", g["hm_rand_96x72"], g["center_b"], g["scale_b"]),
                          ("blobs", blobs, center_c, scale_c)):
        p0, m0 = L.pose_parsing.get_max_preds_hrnet(hm.copy())
        preds, maxvals, coords = L.pose_parsing.get_final_preds_hrnet(hm.copy(), c, s)
        out[f"{tag}_max_preds"] = p0
        out[f"{tag}_max_vals"] = m0
        out[f"{tag}_preds"] = preds
        out[f"{tag}_maxvals"] = maxvals
        out[f"{tag}_coords"] = coords
    np.savez_compressed(os.path.join(GOLDEN_DIR, "decode.npz"), **out)

    # ---- flip-test averaging -----------------------------------------------------------------
    model = _TwoShotModel(g["flip_out"], g["flip_out_f"])
    avg = L.inference.forward_pass(model, torch.zeros(2, 3, 8, 8), "HRNet", device="cpu", flip=True)
    fb = L.transforms.flip_back(g["flip_out_f"].copy(), L.CONSTANTS.FLIP_PAIRS)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "flip.npz"), avg=avg.numpy(), flip_back=fb.numpy())

    # ---- loss --------------------------------------------------------------------------------
    o = torch.from_numpy(g["loss_out"]).requires_grad_(True)
    loss = L.loss.PersonMSELoss()(o, torch.from_numpy(g["loss_tgt"]), torch.from_numpy(g["loss_tw"]))
    loss.backward()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "loss.npz"), loss=loss.detach().numpy(),
                        grad=o.grad.numpy())

    # ---- HRNet forward -----------------------------------------------------------------------
    for width, hw, key in ((32, (256, 192), "x_w32"), (48, (384, 288), "x_w48")):
        m = ref_shim.build_reference_hrnet(width, hw).eval()
        m.load_state_dict(hrnet_oracle.synth_state_dict(width, seed=0), strict=True)
        with torch.no_grad():
            y = m(torch.from_numpy(g[key]))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"hrnet_w{width}_fwd.npz"), y=y.numpy())
        print(f"w{width}: y range [{y.min():.3f}, {y.max():.3f}] std {y.std():.3f}")
    make_pck()
    make_crops()
    make_pose_entries()
    make_targets()
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
