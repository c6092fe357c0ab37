"""Drop-in for the keypoint part of the reference ``lib.nms`` (/root/reference/src/lib/nms.py:10-74), for
``generate_submission_hrnet`` (/root/reference/src/lib/metrics.py:192-265: rescoring + OKS-NMS loop on the device) and for
the COCO result packing it ends with (``convert_keypoints_to_coco_format``, /root/reference/src/data/data_processing.py:52-82)."""
import json

import numpy as np
import torch

from . import _lib

COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def _run(kpts, areas, scores, offsets, sigmas, in_vis_thr, oks_thr, nms_vis_thr, rescore):
    L = _lib.lib()
    M, J = kpts.shape[0], kpts.shape[1]
    if not torch.cuda.is_available():
        raise RuntimeError("stlpose_b200.nms needs a CUDA device (there is no CPU path)")
    dev = "cuda"
    k = torch.as_tensor(np.ascontiguousarray(kpts, dtype=np.float32)).to(dev)
    a = torch.as_tensor(np.ascontiguousarray(areas, dtype=np.float64)).to(dev)
    s = torch.as_tensor(np.ascontiguousarray(scores, dtype=np.float64)).to(dev)
    o = torch.as_tensor(np.ascontiguousarray(offsets, dtype=np.int32)).to(dev)
    sig = COCO_SIGMAS if not isinstance(sigmas, np.ndarray) else sigmas
    v = torch.as_tensor(np.ascontiguousarray((np.asarray(sig, np.float64) * 2) ** 2)).to(dev)
    score_out = torch.empty(M, dtype=torch.float64, device=dev)
    keep_rank = torch.empty(M, dtype=torch.int32, device=dev)
    max_p = int(np.max(np.diff(offsets))) if len(offsets) > 1 else 0
    with torch.cuda.device(k.device):
        _lib.check(L.stl_oks_nms(_lib.ptr(k), _lib.ptr(a), _lib.ptr(s), _lib.ptr(o), len(offsets) - 1, max_p, J, _lib.ptr(v),
                                 float(in_vis_thr), float(oks_thr), float(nms_vis_thr), int(rescore), _lib.ptr(score_out),
                                 _lib.ptr(keep_rank), _lib.current_stream()))
    return score_out.cpu().numpy(), keep_rank.cpu().numpy()


def oks_nms(kpts_db, thresh, sigmas=None, in_vis_thre=None):
    """lib/nms.py:10-46: indices of ``kpts_db`` to keep, in descending-score order."""
    if len(kpts_db) == 0:
        return []
    kpts = np.array([np.asarray(p["keypoints"]).reshape(-1, 3) for p in kpts_db])
    areas = np.array([p["area"] for p in kpts_db])
    scores = np.array([p["score"] for p in kpts_db])
    _, rank = _run(kpts, areas, scores, [0, len(kpts_db)], sigmas, 0.0, thresh,
                   -1.0 if in_vis_thre is None else in_vis_thre, rescore=False)
    kept = np.where(rank >= 0)[0]
    return [int(i) for i in kept[np.argsort(rank[kept])]]


def rescore_and_nms(all_preds, all_bboxes, image_ids, in_vis_thr=0.2, oks_thr=0.9, sigmas=None):
    """The rescoring + OKS-NMS loop of generate_submission_hrnet (lib/metrics.py:232-258) for a whole evaluation.

    all_preds [M,17,3] (x, y, score), all_bboxes [M,6] (center 0:2, scale 2:4, area 4, score 5), image_ids: M ids.
    Returns a list (one entry per image, in first-appearance order) of lists of person dicts with the reference's keys
    ('keypoints', 'center', 'scale', 'area', 'score' [rescored], 'image'), ordered by descending score."""
    all_preds, all_bboxes = np.asarray(all_preds), np.asarray(all_bboxes)
    order_of, groups = {}, []
    for m, img in enumerate(image_ids):
        if img not in order_of:
            order_of[img] = len(groups)
            groups.append([])
        groups[order_of[img]].append(m)
    perm = np.array([m for g in groups for m in g], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum([len(g) for g in groups])])
    if len(perm) == 0:
        return []
    scores, rank = _run(all_preds[perm], all_bboxes[perm, 4], all_bboxes[perm, 5], offsets, sigmas, in_vis_thr, oks_thr,
                        -1.0, rescore=True)
    out = []
    for gi, g in enumerate(groups):
        lo = offsets[gi]
        local = [(rank[lo + i], i) for i in range(len(g)) if rank[lo + i] >= 0]
        persons = []
        for _, i in sorted(local):
            m = g[i]
            persons.append({"keypoints": all_preds[m], "center": all_bboxes[m][0:2], "scale": all_bboxes[m][2:4],
                            "area": all_bboxes[m][4], "score": scores[lo + i], "image": image_ids[m]})
        out.append(persons)
    return out


def convert_keypoints_to_coco_format(keypoints, res_file=None):
    """data/data_processing.py:52-82: list (per image) of lists of person dicts -> flat list of COCO result dicts
    ('keypoints' = 51 float64 values).  ``res_file`` is unused, as in the reference."""
    results = []
    for img_kpts in keypoints:
        for person in img_kpts:
            results.append({"image_id": person["image"], "category_id": 1,
                            "keypoints": list(np.asarray(person["keypoints"], dtype=np.float64).reshape(-1)),
                            "score": person["score"], "center": list(person["center"]), "scale": list(person["scale"])})
    return results


def generate_submission_hrnet(all_preds, all_bboxes, image_ids, preds_file, name=False):
    """lib/metrics.py:192-265 with the reference's arguments: lists of per-batch arrays ([n,17,3] and [n,6]) as
    03_evaluate.py:185-198 collects them, the image ids, the JSON path.  Thresholds as hard-coded there (0.2 / 0.9)."""
    all_preds = np.concatenate(all_preds, axis=0)
    all_bboxes = np.concatenate(all_bboxes, axis=0)
    if name:
        image_ids = [int(n[-16:-4]) for n in image_ids]
    results = convert_keypoints_to_coco_format(rescore_and_nms(all_preds, all_bboxes, image_ids), preds_file)
    with open(preds_file, "w") as f:
        json.dump(results, f)
    return
