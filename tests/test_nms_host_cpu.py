"""Host side of the OKS drop-ins (no GPU): grouping / packing around stl_oks_nms, checked against the JSON the
unmodified reference wrote (tests/golden/submission.npz) with the oracle standing in for the device call."""
import numpy as np

from oracle import pose_oracle
from oracle.make_golden import submission_inputs


def test_coco_result_packing_matches_reference_fixture(golden, monkeypatch):
    from stlpose_b200 import nms
    g = golden("submission.npz")
    preds, boxes, ids = submission_inputs()

    def fake_run(kpts, areas, scores, offsets, sigmas, in_vis_thr, oks_thr, nms_vis_thr, rescore):
        # what the device returns: rescored scores and the position of every person in its image's keep list
        sc = pose_oracle.rescore(kpts, scores, in_vis_thr) if rescore else np.asarray(scores, np.float64)
        rank = np.full(len(kpts), -1, np.int32)
        for i in range(len(offsets) - 1):
            lo, hi = offsets[i], offsets[i + 1]
            keep = pose_oracle.oks_nms(kpts[lo:hi], sc[lo:hi], areas[lo:hi], oks_thr,
                                       in_vis_thre=None if nms_vis_thr < 0 else nms_vis_thr)
            for r, k in enumerate(keep):
                rank[lo + k] = r
        return sc, rank

    monkeypatch.setattr(nms, "_run", fake_run)
    res = nms.convert_keypoints_to_coco_format(nms.rescore_and_nms(preds, boxes, ids))
    assert [r["image_id"] for r in res] == g["image_id"].tolist()
    assert np.array_equal(np.array([r["score"] for r in res]), g["score"])
    assert np.array_equal(np.array([r["keypoints"] for r in res]), g["keypoints"])
    assert np.array_equal(np.array([r["center"] for r in res]), g["center"])
    assert np.array_equal(np.array([r["scale"] for r in res]), g["scale"])
    assert all(r["category_id"] == 1 and len(r["keypoints"]) == 51 for r in res)
    big = [m for m, i in enumerate(ids) if i == 1000 + 7 * 3]
    db = [{"keypoints": preds[m], "area": boxes[m, 4], "score": boxes[m, 5]} for m in big]
    assert nms.oks_nms(db, 0.9) == g["keep_t09"].tolist()
    assert nms.oks_nms(db, 0.7, in_vis_thre=0.4) == g["keep_t07_vis"].tolist()
    assert nms.oks_nms([], 0.9) == [] and nms.rescore_and_nms(preds[:0], boxes[:0], []) == []
