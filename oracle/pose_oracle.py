"""NumPy restatement of the reference's heatmap decode / flip-test / loss arithmetic.

Test infrastructure (see oracle/__init__.py).  Each function cites the reference lines it follows
(paths relative to /root/reference/src).  Vectorised NumPy instead of the reference's Python
double loops; arithmetic order and dtypes follow the reference so results are comparable bit for
bit where the reference's own arithmetic is exact (argmax index, max value, quarter-pixel offsets).

Third-party arithmetic on this path that is NOT under /root/reference:
  * NumPy (pinned 1.17.2, environment.yml:182; 2.3.5 here): np.argmax / np.amax / np.sign.
  * OpenCV (pinned 3.4.2 / 4.2.0.34, environment.yml:186,353; 4.13.0 here): cv2.getAffineTransform,
    the exact 3-point affine solve in float64.  Restated here as a float64 linear solve.
"""
import numpy as np

FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]  # CONSTANTS.py:65


def get_max_preds(heatmaps):
    """lib/pose_parsing.py:16-55 (get_max_preds_hrnet)."""
    n = heatmaps.shape[0]
    if n == 0:
        return [], []
    j, w = heatmaps.shape[1], heatmaps.shape[3]
    flat = heatmaps.reshape(n, j, -1)
    idx = np.argmax(flat, 2)                       # first index on ties (:40)
    maxvals = np.amax(flat, 2).reshape(n, j, 1)    # (:41,43)
    preds = np.empty((n, j, 2), np.float32)
    preds[:, :, 0] = (idx % w).astype(np.float32)  # (:47)
    preds[:, :, 1] = np.floor(idx.astype(np.float32) / w)  # (:48)
    preds *= (maxvals > 0.0).astype(np.float32)    # (:50-53)
    return preds, maxvals


def _third_point(a, b):
    """lib/transforms.py:243-246."""
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def inverse_affine(center, scale, out_w, out_h):
    """lib/transforms.py:197-233 with rot=0, shift=0, inv=1: the 2x3 float64 matrix mapping heatmap -> image.

    Points are assembled in float32 exactly as the reference does (:213-224), then the 3-point affine
    is solved in float64 as cv2.getAffineTransform does (:229).
    """
    center = np.asarray(center)
    scale = np.asarray(scale)
    scale_tmp = scale * 200.0
    src_w = scale_tmp[0]                                          # only scale[0] is used (:209)
    src_dir = np.array([0.0, src_w * -0.5])                       # get_dir with rot_rad=0 (:249-256)
    dst_dir = np.array([0, out_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + src_dir
    dst[0, :] = [out_w * 0.5, out_h * 0.5]
    dst[1, :] = np.array([out_w * 0.5, out_h * 0.5]) + dst_dir
    src[2, :] = _third_point(src[0, :], src[1, :])
    dst[2, :] = _third_point(dst[0, :], dst[1, :])
    # inv=1: transform maps dst (heatmap) -> src (image)
    a = np.concatenate([dst.astype(np.float64), np.ones((3, 1))], axis=1)  # 3x3
    m = np.linalg.solve(a, src.astype(np.float64))                         # 3x2
    return m.T                                                             # 2x3


def transform_preds(coords, center, scale, out_w, out_h):
    """lib/transforms.py:184-194 + affine_transform :236-240 (float64 result)."""
    t = inverse_affine(center, scale, out_w, out_h)
    pts = np.concatenate([coords[:, 0:2].astype(np.float64), np.ones((coords.shape[0], 1))], axis=1)
    return pts @ t.T


def get_final_preds(heatmaps, center, scale):
    """lib/pose_parsing.py:58-92 (get_final_preds_hrnet) -> (preds, maxvals, coords)."""
    coords, maxvals = get_max_preds(heatmaps)
    n, j, h, w = heatmaps.shape
    px = np.floor(coords[:, :, 0] + 0.5).astype(np.int64)         # (:73)
    py = np.floor(coords[:, :, 1] + 0.5).astype(np.int64)
    inside = (px > 1) & (px < w - 1) & (py > 1) & (py < h - 1)     # (:75)
    pxc = np.clip(px, 1, w - 2)
    pyc = np.clip(py, 1, h - 2)
    ni, ji = np.meshgrid(np.arange(n), np.arange(j), indexing="ij")
    dx = heatmaps[ni, ji, pyc, pxc + 1] - heatmaps[ni, ji, pyc, pxc - 1]   # (:78)
    dy = heatmaps[ni, ji, pyc + 1, pxc] - heatmaps[ni, ji, pyc - 1, pxc]   # (:79)
    off = np.stack([np.sign(dx), np.sign(dy)], axis=2) * 0.25             # (:82)
    coords = coords.copy()
    coords = np.where(inside[:, :, None], (coords + off).astype(np.float32), coords)
    preds = coords.copy()
    for i in range(n):                                                    # (:87-90)
        preds[i] = transform_preds(coords[i], center[i], scale[i], w, h)
    return preds, maxvals, coords


def flip_back(output_flipped, matched_parts=FLIP_PAIRS):
    """lib/transforms.py:147-164: reverse W, swap left/right joint channels."""
    assert output_flipped.ndim == 4
    out = output_flipped[:, :, :, ::-1].copy()
    for a, b in matched_parts:
        tmp = out[:, a].copy()
        out[:, a] = out[:, b]
        out[:, b] = tmp
    return out


def flip_average(output, output_flipped, matched_parts=FLIP_PAIRS):
    """lib/inference.py:21-26: flip_back, unconditional 1-px right shift (col 0 kept), average."""
    of = flip_back(output_flipped, matched_parts)
    shifted = of.copy()
    shifted[:, :, :, 1:] = of[:, :, :, 0:-1]                              # (:25)
    return ((output + shifted) * np.float32(0.5)).astype(output.dtype)    # (:26)


def person_mse_loss(output, target, target_weight):
    """lib/loss.py:61-94 (PersonMSELoss): returns (loss, dloss/doutput), float64 accumulation.

    loss = 0.5/(J*B*hw) * sum (tw*(out-tgt))^2 ; the use_target_weight flag is ignored upstream (:71).
    """
    b, j = output.shape[0], output.shape[1]
    o = output.reshape(b, j, -1).astype(np.float64)
    t = target.reshape(b, j, -1).astype(np.float64)
    tw = target_weight.reshape(b, j, 1).astype(np.float64)
    d = (o - t) * tw
    denom = j * b * o.shape[2]
    loss = 0.5 * np.sum(d * d) / denom
    grad = (d * tw / denom).reshape(output.shape)
    return loss, grad


def synth_boxes(n, seed=0):
    """Synthetic person boxes as SURVEY.md 8(d) config 1: center, scale as _xywh2cs would make them
    (data/HRNet_Coco.py:233-248): scale = (0.75*h, h)/200*1.25."""
    rng = np.random.default_rng(seed)
    center = np.stack([rng.uniform(100, 500, n), rng.uniform(100, 400, n)], axis=1)
    h = rng.uniform(80, 400, n)
    scale = np.stack([0.75 * h, h], axis=1) / 200.0 * 1.25
    return center, scale


def blob_heatmaps(n, j, h, w, seed=0, sigma=2.0, noise=0.01):
    """Gaussian-blob heatmaps (data/JointsDataset.py:248-281 style targets) + small noise."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(-2, w + 2, (n, j, 1, 1))
    cy = rng.uniform(-2, h + 2, (n, j, 1, 1))
    ys, xs = np.mgrid[0:h, 0:w]
    g = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sigma ** 2))
    g = g + noise * rng.standard_normal((n, j, h, w))
    return g.astype(np.float32)
