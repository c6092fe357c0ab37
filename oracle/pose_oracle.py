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


# ----------------------------------------------------------------------------------------------
# crop extraction (the step before the hot path): lib/transforms.py:14-82, 197-233, 259-268
# ----------------------------------------------------------------------------------------------
def coords2cs(coords, det_width=192, det_height=256):
    """TransformDetection._coords2cs, lib/transforms.py:60-82: (xmin,ymin,xmax,ymax) -> (center f32[2], scale f32[2])."""
    xmin, ymin, xmax, ymax = coords
    aspect = det_width * 1.0 / det_height
    w, h = (xmax - xmin), (ymax - ymin)
    center = np.zeros((2), dtype=np.float32)
    center[0] = xmin + w * 0.5
    center[1] = ymin + h * 0.5
    if w > aspect * h:
        h = w * 1.0 / aspect
    elif w < aspect * h:
        w = h * aspect
    scale = np.array([w * 1.0 / 200, h * 1.0 / 200], dtype=np.float32)
    if center[0] != -1:
        scale = scale * 1.25
    return center, scale


def cv_affine_solve(src, dst):
    """cv2.getAffineTransform(src, dst): the 6x6 system [x y 1 0 0 0; 0 0 0 x y 1] m = [u; v] solved by OpenCV's own
    LU with partial pivoting (core/matrix_decomp.cpp LUImpl, float64), restated operation by operation so that the
    matrix is bit-identical (checked against cv2 in tests/test_oracle_vs_reference.py) -- a last-bit difference in
    the matrix moves 1/32-pixel source coordinates across rounding boundaries in warp_affine_u8."""
    src = np.asarray(src, np.float64)
    dst = np.asarray(dst, np.float64)
    a = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        a[2 * i, 0:3] = [src[i, 0], src[i, 1], 1.0]
        a[2 * i + 1, 3:6] = [src[i, 0], src[i, 1], 1.0]
        b[2 * i], b[2 * i + 1] = dst[i, 0], dst[i, 1]
    n = 6
    for i in range(n):
        k = i
        for j in range(i + 1, n):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if k != i:
            a[[i, k], i:] = a[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / a[i, i]
        for j in range(i + 1, n):
            alpha = a[j, i] * d
            for kk in range(i + 1, n):
                a[j, kk] += alpha * a[i, kk]
            b[j] += alpha * b[i]
    for i in range(n - 1, -1, -1):
        sacc = b[i]
        for kk in range(i + 1, n):
            sacc -= a[i, kk] * b[kk]
        b[i] = sacc / a[i, i]
    return b.reshape(2, 3)


def forward_affine(center, scale, rot, output_size):
    """get_affine_transform(center, scale, rot, output_size) with inv=0, lib/transforms.py:197-233: the 2x3 float64
    matrix image -> crop.  Points assembled in float32 as the reference does, 3-point solve in float64
    (cv2.getAffineTransform, OpenCV 3.4.2 / 4.2.0 pinned in environment.yml; any version: plain 6x6 solve)."""
    center = np.asarray(center)
    scale = np.asarray(scale)
    scale_tmp = scale * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = [0, src_w * -0.5]
    src_dir = [p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs]                     # get_dir :249-256
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + src_dir
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2, :] = _third_point(src[0, :], src[1, :])
    dst[2, :] = _third_point(dst[0, :], dst[1, :])
    return cv_affine_solve(src, dst)


def invert_affine(m):
    """cv2.warpAffine's inversion of the 2x3 matrix (dst -> src), float64 (imgwarp.cpp, invertAffineTransform inline)."""
    m = np.asarray(m, np.float64).reshape(6).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_affine_u8(img, m, dsize):
    """cv2.warpAffine(img, m, dsize, flags=INTER_LINEAR) for uint8 HWC images, borderMode CONSTANT 0: OpenCV's fixed-point
    algorithm (imgwarp.cpp WarpAffineInvoker + remap bilinear): source coordinates in 1/1024 px, rounded to 1/32 px
    (INTER_BITS = 5), bilinear weights (32-fx)(32-fy)/1024 in 15-bit fixed point, result (sum + 2^14) >> 15.
    Bit-exact against cv2 4.13 in this container (tests/test_oracle_vs_reference.py)."""
    W, H = dsize
    mi = invert_affine(m)
    ab = 1024
    xs = np.arange(W)
    adelta = np.rint(mi[0] * xs * ab).astype(np.int64)
    bdelta = np.rint(mi[3] * xs * ab).astype(np.int64)
    ih, iw, ch = img.shape
    pad = np.zeros((ih + 2, iw + 2, ch), np.int64)
    pad[1:-1, 1:-1] = img
    out = np.zeros((H, W, ch), np.uint8)

    def px(yy, xx):
        ok = (yy >= 0) & (yy < ih) & (xx >= 0) & (xx < iw)
        return pad[np.clip(yy, -1, ih) + 1, np.clip(xx, -1, iw) + 1] * ok[:, None]

    for y in range(H):
        x0 = int(np.rint((mi[1] * y + mi[2]) * ab)) + 16
        y0 = int(np.rint((mi[4] * y + mi[5]) * ab)) + 16
        X, Y = (x0 + adelta) >> 5, (y0 + bdelta) >> 5
        sx, sy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
        acc = (px(sy, sx) * ((32 - fx) * (32 - fy))[:, None] + px(sy, sx + 1) * (fx * (32 - fy))[:, None] +
               px(sy + 1, sx) * ((32 - fx) * fy)[:, None] + px(sy + 1, sx + 1) * (fx * fy)[:, None])
        out[y] = ((acc * 32 + 16384) >> 15).astype(np.uint8)
    return out


def warp_affine_f32(img, m, dsize):
    """cv2.warpAffine(img, m, dsize, flags=INTER_LINEAR) for float32 HWC images (what 04_evaluate_vases_qualitatively.py:
    209-213 feeds TransformDetection): the same fixed-point source coordinates as the uint8 path (1/32 px), but the
    interpolation runs in float32 with OpenCV's bilinear table (imgwarp.cpp initInterTab2D: w = vy[k1] * vx[k2] with
    v = {1 - i/32, i/32}) and the sum S00 w00 + S01 w01 + S10 w10 + S11 w11 evaluated left to right.
    Bit-exact against cv2 4.13 in this container (tests/test_oracle_vs_reference.py)."""
    W, H = dsize
    mi = invert_affine(m)
    ab = 1024
    xs = np.arange(W)
    adelta = np.rint(mi[0] * xs * ab).astype(np.int64)
    bdelta = np.rint(mi[3] * xs * ab).astype(np.int64)
    img = np.asarray(img, np.float32)
    ih, iw, ch = img.shape
    pad = np.zeros((ih + 2, iw + 2, ch), np.float32)
    pad[1:-1, 1:-1] = img
    out = np.zeros((H, W, ch), np.float32)
    tab = np.arange(32, dtype=np.float32) / np.float32(32)

    def px(yy, xx):
        ok = (yy >= 0) & (yy < ih) & (xx >= 0) & (xx < iw)
        return pad[np.clip(yy, -1, ih) + 1, np.clip(xx, -1, iw) + 1] * ok[:, None].astype(np.float32)

    for y in range(H):
        x0 = int(np.rint((mi[1] * y + mi[2]) * ab)) + 16
        y0 = int(np.rint((mi[4] * y + mi[5]) * ab)) + 16
        X, Y = (x0 + adelta) >> 5, (y0 + bdelta) >> 5
        sx, sy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
        wx1, wy1 = tab[fx], tab[fy]
        wx0, wy0 = np.float32(1) - wx1, np.float32(1) - wy1
        out[y] = (px(sy, sx) * (wy0 * wx0)[:, None] + px(sy, sx + 1) * (wy0 * wx1)[:, None] +
                  px(sy + 1, sx) * (wy1 * wx0)[:, None] + px(sy + 1, sx + 1) * (wy1 * wx1)[:, None])
    return out


def transform_detection(img, list_coords, det_width=192, det_height=256):
    """TransformDetection.__call__, lib/transforms.py:30-58 -> (detections u8 [N,3,H,W], centers [N,2], scales [N,2])."""
    dets, centers, scales = [], [], []
    for coords in list_coords:
        c, s = coords2cs(coords, det_width, det_height)
        m = forward_affine(c, s, 0, (det_width, det_height))
        warp = warp_affine_u8 if np.asarray(img).dtype == np.uint8 else warp_affine_f32
        dets.append(warp(img, m, (det_width, det_height)))
        centers.append(c)
        scales.append(s)
    dets, centers, scales = np.array(dets), np.array(centers), np.array(scales)
    if len(dets) == 0:
        return dets, centers, scales
    return dets.transpose(0, 3, 1, 2), centers, scales


# ----------------------------------------------------------------------------------------------
# training targets: data/JointsDataset.py:230-286
# ----------------------------------------------------------------------------------------------
def generate_target(joints, joints_vis, image_size=(192, 256), heatmap_size=(48, 64), sigma=2, joints_weight=None):
    """JointsDataset.generate_target for one sample: joints, joints_vis [J,3] -> (target f32 [J,h,w], target_weight f32
    [J,1]).  A 13x13 (sigma 2) unnormalised Gaussian centred on the rounded heatmap position of every visible joint
    whose patch touches the map; joints whose patch lies outside get weight 0."""
    J = joints.shape[0]
    W, H = heatmap_size
    weight = np.ones((J, 1), dtype=np.float32)
    weight[:, 0] = joints_vis[:, 0]
    target = np.zeros((J, H, W), dtype=np.float32)
    rad = sigma * 3
    stride = np.asarray(image_size) / np.asarray(heatmap_size)
    ax = np.arange(0, 2 * rad + 1, 1, np.float32)
    gauss = np.exp(-((ax - rad) ** 2 + (ax[:, None] - rad) ** 2) / (2 * sigma ** 2))
    for j in range(J):
        mx = int(joints[j][0] / stride[0] + 0.5)
        my = int(joints[j][1] / stride[1] + 0.5)
        x0, y0, x1, y1 = int(mx - rad), int(my - rad), int(mx + rad + 1), int(my + rad + 1)
        if x0 >= W or y0 >= H or x1 < 0 or y1 < 0:
            weight[j] = 0
            continue
        if weight[j] > 0.5:
            xa, xb, ya, yb = max(0, x0), min(x1, W), max(0, y0), min(y1, H)
            target[j, ya:yb, xa:xb] = gauss[ya - y0:yb - y0, xa - x0:xb - x0]
    if joints_weight is not None:
        weight = np.multiply(weight, joints_weight)
    return target, weight


# ----------------------------------------------------------------------------------------------
# second decode path: lib/pose_parsing.py:107-151
# ----------------------------------------------------------------------------------------------
def create_pose_entries(keypoints, max_vals=None, thr=0.1):
    """pose_parsing.py:107-135."""
    if len(keypoints) == 0:
        all_keypoints = []
    else:
        all_keypoints = np.array([(*item, 1, 1) for sublist in keypoints for item in sublist])
        idx = np.argwhere(all_keypoints == -1)
        all_keypoints[idx[:, 0], :] = -1
        if max_vals is not None:
            idx = np.argwhere(max_vals[:, :, 0] < thr)
            all_keypoints[idx[:, 0] * 17 + idx[:, 1], -1] = 0
    pose_entries = []
    for idx, cur_pose in enumerate(keypoints):
        entry = np.ones(19) * -1
        for i, kpt in enumerate(cur_pose):
            if kpt[0] != -1:
                entry[i] = 17 * idx + i
        entry[-2] = len(np.where(entry[:-2] != -1)[0])
        pose_entries.append(entry)
    return pose_entries, all_keypoints


def upsampled_max_preds(dets, size=(256, 192)):
    """pose_parsing.py:143-144: F.interpolate(dets, size, bilinear, align_corners=True) (torch, third-party: the same
    library call the reference makes) -> get_max_preds."""
    import torch
    import torch.nn.functional as F
    scaled = F.interpolate(torch.as_tensor(dets).clone(), size, mode="bilinear", align_corners=True)
    return get_max_preds(scaled.numpy())


def create_pose_from_outputs(dets, keypoint_thr=0.1):
    """pose_parsing.py:138-151."""
    coords, max_vals = upsampled_max_preds(dets)
    pose_entries, all_keypoints = create_pose_entries(coords, max_vals, thr=keypoint_thr)
    all_keypoints = np.array([all_keypoints[:, 1], all_keypoints[:, 0], all_keypoints[:, 2], all_keypoints[:, 3]]).T
    return pose_entries, all_keypoints


def calc_dists(preds, target, normalize):
    """metrics.py:268-296: normalised distance per (joint, sample); -1 where the target is not > 1 in x and y."""
    preds = preds.astype(np.float32)
    target = target.astype(np.float32)
    dists = np.zeros((preds.shape[1], preds.shape[0]))
    for n in range(preds.shape[0]):
        for c in range(preds.shape[1]):
            if target[n, c, 0] > 1 and target[n, c, 1] > 1:
                d = preds[n, c, :] / normalize[n] - target[n, c, :] / normalize[n]
                dists[c, n] = np.sqrt((d * d).sum())
            else:
                dists[c, n] = -1
    return dists


def dist_acc(dists, thr=0.5):
    """metrics.py:299-317: fraction of the counted distances below thr, -1 when nothing is counted."""
    counted = dists != -1
    n = counted.sum()
    return (dists[counted] < thr).sum() * 1.0 / n if n > 0 else -1


def accuracy(output, target, thr=0.5):
    """metrics.py:321-364 (hm_type='gaussian').  Lines 355-356 of the reference are corrupted; the evident intent
    (upstream HRNet) is ``acc[i + 1] = dist_acc(dists[idx[i]])``.  -> (acc [J+1], avg_acc, cnt, pred)."""
    pred, _ = get_max_preds(output)
    tgt, _ = get_max_preds(target)
    h, w = output.shape[2], output.shape[3]
    norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10
    dists = calc_dists(pred, tgt, norm)
    J = output.shape[1]
    acc = np.zeros(J + 1)
    avg_acc, cnt = 0, 0
    for i in range(J):
        acc[i + 1] = dist_acc(dists[i], thr)
        if acc[i + 1] >= 0:
            avg_acc += acc[i + 1]
            cnt += 1
    avg_acc = avg_acc / cnt if cnt != 0 else 0
    if cnt != 0:
        acc[0] = avg_acc
    return acc, avg_acc, cnt, pred


COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def oks_iou(g, d, a_g, a_d, sigmas=None, in_vis_thre=None):
    """lib/nms.py:49-74: OKS of one person g [3J] against persons d [n,3J] (x, y, v interleaved).  Vectorised over the
    persons; dtypes as the reference (float32 keypoints: differences and squares in float32, the divisions in float64).
    The reference's mask ``list(vg > t) and list(vd > t)`` evaluates to the SECOND list, i.e. the candidate's mask."""
    sig = sigmas if isinstance(sigmas, np.ndarray) else COCO_SIGMAS
    var = (sig * 2) ** 2
    ious = np.zeros(d.shape[0])
    dx = d[:, 0::3] - g[0::3]
    dy = d[:, 1::3] - g[1::3]
    e = (dx ** 2 + dy ** 2) / var / ((a_g + a_d) / 2 + np.spacing(1))[:, None] / 2
    for n in range(d.shape[0]):
        en = e[n][d[n, 2::3] > in_vis_thre] if in_vis_thre is not None else e[n]
        ious[n] = np.sum(np.exp(-np.ascontiguousarray(en))) / en.shape[0] if en.shape[0] != 0 else 0.0
    return ious


def oks_nms(kpts, scores, areas, thresh, sigmas=None, in_vis_thre=None):
    """lib/nms.py:10-46 on arrays: kpts [n,J,3], scores [n], areas [n] -> indices to keep, best first."""
    if len(kpts) == 0:
        return []
    flat = np.asarray(kpts).reshape(len(kpts), -1)
    order = np.asarray(scores).argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        ovr = oks_iou(flat[i], flat[order[1:]], areas[i], areas[order[1:]], sigmas, in_vis_thre)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep


def rescore(all_preds, box_scores, in_vis_thr=0.2):
    """lib/metrics.py:239-250: score = mean of the joint scores above in_vis_thr (float32 running sum in joint order,
    0 when there is none) times the box score (float64)."""
    out = np.zeros(len(all_preds))
    for m, kpt in enumerate(all_preds):
        s, n = 0, 0
        for j in range(kpt.shape[0]):
            if kpt[j][2] > in_vis_thr:
                s = s + kpt[j][2]
                n += 1
        if n:
            s = s / n
        out[m] = s * box_scores[m]
    return out


def rescore_and_nms(all_preds, all_bboxes, image_ids, in_vis_thr=0.2, oks_thr=0.9):
    """lib/metrics.py:211-258 (generate_submission_hrnet up to the JSON packing): group by image in first-appearance
    order, rescore, OKS-NMS.  -> list per image of (person index m, rescored score), best first."""
    groups = {}
    for m, img in enumerate(image_ids):
        groups.setdefault(img, []).append(m)
    scores = rescore(all_preds, all_bboxes[:, 5], in_vis_thr)
    out = []
    for img, ms in groups.items():
        ms = np.array(ms)
        keep = oks_nms(all_preds[ms], scores[ms], all_bboxes[ms, 4], oks_thr)
        out.append([(int(ms[k]), scores[ms[k]]) for k in keep])
    return out


def coco_results(all_preds, all_bboxes, image_ids, kept):
    """data/data_processing.py:52-82 (convert_keypoints_to_coco_format) on the output of rescore_and_nms: one result dict
    per kept person, 'keypoints' = 51 float64 values (x, y, score per joint)."""
    results = []
    for persons in kept:
        for m, score in persons:
            results.append({"image_id": image_ids[m], "category_id": 1,
                            "keypoints": list(all_preds[m].astype(np.float64).reshape(-1)), "score": score,
                            "center": list(all_bboxes[m][0:2]), "scale": list(all_bboxes[m][2:4])})
    return results


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
