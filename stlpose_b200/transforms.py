"""Drop-in for the hot-path part of the reference ``lib.transforms`` (/root/reference/src/lib/transforms.py:147-164)."""
import ctypes

import torch

from . import _lib

FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]  # CONSTANTS.py:65


def _pairs_array(matched_parts):
    flat = [int(v) for pair in matched_parts for v in pair]
    return (ctypes.c_int * max(1, len(flat)))(*flat), len(flat) // 2


def _as_cuda_f32(a, device=None):
    """numpy array / CPU tensor / CUDA tensor -> contiguous fp32 CUDA tensor."""
    t = a.detach() if torch.is_tensor(a) else torch.as_tensor(a)
    if not t.is_cuda:
        t = t.to(device or "cuda")
    return t.float().contiguous()


def flip_back(output_flipped, matched_parts):
    """Reverse W and swap left/right joint channels.  Returns a CPU tensor, like the reference.

    (forward_pass does not call this: it fuses the permutation into stl_flip_avg / stl_decode.)
    """
    assert output_flipped.ndim == 4, 'output_flipped should be [batch_size, num_joints, height, width]'
    x = _as_cuda_f32(output_flipped)
    out = torch.empty_like(x)
    B, J, h, w = x.shape
    pairs, n_pairs = _pairs_array(matched_parts)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().stl_flip_back(_lib.ptr(x), _lib.ptr(out), B, J, h, w, pairs, n_pairs,
                                            _lib.current_stream()))
    return out.cpu()


# ----------------------------------------------------------------------------------------------------------------
# Crop extraction (the step in front of the network): lib/transforms.py:14-82, 197-268
# ----------------------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402


def _solve_affine_3pt(src, dst):
    """The 2x3 float64 matrix with dst_i = M [src_i; 1] for three point pairs, solved the way cv2.getAffineTransform
    does (6x6 system, Gaussian elimination with partial pivoting in float64, same operation order), so that the
    1/32-pixel source coordinates derived from it round exactly like the reference's."""
    a = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        a[2 * i, 0], a[2 * i, 1], a[2 * i, 2] = float(src[i][0]), float(src[i][1]), 1.0
        a[2 * i + 1, 3], a[2 * i + 1, 4], a[2 * i + 1, 5] = float(src[i][0]), float(src[i][1]), 1.0
        b[2 * i], b[2 * i + 1] = float(dst[i][0]), float(dst[i][1])
    for i in range(6):
        piv = i + int(np.argmax(np.abs(a[i:, i])))          # first maximum, like the strict '>' scan
        if piv != i:
            a[[i, piv], i:] = a[[piv, i], i:]
            b[[i, piv]] = b[[piv, i]]
        d = -1.0 / a[i, i]
        for j in range(i + 1, 6):
            alpha = a[j, i] * d
            a[j, i + 1:] += alpha * a[i, i + 1:]
            b[j] += alpha * b[i]
    for i in range(5, -1, -1):
        acc = b[i]
        for k in range(i + 1, 6):
            acc -= a[i, k] * b[k]
        b[i] = acc / a[i, i]
    return b.reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """lib/transforms.py:197-233 -> 2x3 float64 matrix (image -> crop, or crop -> image with inv=1)."""
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])
    scale_tmp = np.asarray(scale) * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p0, p1 = 0, src_w * -0.5
    src_dir = [p0 * cs - p1 * sn, p0 * sn + p1 * cs]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    for pts in (src, dst):
        d = pts[0, :] - pts[1, :]
        pts[2, :] = pts[1, :] + np.array([-d[1], d[0]], dtype=np.float32)
    return _solve_affine_3pt(dst, src) if inv else _solve_affine_3pt(src, dst)


def _invert_for_warp(m):
    """The crop -> image matrix cv2.warpAffine derives from M (float64, same operation order)."""
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


def warp_affine_crops(img, matrices, output_size, normalize=None, as_tensor=False):
    """cv2.warpAffine(img, M, output_size, flags=INTER_LINEAR) for every M, on the device.

    img: uint8 or float32 HWC image (NumPy array or CUDA tensor); matrices: iterable of 2x3 image -> crop matrices;
    output_size = (width, height).  Returns uint8 crops [N,3,h,w] (NumPy, or CUDA tensor with ``as_tensor``), or, with
    ``normalize=(mean3, std3)``, the network input fp32 CUDA tensor [N,3,h,w] = (v/255 - mean) / std.  A float32 image
    (what 04_evaluate_vases_qualitatively.py:209-210 passes) gives float32 crops, interpolated like cv2 does for floats."""
    t = img if torch.is_tensor(img) else torch.as_tensor(np.ascontiguousarray(img))
    if t.dtype == torch.float64:
        t = t.float()
    if t.dtype not in (torch.uint8, torch.float32) or t.ndim != 3 or t.shape[2] != 3:
        raise ValueError("img must be a uint8 or float32 [H,W,3] image")
    t = t.cuda().contiguous() if not t.is_cuda else t.contiguous()
    if t.dtype == torch.float32:
        if normalize is not None:
            raise ValueError("normalize= fuses ToTensor (v / 255) and applies to uint8 images only")
        mats = [np.asarray(m, np.float64) for m in matrices]
        out_w, out_h = int(output_size[0]), int(output_size[1])
        out = torch.empty((len(mats), 3, out_h, out_w), dtype=torch.float32, device=t.device)
        if mats:
            minv = torch.as_tensor(np.stack([_invert_for_warp(m) for m in mats])).to(t.device)
            with torch.cuda.device(t.device):
                _lib.check(_lib.lib().stl_warp_affine_crops_f32(_lib.ptr(t), t.shape[0], t.shape[1], _lib.ptr(minv),
                                                                len(mats), out_h, out_w, _lib.ptr(out),
                                                                _lib.current_stream()))
        return out if as_tensor else out.cpu().numpy()
    mats = [np.asarray(m, np.float64) for m in matrices]
    n = len(mats)
    out_w, out_h = int(output_size[0]), int(output_size[1])
    dev = t.device
    if n == 0:
        dt = torch.float32 if normalize is not None else torch.uint8
        empty = torch.empty((0, 3, out_h, out_w), dtype=dt, device=dev)
        return empty if (as_tensor or normalize is not None) else empty.cpu().numpy()
    minv = torch.as_tensor(np.stack([_invert_for_warp(m) for m in mats])).to(dev)
    out_u8 = out_f = mean = std = None
    if normalize is None:
        out_u8 = torch.empty((n, 3, out_h, out_w), dtype=torch.uint8, device=dev)
    else:
        out_f = torch.empty((n, 3, out_h, out_w), dtype=torch.float32, device=dev)
        mean = (ctypes.c_float * 3)(*[float(v) for v in normalize[0]])
        std = (ctypes.c_float * 3)(*[float(v) for v in normalize[1]])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().stl_warp_affine_crops(_lib.ptr(t), t.shape[0], t.shape[1], _lib.ptr(minv), n, out_h, out_w,
                                                    _lib.ptr(out_u8), _lib.ptr(out_f), mean, std, _lib.current_stream()))
    if normalize is not None:
        return out_f
    return out_u8 if as_tensor else out_u8.cpu().numpy()


def crop(img, center, scale, output_size, rot=0):
    """lib/transforms.py:259-268 -> uint8 [h,w,3] like cv2.warpAffine returns."""
    m = get_affine_transform(center, scale, rot, output_size)
    return warp_affine_crops(img, [m], output_size)[0].transpose(1, 2, 0)


class TransformDetection:
    """lib/transforms.py:14-82: person boxes of one image -> aspect-corrected, 1.25x enlarged crops (+ center, scale).

    ``__call__`` returns what the reference returns (uint8 [N,3,H,W] NumPy array, centers, scales);
    ``extract_normalized`` returns the network input directly on the device (ToTensor + Normalize fused into the warp,
    no host round trip of the crops)."""

    def __init__(self, det_width=192, det_height=256):
        self.det_width = det_width
        self.det_height = det_height
        self.image_size = np.array([det_width, det_height])
        self.aspect_ratio = self.det_width * 1.0 / self.det_height
        self.pixel_std = 200

    def _coords2cs(self, coords):
        xmin, ymin, xmax, ymax = coords
        x, y = xmin, ymin
        w, h = (xmax - xmin), (ymax - ymin)
        center = np.zeros((2), dtype=np.float32)
        center[0] = x + w * 0.5
        center[1] = y + h * 0.5
        if w > self.aspect_ratio * h:
            h = w * 1.0 / self.aspect_ratio
        elif w < self.aspect_ratio * h:
            w = h * self.aspect_ratio
        scale = np.array([w * 1.0 / self.pixel_std, h * 1.0 / self.pixel_std], dtype=np.float32)
        if center[0] != -1:
            scale = scale * 1.25
        return center, scale

    def _matrices(self, list_coords):
        centers, scales, mats = [], [], []
        for coords in list_coords:
            c, s = self._coords2cs(coords)
            mats.append(get_affine_transform(center=c, scale=s, rot=0, output_size=self.image_size))
            centers.append(c)
            scales.append(s)
        return mats, np.array(centers), np.array(scales)

    def __call__(self, img, list_coords):
        mats, centers, scales = self._matrices(list_coords)
        if len(mats) == 0:
            return np.array([]), centers, scales
        dets = warp_affine_crops(img, mats, (int(self.image_size[0]), int(self.image_size[1])))
        return dets, centers, scales

    def extract_normalized(self, img, list_coords, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        """-> (fp32 CUDA tensor [N,3,H,W] ready for forward_pass, centers, scales); ImageNet statistics by default
        (data/data_loaders.py:59-61)."""
        mats, centers, scales = self._matrices(list_coords)
        x = warp_affine_crops(img, mats, (int(self.image_size[0]), int(self.image_size[1])), normalize=(mean, std))
        return x, centers, scales
