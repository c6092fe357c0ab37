"""Drop-in for the hot-path part of the reference ``lib.pose_parsing`` (/root/reference/src/lib/pose_parsing.py:16-92).

Inputs may be NumPy arrays (as the reference's callers pass, 03_evaluate.py:151) or CUDA tensors (skips the
host round trip).  Outputs are NumPy float32 arrays with the reference's shapes, or CUDA tensors when
``as_tensor=True``.
"""
import numpy as np
import torch

from . import _lib
from .transforms import _as_cuda_f32, _pairs_array


def _decode(heat, center, scale, refine, heat_flipped=None, pairs=(), want_avg=False):
    heat = _as_cuda_f32(heat)
    B, J, h, w = heat.shape
    dev = heat.device
    maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
    coords = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
    preds = c = s = avg = None
    if center is not None:
        c = _as_cuda_f32(np.asarray(center, dtype=np.float64) if not torch.is_tensor(center) else center, dev)
        s = _as_cuda_f32(np.asarray(scale, dtype=np.float64) if not torch.is_tensor(scale) else scale, dev)
        if c.shape != (B, 2) or s.shape != (B, 2):
            raise ValueError(f"center/scale must be [{B},2], got {tuple(c.shape)} / {tuple(s.shape)}")
        preds = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
    hf = None
    if heat_flipped is not None:
        hf = _as_cuda_f32(heat_flipped, dev)
        if want_avg:
            avg = torch.empty_like(heat)
    pa, n_pairs = _pairs_array(pairs)
    if B > 0:
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().stl_decode(_lib.ptr(heat), _lib.ptr(hf), _lib.ptr(c), _lib.ptr(s), B, J, h, w, pa,
                                             n_pairs, int(refine), _lib.ptr(avg), _lib.ptr(preds), _lib.ptr(maxvals),
                                             _lib.ptr(coords), _lib.current_stream()))
    return preds, maxvals, coords, avg


def get_max_preds_hrnet(scaled_heats, thr=0.1, as_tensor=False):
    """pose_parsing.py:16-55: per-joint argmax -> (preds [N,J,2], maxvals [N,J,1]); ``thr`` is unused upstream."""
    if scaled_heats.shape[0] == 0:
        return [], []
    _, maxvals, coords, _ = _decode(scaled_heats, None, None, refine=False)
    if as_tensor:
        return coords, maxvals
    return coords.cpu().numpy(), maxvals.cpu().numpy()


def get_final_preds_hrnet(heatmaps, center, scale, as_tensor=False):
    """pose_parsing.py:58-92 -> (preds image-space [N,J,2], maxvals [N,J,1], coords heatmap-space [N,J,2])."""
    preds, maxvals, coords, _ = _decode(heatmaps, center, scale, refine=True)
    if as_tensor:
        return preds, maxvals, coords
    return preds.cpu().numpy(), maxvals.cpu().numpy(), coords.cpu().numpy()


def get_max_preds_upsampled(heatmaps, size=(256, 192), as_tensor=False):
    """get_max_preds_hrnet(F.interpolate(heatmaps, size, mode="bilinear", align_corners=True)) without materialising the
    upsampled tensor (pose_parsing.py:143-144).  -> (preds [N,J,2], maxvals [N,J,1])."""
    heat = _as_cuda_f32(heatmaps)
    B, J, h, w = heat.shape
    if B == 0:
        return [], []
    coords = torch.empty((B, J, 2), dtype=torch.float32, device=heat.device)
    maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=heat.device)
    with torch.cuda.device(heat.device):
        _lib.check(_lib.lib().stl_upsampled_argmax(_lib.ptr(heat), B, J, h, w, int(size[0]), int(size[1]),
                                                   _lib.ptr(coords), _lib.ptr(maxvals), _lib.current_stream()))
    if as_tensor:
        return coords, maxvals
    return coords.cpu().numpy(), maxvals.cpu().numpy()


def create_pose_entries(keypoints, max_vals=None, thr=0.1):
    """pose_parsing.py:107-135: (pose_entries list of 19-vectors, all_keypoints [N*17,4]) from per-person keypoints;
    keypoints below the confidence threshold get visibility 0.  Small host-side bookkeeping, NumPy like the reference."""
    if len(keypoints) == 0:
        all_keypoints = []
    else:
        kp = np.asarray(keypoints)
        all_keypoints = np.concatenate([kp.reshape(-1, kp.shape[-1]), np.ones((kp.shape[0] * kp.shape[1], 2), kp.dtype)], axis=1)
        rows = np.argwhere(all_keypoints == -1)[:, 0]
        all_keypoints[rows, :] = -1
        if max_vals is not None:
            low = np.argwhere(np.asarray(max_vals)[:, :, 0] < thr)
            all_keypoints[low[:, 0] * 17 + low[:, 1], -1] = 0
    pose_entries = []
    for idx, cur_pose in enumerate(keypoints):
        entry = np.ones(19) * -1
        for i, kpt in enumerate(cur_pose):
            if kpt[0] != -1:
                entry[i] = 17 * idx + i
        entry[-2] = len(np.where(entry[:-2] != -1)[0])
        pose_entries.append(entry)
    return pose_entries, all_keypoints


def create_pose_from_outputs(dets, keypoint_thr=0.1):
    """pose_parsing.py:138-151: upsample to 256x192, arg-max, pose entries; all_keypoints columns (y, x, 1, visible)."""
    keypoint_coords, max_vals = get_max_preds_upsampled(dets, (256, 192))
    pose_entries, all_keypoints = create_pose_entries(keypoint_coords, max_vals, thr=keypoint_thr)
    all_keypoints = np.array([all_keypoints[:, 1], all_keypoints[:, 0], all_keypoints[:, 2], all_keypoints[:, 3]]).T
    return pose_entries, all_keypoints
