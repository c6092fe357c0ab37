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
