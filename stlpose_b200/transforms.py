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
