"""Drop-in for the reference ``lib.inference`` (/root/reference/src/lib/inference.py:11-32)."""
import torch

from . import _lib
from .hrnet import PoseHighResolutionNet
from .transforms import FLIP_PAIRS, _pairs_array


def _unwrap(model):
    # the reference wraps the model in torch.nn.DataParallel (02_train.py:109, 03_evaluate.py:100)
    return model.module if isinstance(model, torch.nn.DataParallel) else model


def forward_pass(model, img, model_name, device=None, flip=False):
    """Forward pass (+ optional flip test) -> heatmaps f32 [B,J,h,w] on ``img``'s device.

    Same signature and result as the reference.  With ``flip=True`` the reference runs the model twice and
    round-trips the flipped heatmaps through the CPU (flip_back, inference.py:23-24); here both passes run as
    one 2B batch and the flip-back / 1-px shift / average is a single device kernel.
    """
    if model_name != "HRNet":
        raise NotImplementedError("Wrong model name. Only ['HRNet'] supported")
    net = _unwrap(model)
    if not isinstance(net, PoseHighResolutionNet):
        raise TypeError("forward_pass expects stlpose_b200.PoseHighResolutionNet")
    if flip is not True:
        return net(img)
    both = net.forward_flip_pair(img)
    B = img.shape[0]
    out = torch.empty_like(both[:B])
    if B == 0:
        return out
    J, h, w = out.shape[1:]
    pairs, n_pairs = _pairs_array(FLIP_PAIRS)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().stl_flip_avg(_lib.ptr(both[:B]), _lib.ptr(both[B:]), _lib.ptr(out), B, J, h, w,
                                           pairs, n_pairs, _lib.current_stream()))
    return out
