"""Drop-in for the hot-loop part of the reference ``lib.metrics`` (/root/reference/src/lib/metrics.py:268-364): the PCK
accuracy that 02_train.py:223 / :277 and 03_evaluate.py:142 compute every iteration on ``output.cpu()``.

Here the two arg-max decodes and the distance / threshold statistics run on the device; with ``as_tensor=True`` nothing
is copied to the host (no synchronisation inside the training loop).
"""
import numpy as np
import torch

from . import _lib
from .pose_parsing import _decode


def accuracy(output, target, hm_type="gaussian", thr=0.5, as_tensor=False):
    """-> (acc [J+1], avg_acc, cnt, pred [B,J,2]); NumPy / Python scalars like the reference, or CUDA tensors.

    The reference's lines 355-356 are corrupted (``acc[i , avg_acc, cnt, pred\\n+ 1] = ...``); this follows the evident
    intent, ``acc[i + 1] = dist_acc(dists[idx[i]])`` (upstream HRNet).  ``dist_acc`` is called without ``thr`` there, so
    the threshold is always its default 0.5 upstream; here ``thr`` is honoured.
    """
    if hm_type != "gaussian":
        raise NotImplementedError("only hm_type='gaussian' (arg-max of heatmaps) exists on the device path")
    _, _, pred, _ = _decode(output, None, None, refine=False)
    _, _, tgt, _ = _decode(target, None, None, refine=False)
    B, J = pred.shape[0], pred.shape[1]
    h, w = output.shape[2], output.shape[3]
    dev = pred.device
    acc = torch.empty(J + 1, dtype=torch.float32, device=dev)
    avg = torch.empty((), dtype=torch.float32, device=dev)
    cnt = torch.empty((), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().stl_pck_accuracy(_lib.ptr(pred), _lib.ptr(tgt), B, J, h, w, float(thr), _lib.ptr(acc),
                                               _lib.ptr(avg), _lib.ptr(cnt), _lib.current_stream()))
    if as_tensor:
        return acc, avg, cnt, pred
    return acc.cpu().numpy().astype(np.float64), float(avg.item()), int(cnt.item()), pred.cpu().numpy()
