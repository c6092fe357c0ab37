"""Drop-in for ``JointsDataset.generate_target`` (/root/reference/src/data/JointsDataset.py:230-286), batched on the device.

The reference builds the [17,64,48] Gaussian target of every sample in the data-loader workers and ships 209 KB per crop
to the GPU; here the joints (408 B per crop) are shipped and the targets are generated where the loss consumes them.
"""
import numpy as np
import torch

from . import _lib


def generate_target(joints, joints_vis, image_size=(192, 256), heatmap_size=(48, 64), sigma=2, joints_weight=None,
                    device="cuda"):
    """joints, joints_vis: [B,J,3] (or [J,3] for one sample, like the reference method) -> (target f32 [B,J,h,w],
    target_weight f32 [B,J,1]) as CUDA tensors.  image_size / heatmap_size are (width, height) like the reference's
    ``self.image_size`` / ``self.heatmap_size``; joints_weight: [J] or [J,1] per-joint weights or None."""
    j = torch.as_tensor(np.asarray(joints, dtype=np.float64) if not torch.is_tensor(joints) else joints)
    v = torch.as_tensor(np.asarray(joints_vis, dtype=np.float64) if not torch.is_tensor(joints_vis) else joints_vis)
    single = j.dim() == 2
    if single:
        j, v = j[None], v[None]
    j = j.to(device=device, dtype=torch.float64).contiguous()
    v = v.to(device=device, dtype=torch.float64).contiguous()
    B, J = j.shape[0], j.shape[1]
    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    jw = None
    if joints_weight is not None:
        jw = torch.as_tensor(np.asarray(joints_weight, dtype=np.float32)).reshape(-1).to(j.device).contiguous()
        if jw.numel() != J:
            raise ValueError(f"joints_weight must have {J} entries")
    target = torch.empty((B, J, h, w), dtype=torch.float32, device=j.device)
    weight = torch.empty((B, J, 1), dtype=torch.float32, device=j.device)
    if B > 0:
        with torch.cuda.device(j.device):
            _lib.check(_lib.lib().stl_generate_target(_lib.ptr(j), _lib.ptr(v), _lib.ptr(jw), B, J, h, w, int(image_size[1]),
                                                      int(image_size[0]), int(sigma), _lib.ptr(target), _lib.ptr(weight),
                                                      _lib.current_stream()))
    return (target[0], weight[0]) if single else (target, weight)
