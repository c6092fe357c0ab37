"""Batch sharding across the GPUs of one box: one process per GPU, no collective on the data path.

Replaces the reference's single-process ``torch.nn.DataParallel`` wrapper (02_train.py:109, 03_evaluate.py:100:
scatter -> replicate -> parallel_apply -> gather every forward).  Person crops are independent in inference, so
each rank runs the whole pipeline on a contiguous slice of the batch with its own resident copy of the weights;
the only exchange is the final all-gather of the keypoints (204 B per crop).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous slice [lo, hi) of a batch of n for `rank`, split like DataParallel.scatter / torch.chunk:
    chunks of ceil(n / world_size); trailing ranks may get a short or empty slice."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    chunk = -(-n // world_size) if n > 0 else 0
    lo = min(n, rank * chunk)
    hi = min(n, lo + chunk)
    return lo, hi


def gather_keypoints(preds_local, maxvals_local, n_total, group=None):
    """All-gather per-rank results ([n_r,J,2], [n_r,J,1]) back into batch order -> ([n_total,J,2], [n_total,J,1]).

    Works on whatever backend the process group uses (NCCL for CUDA tensors, gloo for CPU tensors)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return preds_local, maxvals_local
    J = preds_local.shape[1]
    chunk = -(-n_total // world) if n_total > 0 else 0
    packed = torch.zeros((chunk, J, 3), dtype=preds_local.dtype, device=preds_local.device)
    n_r = preds_local.shape[0]
    packed[:n_r, :, :2] = preds_local
    packed[:n_r, :, 2:] = maxvals_local
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        parts.append(out[r][: hi - lo])
    full = torch.cat(parts, dim=0)
    return full[..., :2].contiguous(), full[..., 2:].contiguous()


class ShardedKeypointInference:
    """forward_pass(flip=True) + get_final_preds_hrnet over a batch sharded across the ranks of the process group.

    Every rank calls ``run`` with the full batch (host or device tensors); each computes only its slice and all
    ranks return the full, ordered result.  The model must already live on this rank's GPU.
    """

    def __init__(self, model, flip=True, group=None):
        self.model, self.flip, self.group = model, flip, group

    def run(self, imgs, center, scale):
        from .inference import forward_pass
        from .pose_parsing import get_final_preds_hrnet
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = imgs.shape[0]
        lo, hi = shard_bounds(n, world, rank)
        dev = self.model.conv1.weight.device
        J = self.model.num_joints
        if hi > lo:
            x = imgs[lo:hi].to(dev, non_blocking=True)
            heat = forward_pass(self.model, x, "HRNet", device=dev, flip=self.flip)
            c = torch.as_tensor(center[lo:hi]).to(dev)
            s = torch.as_tensor(scale[lo:hi]).to(dev)
            preds, maxvals, _ = get_final_preds_hrnet(heat, c, s, as_tensor=True)
        else:
            preds = torch.zeros((0, J, 2), dtype=torch.float32, device=dev)
            maxvals = torch.zeros((0, J, 1), dtype=torch.float32, device=dev)
        return gather_keypoints(preds, maxvals, n, self.group)
