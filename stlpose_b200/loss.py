"""Drop-in for the reference ``lib.loss.PersonMSELoss`` (/root/reference/src/lib/loss.py:61-94)."""
import torch
import torch.nn as nn

from . import _lib


class _PersonMSE(torch.autograd.Function):
    """Fused forward + gradient: one pass over output/target produces the loss and dloss/doutput."""

    @staticmethod
    def forward(ctx, output, target, target_weight):
        if not output.is_cuda:
            raise _lib.StlError("PersonMSELoss runs on CUDA only (there is no CPU fallback)")
        B, J = output.shape[0], output.shape[1]
        o = output.detach().reshape(B, J, -1).float().contiguous()
        t = target.detach().reshape(B, J, -1).to(o.device).float().contiguous()
        tw = target_weight.detach().reshape(B, J).to(o.device).float().contiguous()
        hw = o.shape[2]
        L = _lib.lib()
        loss = torch.empty((), dtype=torch.float32, device=o.device)
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(o) if need_grad else None
        ws = torch.empty(L.stl_mse_workspace_bytes(), dtype=torch.uint8, device=o.device)
        with torch.cuda.device(o.device):
            _lib.check(L.stl_mse_loss_fwd_bwd(_lib.ptr(o), _lib.ptr(t), _lib.ptr(tw), B, J, hw, _lib.ptr(loss),
                                              _lib.ptr(grad), _lib.ptr(ws), _lib.current_stream()))
        ctx.grad = grad                      # dloss/doutput for an upstream gradient of 1, consumed by backward
        ctx.out_shape = output.shape
        ctx.out_dtype = output.dtype
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        grad = ctx.grad
        if grad is None:
            raise RuntimeError("PersonMSELoss: backward called twice; the fused gradient buffer is consumed by the "
                               "first call (run the forward again)")
        ctx.grad = None
        g = g.detach().reshape(1).float().contiguous()
        with torch.cuda.device(grad.device):     # scales in place, and touches no memory when g == 1
            _lib.check(_lib.lib().stl_scale_inplace(_lib.ptr(grad), _lib.ptr(g), grad.numel(), _lib.current_stream()))
        return grad.reshape(ctx.out_shape).to(ctx.out_dtype), None, None


class PersonMSELoss(nn.Module):
    """loss = 0.5/(J*B*h*w) * sum (tw * (output - target))^2.

    As in the reference, ``use_target_weight`` is stored but ignored (loss.py:64-68,71) and ``target_weight``
    must be a [B,J,1] tensor.
    """

    def __init__(self, use_target_weight=1):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight=1):
        if not torch.is_tensor(target_weight):
            raise TypeError("target_weight must be a [B,J,1] tensor (the reference's default of 1 is unusable too)")
        return _PersonMSE.apply(output, target, target_weight)
