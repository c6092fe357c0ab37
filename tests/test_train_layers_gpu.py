"""Layer-by-layer parity of the training path on the UNDAMPED synthetic checkpoint.

Whole-network train-mode comparisons are limited by conditioning: batch-statistics BatchNorm over a few crops amplifies
bf16 storage rounding with depth (tests/test_train_gpu.py).  Here every one of the 292 conv + BatchNorm [+ residual]
[+ ReLU] units is run on the device IN ISOLATION on the inputs the bf16-storage oracle saw at that point of a real
forward / backward pass (activation, residual and incoming gradient, all bf16-representable), and compared with the
same unit of the oracle (== the reference module, test_oracle_vs_reference.py) evaluated on the same inputs.  Nothing
is amplified, so the bounds are tight and a wrong kernel at any depth, shape or stride shows up at its own layer.
"""
import numpy as np
import pytest
import torch

from oracle import hrnet_oracle, pose_oracle
from gpu_util import from_padded, to_padded

pytestmark = pytest.mark.gpu

# bounds relative to the largest magnitude of the compared tensor (operands are identical bf16 values on both sides;
# differences come from fp32 summation order and from 1-ulp bf16 rounding flips of stored tensors: 2^-8 = 3.9e-3)
TOL_Y = 1.2e-2          # unit output (bf16, after BatchNorm: one rounding flip of z moves y by up to gamma*rstd ulps)
TOL_DX = 2.5e-2         # input gradient (bf16; built from bf16-rounded dz)
TOL_PARAM = 1.5e-2      # dW, dgamma, dbeta (fp32 sums over all pixels of bf16-rounded factors)
TOL_STAT = 5e-3         # updated running mean / var, relative to the largest entry (the lowest-resolution branches see
                        # 36-48 samples per channel here: one 1-ulp flip of a stored conv output moves a variance by ~1e-3)


def _padded(t, c_pad=None):
    n, c, h, w = t.shape
    c_pad = c_pad or c
    return to_padded(t, c_pad).view(torch.bfloat16).view(n, h + 1, w + 1, c_pad)


def _unpadded(p, c=None):
    n, hp, wp, cp = p.shape
    return from_padded(p.contiguous().view(torch.uint8).view(-1), n, c or cp, hp - 1, wp - 1, cp)


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


@pytest.mark.parametrize("width,hw", [(32, (128, 96)), (48, (64, 64))])
def test_every_unit_matches_the_oracle_unit(width, hw):
    from stlpose_b200 import training
    B = 3
    H, W = hw
    sd0 = hrnet_oracle.synth_state_dict(width, seed=0)                  # undamped: unit-gain residual branches
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, H, W, generator=g)
    tgt = torch.from_numpy(pose_oracle.blob_heatmaps(B, 17, H // 4, W // 4, seed=2, noise=0.0))
    tw = torch.ones(B, 17, 1)
    sd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone())
          for k, v in sd0.items()}
    trace = []
    heat = hrnet_oracle.hrnet_forward_train(sd, x, width, bf16_storage=True, trace=trace)
    d = (heat - tgt).reshape(B, 17, -1) * tw
    (0.5 * (d * d).mean(dim=(0, 2)).sum() / 17).backward()
    assert len(trace) == 292
    worst = {}
    for rec in trace:
        conv, bn, stride, relu = rec["conv"], rec["bn"], rec["stride"], rec["relu"]
        xin, res = rec["x"], rec["res"]
        dy = rec["y"].grad
        if dy is None:
            continue
        dy = dy.bfloat16().float()                                     # the device stores activation gradients in bf16
        # ---- oracle unit in isolation
        usd = {k: sd0[k].clone() for k in (conv + ".weight", bn + ".weight", bn + ".bias", bn + ".running_mean",
                                           bn + ".running_var")}
        for k in (conv + ".weight", bn + ".weight", bn + ".bias"):
            usd[k].requires_grad_(True)
        xo = xin.clone().requires_grad_(True)
        ro = res.clone().requires_grad_(True) if res is not None else None
        yo = hrnet_oracle.conv_bn_unit(usd, xo, conv, bn, stride, relu, ro)
        yo.backward(dy)
        # ---- device unit on the same operands
        w = sd0[conv + ".weight"].cuda()
        cin = w.shape[1]
        cin_pad = (cin + 15) // 16 * 16
        xp = _padded(xin.cuda(), cin_pad).requires_grad_(True)
        rp = _padded(res.cuda()).requires_grad_(True) if res is not None else None
        wd = w.clone().requires_grad_(True)
        gd, bd = sd0[bn + ".weight"].cuda().requires_grad_(True), sd0[bn + ".bias"].cuda().requires_grad_(True)
        rm, rv = sd0[bn + ".running_mean"].cuda().clone(), sd0[bn + ".running_var"].cuda().clone()
        tickets = torch.zeros(8, dtype=torch.int32, device="cuda")
        yd = training._ConvBN.apply(xp, wd, gd, bd, rp, rm, rv, stride, relu, 0.1, tickets, None)
        yd.backward(_padded(dy.cuda()))
        torch.cuda.synchronize()
        errs = {
            "y": (_rel(_unpadded(yd.detach()).cpu(), yo.detach()), TOL_Y),
            "dW": (_rel(wd.grad.cpu(), usd[conv + ".weight"].grad), TOL_PARAM),
            "dgamma": (_rel(gd.grad.cpu(), usd[bn + ".weight"].grad), TOL_PARAM),
            "dbeta": (_rel(bd.grad.cpu(), usd[bn + ".bias"].grad), TOL_PARAM),
            "run_mean": (_rel(rm.cpu(), usd[bn + ".running_mean"]), TOL_STAT),
            "run_var": (_rel(rv.cpu(), usd[bn + ".running_var"]), TOL_STAT),
        }
        if cin == cin_pad:                                              # (the 3-channel stem input needs no gradient)
            errs["dx"] = (_rel(_unpadded(xp.grad).cpu(), xo.grad), TOL_DX)
        if res is not None:
            errs["dres"] = (_rel(_unpadded(rp.grad).cpu(), ro.grad), TOL_DX)
        bad = {k: v for k, v in errs.items() if not (v[0] <= v[1])}
        assert not bad, f"{conv} (stride {stride}, relu {relu}, residual {res is not None}): {bad}"
        for k, (e, _) in errs.items():
            worst[k] = max(worst.get(k, 0.0), e)
        assert (yd[:, -1] == 0).all() and (yd[:, :, -1] == 0).all()   # zero cells of the padded layout stay zero
    print("worst relative errors over 292 units:", {k: round(v, 5) for k, v in worst.items()})
