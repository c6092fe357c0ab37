"""Informative same-box comparator for the training step: the reference algorithm (oracle restatement) under
torch autograd + cuDNN on the B200, bf16 autocast, channels_last, SGD -- "the reference's kernels on Blackwell"."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import hrnet_oracle

torch.backends.cudnn.benchmark = True
sd0 = hrnet_oracle.synth_state_dict(32, seed=0)
for B in [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "32,128").split(",")]:
    sd = {k: (v.cuda().contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v.cuda()) for k, v in sd0.items()}
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k]
    opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4)
    x = torch.randn(B, 3, 256, 192, device="cuda").contiguous(memory_format=torch.channels_last)
    tgt = torch.rand(B, 17, 64, 48, device="cuda"); tw = torch.ones(B, 17, 1, device="cuda")
    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            heat = hrnet_oracle.hrnet_forward_train(sd, x, 32)
        d = (heat.float() - tgt).reshape(B, 17, -1) * tw
        loss = 0.5 * (d * d).mean(dim=(0, 2)).sum() / 17
        opt.zero_grad(); loss.backward(); opt.step()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps(dict(comparator="torch autograd + cuDNN, bf16 autocast, channels_last, eager", batch=B,
                          ms_per_step=round(ms, 2), crops_per_s=round(B / ms * 1e3, 1))))
