"""Informative same-box comparator: the reference algorithm (oracle restatement, identical graph)
executed by the container's torch + cuDNN on the B200 in bf16 channels_last, flip-test + decode not included.
This is "the reference's kernels on Blackwell"; it is not part of any product path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import hrnet_oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
width = int(sys.argv[2]) if len(sys.argv) > 2 else 32
hw = (256, 192) if width == 32 else (384, 288)
torch.backends.cudnn.benchmark = True
sd = hrnet_oracle.synth_state_dict(width, seed=0)
for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "fp32(tf32 off)")):
    sdc = {k: (v.cuda().to(dt) if v.is_floating_point() else v.cuda()) for k, v in sd.items()}
    sdc = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sdc.items()}
    bs = B if dt == torch.bfloat16 else B // 4
    x = torch.randn(bs, 3, *hw, device="cuda").to(dt).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        y = hrnet_oracle.hrnet_forward(sdc, x, width)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = hrnet_oracle.hrnet_forward(sdc, x, width)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = hrnet_oracle.conv_flops_per_crop(width, hw)
    print(json.dumps(dict(comparator="torch+cuDNN eager, unfused BN, channels_last", dtype=name, width=width, batch=bs,
                          ms_per_forward=round(ms, 2), forwards_per_s=round(bs / ms * 1e3, 1),
                          crops_per_s_with_flip=round(bs / ms * 1e3 / 2, 1), tflops=round(fl * bs / ms / 1e9, 1))))
