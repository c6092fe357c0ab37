"""Kernel-level breakdown of one fine-tuning step (torch.profiler, CUDA activity only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import stlpose_b200 as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = S.PoseHighResolutionNet(width=32).cuda().train()
opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)
crit = S.PersonMSELoss()
x = torch.randn(B, 3, 256, 192, device="cuda"); tgt = torch.rand(B, 17, 64, 48, device="cuda"); tw = torch.ones(B, 17, 1, device="cuda")
def step():
    loss = crit(m(x), tgt, tw); opt.zero_grad(); loss.backward(); opt.step()
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
