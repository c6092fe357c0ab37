"""One eager fine-tuning step bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off` launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stlpose_b200 as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = S.PoseHighResolutionNet(width=32).cuda().train()
opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)
crit = S.PersonMSELoss()
x = torch.randn(B, 3, 256, 192, device="cuda"); tgt = torch.rand(B, 17, 64, 48, device="cuda"); tw = torch.ones(B, 17, 1, device="cuda")
def step():
    loss = crit(S.forward_pass(m, x, "HRNet", device="cuda"), tgt, tw)
    opt.zero_grad(); loss.backward(); opt.step()
    return loss
step(); torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
