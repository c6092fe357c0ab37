import sys, os, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import stlpose_b200 as S
from stlpose_b200 import _lib
mode = sys.argv[1]
m = S.PoseHighResolutionNet(width=32).cuda().train()
opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9)
crit = S.PersonMSELoss()
def data(B):
    return torch.randn(B, 3, 256, 192, device="cuda"), torch.rand(B, 17, 64, 48, device="cuda"), torch.ones(B, 17, 1, device="cuda")
a = S.TrainStep(m, opt, crit, batch=8)
print("first ok", a(*data(8)).item())
if mode == "eager_between":
    x, t, w = data(8)
    loss = crit(S.forward_pass(m, x, "HRNet", device="cuda"), t, w); opt.zero_grad(); loss.backward(); opt.step()
    print("eager ok", loss.item())
if mode == "delete_first":
    del a; gc.collect(); torch.cuda.empty_cache()
try:
    b = S.TrainStep(m, opt, crit, batch=4)
    print("second ok", b(*data(4)).item())
except Exception as e:
    print("second FAILED:", str(e)[:300]); print("lib error:", _lib.lib().stl_last_error().decode())
