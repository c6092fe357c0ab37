import faulthandler, os, sys
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import stlpose_b200 as S
from stlpose_b200.parallel import GradientReducer
B = 2
torch.manual_seed(0)
m = S.PoseHighResolutionNet(width=32).cuda().train()
opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
red = GradientReducer(m.parameters(), local_batch=B, bucket_bytes=4 << 20)
print("buckets", len(red.buckets), flush=True)
red.bind(m)
print("bound", flush=True)
x = torch.randn(B, 3, 256, 192).cuda(); t = torch.rand(B, 17, 64, 48).cuda(); w = torch.ones(B, 17, 1).cuda()
crit = S.PersonMSELoss()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    out = S.forward_pass(m, x, "HRNet", device="cuda", flip=False)
    print("fwd", flush=True)
    loss = crit(out, t, w)
    print("loss", float(loss), flush=True)
    loss.backward(gradient=red.scale_tensor)
    print("bwd", flush=True)
    red.finish_backward()
    opt.step()
torch.cuda.synchronize()
print("eager ok", flush=True)
red.unbind()
step = S.TrainStep(m, opt, crit, batch=B, reducer=red)
print("captured", step.optimizer_in_graph, flush=True)
print(float(step(x, t, w)), flush=True)
