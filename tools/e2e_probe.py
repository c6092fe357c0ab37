"""Where the end-to-end call loses time against the device-resident step: variants of KeypointPipeline.__call__."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stlpose_b200 as S
from stlpose_b200.pipeline import KeypointPipeline

B = 512
torch.manual_seed(0)
model = S.PoseHighResolutionNet(width=32).cuda().eval()
pipe = KeypointPipeline(model, B, (256, 192), flip=True)
x = torch.randn(B, 3, 256, 192).pin_memory()
c = (torch.rand(B, 2) * 300 + 100).pin_memory(); s = (torch.rand(B, 2) + 0.5).pin_memory()
p = torch.empty(B, 17, 2).pin_memory(); m = torch.empty(B, 17, 1).pin_memory()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def full():
    pipe(x, c, s, p, m)

def no_small():            # crops only: no box / result copies
    cur = torch.cuda.current_stream()
    j = pipe._calls & 1; pipe._calls += 1
    with torch.cuda.stream(pipe._copy_stream):
        pipe._copy_stream.wait_event(pipe._stage_free[j])
        pipe._stage[j].copy_(x, non_blocking=True)
        pipe._stage_ready[j].record(pipe._copy_stream)
    cur.wait_event(pipe._stage_ready[j])
    pipe.x.copy_(pipe._stage[j], non_blocking=True)
    pipe._stage_free[j].record(cur)
    pipe.step()

def no_h2d():              # everything but the 302 MB host copy
    pipe.x.copy_(pipe._stage[0], non_blocking=True)
    pipe.center.copy_(c, non_blocking=True); pipe.scale.copy_(s, non_blocking=True)
    pipe.step()
    p.copy_(pipe.preds, non_blocking=True); m.copy_(pipe.maxvals, non_blocking=True)

def h2d_only_overlap():    # host copy on the side stream, never consumed: does it slow the step down?
    with torch.cuda.stream(pipe._copy_stream):
        pipe._stage[0].copy_(x, non_blocking=True)
    pipe.step()

print("resident step        %.2f ms" % timed(pipe.step))
print("full e2e call        %.2f ms" % timed(full))
print("crops only           %.2f ms" % timed(no_small))
print("no 302 MB host copy  %.2f ms" % timed(no_h2d))
print("step + unrelated H2D %.2f ms" % timed(h2d_only_overlap))
print("resident step again  %.2f ms" % timed(pipe.step))
print("full e2e call again  %.2f ms" % timed(full))
print("resident, 60 steps   %.2f ms" % timed(pipe.step, 60))
print("full e2e, 60 steps   %.2f ms" % timed(full, 60))
