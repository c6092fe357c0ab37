"""One eager fine-tuning step with the stream-parallel schedule (module branches and weight gradients on side streams),
then a graph-captured TrainStep replay - the smallest program that exercises every training kernel and every
cross-stream hand-off.  Run it under ``compute-sanitizer --tool {memcheck,racecheck,synccheck}`` (one tool per gpurun
call); the logs live in profiles/r02_sanitizer_*.txt."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import stlpose_b200 as S
from stlpose_b200 import training
from oracle import hrnet_oracle, pose_oracle


def main():
    B = int(os.environ.get("SAN_BATCH", "2"))
    graph = os.environ.get("SAN_GRAPH", "1") != "0"
    assert training.BRANCH_STREAMS and training.SIDE_WGRAD
    sd = hrnet_oracle.synth_state_dict(32, seed=0)
    m = S.PoseHighResolutionNet(width=32)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, 256, 192, generator=g).cuda()
    tgt = torch.from_numpy(pose_oracle.blob_heatmaps(B, 17, 64, 48, seed=1, noise=0.0)).cuda().float()
    tw = torch.ones(B, 17, 1).cuda()
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)
    crit = S.PersonMSELoss()
    if graph:
        step = S.TrainStep(m, opt, crit, batch=B, warmup=1)
        loss = step(x, tgt, tw)
        torch.cuda.synchronize()
        print("graph step loss", float(loss))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        loss = crit(S.forward_pass(m, x, "HRNet", device="cuda", flip=False), tgt, tw)
        opt.zero_grad()
        loss.backward()
        opt.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    print("eager step loss", float(loss))
    # inference path too (eval-mode forward + flip + decode)
    m.eval()
    m.invalidate_packed_weights()
    c, s = pose_oracle.synth_boxes(B, seed=0)
    heat = S.forward_pass(m, x, "HRNet", device="cuda", flip=True)
    S.get_final_preds_hrnet(heat, c, s)
    torch.cuda.synchronize()
    print("SANITIZE_PROGRAM_OK")


if __name__ == "__main__":
    main()
