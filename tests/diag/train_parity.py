"""Diagnostic: device train step vs the fp32 oracle and vs the bf16-storage oracle (prints, asserts nothing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from oracle import hrnet_oracle, pose_oracle
import stlpose_b200 as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sd0 = hrnet_oracle.synth_state_dict(32, seed=0)
GAIN = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for k in sd0:
    if (k.endswith("bn2.weight") and "branches" in k) or k.endswith("bn3.weight"):
        sd0[k] = sd0[k] * GAIN
print("residual gain", GAIN)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 3, 256, 192, generator=g)
tgt = torch.from_numpy(pose_oracle.blob_heatmaps(B, 17, 64, 48, seed=1, noise=0.0))
tw = torch.tensor([0.0, 1.0, 1.2, 1.5])[torch.randint(0, 4, (B, 17, 1), generator=g)]

def oracle(bf16):
    sd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone())
          for k, v in sd0.items()}
    t = time.time()
    heat = hrnet_oracle.hrnet_forward_train(sd, x, 32, bf16_storage=bf16)
    d = (heat - tgt).reshape(B, 17, -1) * tw
    loss = 0.5 * (d * d).mean(dim=(0, 2)).sum() / 17
    loss.backward()
    print(f"oracle bf16={bf16}: {time.time()-t:.1f}s loss {loss.item():.6f}")
    return sd, heat.detach()

def cmp(name, ga, gb, ha, hb):
    e = ha - hb
    print(f"[{name}] heat max-abs {e.abs().max():.4f} rel-RMS {(e.norm()/hb.norm()):.4f}")
    st = []
    for k in ga:
        a, b = ga[k], gb[k]
        if b.norm() < 1e-7: continue
        st.append((torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item(),
                   ((a-b).norm()/b.norm()).item(), k))
    cs = sorted(c for c, _, _ in st)
    print(f"   grads n={len(st)} cos min {cs[0]:.4f} p1 {cs[len(cs)//100]:.4f} p10 {cs[len(cs)//10]:.4f} median {cs[len(cs)//2]:.4f}")
    for c, r, k in sorted(st)[:6]:
        print(f"      {k}: cos {c:.4f} rel {r:.3f}")
    # by depth: first 10 / last 10 params in order
    order = [s for s in st]
    print("   head:", " ".join(f"{c:.3f}" for c, _, _ in order[:12]))
    print("   tail:", " ".join(f"{c:.3f}" for c, _, _ in order[-12:]))

sd32, h32 = oracle(False)
sd16, h16 = oracle(True)
m = S.PoseHighResolutionNet(width=32); m.load_state_dict(sd0, strict=True); m = m.cuda().train()
heat = m(x.cuda())
loss = S.PersonMSELoss()(heat, tgt.cuda(), tw.cuda()); loss.backward()
print("device loss", loss.item())
gd = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()}
g32 = {k: sd32[k].grad for k in gd}; g16 = {k: sd16[k].grad for k in gd}
hd = heat.detach().cpu()
cmp("device vs fp32 oracle", gd, g32, hd, h32)
cmp("device vs bf16 oracle", gd, g16, hd, h16)
cmp("bf16 oracle vs fp32 oracle", g16, g32, h16, h32)
new = m.state_dict()
for key in ("bn1.running_mean", "bn1.running_var", "stage3.2.branches.1.3.bn2.running_var", "stage4.2.fuse_layers.0.3.1.running_mean"):
    print(key, (new[key].cpu()-sd16[key]).abs().max().item(), (new[key].cpu()-sd32[key]).abs().max().item())
print("per-BN batch-variance relative deviation (device vs bf16 oracle | bf16 oracle vs fp32 oracle)")
keys = [k for k in sd0 if k.endswith("running_var")]
for i, k in enumerate(keys):
    vd = (new[k].cpu() - 0.9) / 0.1; v16 = (sd16[k] - 0.9) / 0.1; v32 = (sd32[k] - 0.9) / 0.1
    a = ((vd - v16).norm() / v16.norm()).item(); b = ((v16 - v32).norm() / v32.norm()).item()
    if i < 12 or i % 32 == 0:
        print(f"  {i:3d} {k:55s} {a:.2e} | {b:.2e}")
