"""Generate tests/golden/*.npz by running the UNMODIFIED reference code (build container only).

Test infrastructure.  Run:  python -m oracle.make_golden
Needs /root/reference; the fixtures it writes are committed so that the GPU box (which has no
/root/reference) can still check the oracle and the CUDA path against the reference's outputs.

Inputs are regenerated from seeds by tests (``golden_inputs``) except where platform libm could
perturb them (Gaussian blobs), which are stored as float16-exact values.
"""
import os
import sys

import numpy as np
import torch

from oracle import hrnet_oracle, pose_oracle, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_inputs():
    """Seeded inputs shared by the generator and the tests."""
    rng = np.random.default_rng(1234)
    d = {}
    d["hm_rand_64x48"] = rng.standard_normal((2, 17, 64, 48)).astype(np.float32)
    d["hm_rand_96x72"] = rng.standard_normal((1, 17, 96, 72)).astype(np.float32)
    d["center_a"], d["scale_a"] = pose_oracle.synth_boxes(2, seed=7)
    d["center_b"], d["scale_b"] = pose_oracle.synth_boxes(1, seed=8)
    d["flip_out"] = rng.standard_normal((2, 17, 64, 48)).astype(np.float32)
    d["flip_out_f"] = rng.standard_normal((2, 17, 64, 48)).astype(np.float32)
    d["loss_out"] = rng.standard_normal((3, 17, 64, 48)).astype(np.float32)
    d["loss_tgt"] = rng.random((3, 17, 64, 48)).astype(np.float32)
    d["loss_tw"] = rng.choice(np.array([0, 1, 1.2, 1.5], np.float32), size=(3, 17, 1))
    d["x_w32"] = rng.standard_normal((2, 3, 256, 192)).astype(np.float32)
    d["x_w48"] = rng.standard_normal((1, 3, 384, 288)).astype(np.float32)
    return d


class _TwoShotModel:
    """Stand-in 'model' for lib.inference.forward_pass: returns fixed heatmaps for the plain and the
    flipped call, so the reference's flip_back / shift / average code runs on known inputs."""

    def __init__(self, out, out_f):
        self.outs = [torch.from_numpy(out.copy()), torch.from_numpy(out_f.copy())]
        self.calls = 0

    def __call__(self, img):
        r = self.outs[self.calls]
        self.calls += 1
        return r


def main():
    if not ref_shim.available():
        sys.exit("reference not available; goldens can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    L = ref_shim.lib()
    g = golden_inputs()

    # ---- decode ------------------------------------------------------------------------------
    blobs = pose_oracle.blob_heatmaps(2, 17, 64, 48, seed=3).astype(np.float16).astype(np.float32)
    center_c, scale_c = pose_oracle.synth_boxes(2, seed=9)
    out = {"hm_blobs_f16": blobs.astype(np.float16)}
    for tag, hm, c, s in (("rand64", g["hm_rand_64x48"], g["center_a"], g["scale_a"]),
                          ("rand96", g["hm_rand_96x72"], g["center_b"], g["scale_b"]),
                          ("blobs", blobs, center_c, scale_c)):
        p0, m0 = L.pose_parsing.get_max_preds_hrnet(hm.copy())
        preds, maxvals, coords = L.pose_parsing.get_final_preds_hrnet(hm.copy(), c, s)
        out[f"{tag}_max_preds"] = p0
        out[f"{tag}_max_vals"] = m0
        out[f"{tag}_preds"] = preds
        out[f"{tag}_maxvals"] = maxvals
        out[f"{tag}_coords"] = coords
    np.savez_compressed(os.path.join(GOLDEN_DIR, "decode.npz"), **out)

    # ---- flip-test averaging -----------------------------------------------------------------
    model = _TwoShotModel(g["flip_out"], g["flip_out_f"])
    avg = L.inference.forward_pass(model, torch.zeros(2, 3, 8, 8), "HRNet", device="cpu", flip=True)
    fb = L.transforms.flip_back(g["flip_out_f"].copy(), L.CONSTANTS.FLIP_PAIRS)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "flip.npz"), avg=avg.numpy(), flip_back=fb.numpy())

    # ---- loss --------------------------------------------------------------------------------
    o = torch.from_numpy(g["loss_out"]).requires_grad_(True)
    loss = L.loss.PersonMSELoss()(o, torch.from_numpy(g["loss_tgt"]), torch.from_numpy(g["loss_tw"]))
    loss.backward()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "loss.npz"), loss=loss.detach().numpy(),
                        grad=o.grad.numpy())

    # ---- HRNet forward -----------------------------------------------------------------------
    for width, hw, key in ((32, (256, 192), "x_w32"), (48, (384, 288), "x_w48")):
        m = ref_shim.build_reference_hrnet(width, hw).eval()
        m.load_state_dict(hrnet_oracle.synth_state_dict(width, seed=0), strict=True)
        with torch.no_grad():
            y = m(torch.from_numpy(g[key]))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"hrnet_w{width}_fwd.npz"), y=y.numpy())
        print(f"w{width}: y range [{y.min():.3f}, {y.max():.3f}] std {y.std():.3f}")
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
