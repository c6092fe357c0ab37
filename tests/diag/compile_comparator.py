"""Same-box comparator for the headline workload: the UNMODIFIED reference nn.Module (staged under oracle/_ref) under the
best stock PyTorch recipe for inference on this GPU - eval mode, bf16, channels_last, torch.compile(mode="max-autotune")
with inductor freezing (BatchNorm folded into the convolutions at compile time) and CUDA graphs - running the same
step as bench.py: 512 crops, plain + mirrored forward.  Diagnostic only (cuDNN / Triton are not used by the product).

    python tests/diag/compile_comparator.py [--batch 512] [--mode max-autotune|reduce-overhead|eager]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import hrnet_oracle, ref_shim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--mode", default="max-autotune")
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    m = ref_shim.build_reference_hrnet(32, (256, 192))
    m.load_state_dict(hrnet_oracle.synth_state_dict(32, seed=0), strict=True)
    m = m.cuda().eval().to(memory_format=torch.channels_last).bfloat16()
    x = torch.randn(a.batch, 3, 256, 192, device="cuda").to(memory_format=torch.channels_last).bfloat16()
    fn = m
    t0 = time.time()
    if a.mode != "eager":
        import torch._inductor.config as icfg
        icfg.freezing = True                      # constant-fold BatchNorm into the conv weights
        fn = torch.compile(m, mode=a.mode)
    with torch.no_grad():
        for _ in range(3):
            fn(x); fn(x.flip(3))
        torch.cuda.synchronize()
        compile_s = time.time() - t0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn(x); fn(x.flip(3))                 # the flip test: two forwards per crop (lib/inference.py:18-22)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(json.dumps({"comparator": f"reference nn.Module, torch {torch.__version__} compile mode={a.mode}, freezing, bf16 "
                                    "channels_last, 2 forwards per crop (no flip-average / decode)",
                      "batch": a.batch, "ms_per_step": ms, "crops_per_s": a.batch / ms * 1e3,
                      "tflops": 2 * 15.29e9 * a.batch / ms / 1e9, "compile_s": round(compile_s, 1)}))


if __name__ == "__main__":
    main()
