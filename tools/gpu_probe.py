"""First-contact probe for the tensor-core conv kernel on a real B200.

Runs each case in its own subprocess (a device trap kills the CUDA context) with a timeout, and prints one line
per case: impl 0 = shifted-descriptor taps, 1 = TMA load per tap, 2 = CUDA-core reference kernel.
Usage: python tools/gpu_probe.py [--quick]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # name, kwargs
    ("1x1 c64->64 64x48 N2", dict(N=2, cin=64, cout=64, H=64, W=48, k=1, stride=1)),
    ("1x1 c32->32 16x12 N3", dict(N=3, cin=32, cout=32, H=16, W=12, k=1, stride=1)),
    ("3x3 c32->32 64x48 N2", dict(N=2, cin=32, cout=32, H=64, W=48, k=3, stride=1)),
    ("3x3 c64->64 32x24 N3 res", dict(N=3, cin=64, cout=64, H=32, W=24, k=3, stride=1, with_res=True)),
    ("3x3 c128->128 16x12 N5", dict(N=5, cin=128, cout=128, H=16, W=12, k=3, stride=1)),
    ("3x3 c256->256 8x6 N9", dict(N=9, cin=256, cout=256, H=8, W=6, k=3, stride=1)),
    ("3x3 c256->32 64x48 N1", dict(N=1, cin=256, cout=32, H=64, W=48, k=3, stride=1)),
    ("3x3s2 c64->64 128x96 N2", dict(N=2, cin=64, cout=64, H=128, W=96, k=3, stride=2)),
    ("3x3s2 c32->64 64x48 N3 res up", dict(N=3, cin=32, cout=64, H=64, W=48, k=3, stride=2, with_res=True, n_up=2)),
    ("3x3s2 c128->256 16x12 N5", dict(N=5, cin=128, cout=256, H=16, W=12, k=3, stride=2)),
    ("1x1 c64->256 64x48 N2 res", dict(N=2, cin=64, cout=256, H=64, W=48, k=1, stride=1, with_res=True)),
    ("head 1x1 c32->17 nchw", dict(N=3, cin=32, cout=17, H=64, W=48, k=1, stride=1, relu=False, out_nchw=True,
                                   with_bias=True)),
    ("3x3 c48->48 24x18 N4 (W48)", dict(N=4, cin=48, cout=48, H=24, W=18, k=3, stride=1)),
    ("3x3 c96->96 48x36 N2 (W48)", dict(N=2, cin=96, cout=96, H=48, W=36, k=3, stride=1)),
    ("3x3 c192->192 24x18 N3 (W48)", dict(N=3, cin=192, cout=192, H=24, W=18, k=3, stride=1)),
]

CHILD = r"""
import json, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from gpu_util import conv_case
kw = json.loads({kw!r})
err, scale = conv_case(**kw)
print("RESULT", err, scale)
"""


def run_case(kw, timeout=120):
    code = CHILD.format(root=ROOT, kw=json.dumps(kw))
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return "TIMEOUT"
    for line in r.stdout.splitlines():
        if line.startswith("RESULT"):
            _, err, scale = line.split()
            err, scale = float(err), float(scale)
            ok = err <= 1e-2 * max(scale, 1.0)
            return f"{'ok  ' if ok else 'BAD '} err {err:.4g} (ref max {scale:.3g})"
    tail = (r.stdout + r.stderr).strip().splitlines()[-6:]
    return "FAIL rc=%d :: %s" % (r.returncode, " | ".join(tail))


def main():
    quick = "--quick" in sys.argv
    cases = CASES[:4] if quick else CASES
    for name, kw in cases:
        for impl in (2, 1, 0):
            if kw["stride"] == 2 and impl == 0:
                continue  # stride-2 has a single tensor-core path (impl 1 == impl 0)
            res = run_case(dict(kw, impl=impl))
            print(f"[impl {impl}] {name:34s} {res}", flush=True)


if __name__ == "__main__":
    main()
