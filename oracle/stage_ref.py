"""Stage the UNMODIFIED reference hot-path modules into oracle/_ref/src so that they travel to the GPU box.

Test / measurement infrastructure (never imported by stlpose_b200/).  /root/reference does not exist on the GPU box, and
nothing of the reference may be committed; ``oracle/_ref/`` is git-ignored but NOT gpurun-ignored, so files copied there
by this recipe ship with the snapshot like a built .so does.  ``__graft_entry__.build()`` runs this whenever
/root/reference is mounted; ``bench.py --impl reference`` and the ``cpu_baseline`` leg then drive the reference's own

    lib.inference.forward_pass(model, img, "HRNet", flip=True)        (lib/inference.py:11-27)
    lib.pose_parsing.get_final_preds_hrnet(heatmaps, center, scale)   (lib/pose_parsing.py:58-92)

over models.HRnet.PoseHighResolutionNet (models/HRnet.py:275-468) through oracle/ref_shim.py (``kind: "reference"``).
Files are copied byte for byte (sha256 listed in oracle/_ref/MANIFEST.json); the shim works around the three import
problems described in ref_shim's header and nothing else.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "src")
# hot-path modules + the modules they import at module level
FILES = [
    "CONSTANTS.py", "CONFIG.py",
    "models/HRnet.py", "models/utils/hrnet_config.py",
    "lib/inference.py", "lib/pose_parsing.py", "lib/transforms.py", "lib/loss.py", "lib/logger.py", "lib/nms.py",
]


def stage(src="/root/reference/src", dest=DEST, quiet=False):
    if not os.path.isfile(os.path.join(src, "models", "HRnet.py")):
        raise FileNotFoundError(f"reference sources not found under {src}")
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(os.path.dirname(dest), "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1)
    if not quiet:
        print(f"staged {len(FILES)} reference files into {dest}")
    return manifest


if __name__ == "__main__":
    stage(*(sys.argv[1:2] or ["/root/reference/src"]))
