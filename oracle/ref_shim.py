"""Import the UNMODIFIED reference hot-path modules from /root/reference (build container only).

Test infrastructure.  The reference cannot be imported as shipped:
``models/__init__.py:8`` imports a missing module, ``yacs`` is not installed, and the HRNet
architecture YAML lives outside the repo at a cwd-relative, file-name-fixed path
(``models/HRnet.py:280-283``, ``CONFIG.py:14``).  This shim works around exactly those three
things and nothing else; every FLOP still runs through the reference's own code.

Used only by ``oracle/make_golden.py``, by tests that are skipped when the reference is absent, and by
``bench.py``'s reference arm / ``cpu_baseline`` leg.  On the GPU box /root/reference does not exist; the shim
then imports the byte-for-byte copies that ``oracle/stage_ref.py`` placed under ``oracle/_ref/src``.
"""
import copy
import importlib
import os
import sys
import tempfile
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")   # oracle/stage_ref.py


def _default_src():
    """/root/reference/src in the build container; on the GPU box the byte-for-byte copies staged under oracle/_ref."""
    env = os.environ.get("STLPOSE_REFERENCE_SRC")
    if env:
        return env
    if os.path.isfile("/root/reference/src/models/HRnet.py"):
        return "/root/reference/src"
    return _STAGED


REF_SRC = _default_src()


def available():
    return os.path.isfile(os.path.join(REF_SRC, "models", "HRnet.py"))


class _CfgNode(dict):
    """Minimal stand-in for yacs.config.CfgNode (dict + attribute access).

    Attribute misses must raise AttributeError (not KeyError): the reference stores cfg nodes
    as module attributes (HRnet.py:299,309,320) and copy.deepcopy probes dunder attributes.
    """

    def __init__(self, init=None, new_allowed=False):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = _CfgNode(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return _CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def defrost(self):
        pass

    def freeze(self):
        pass

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            data = yaml.safe_load(f)
        self._merge(data)

    def _merge(self, data):
        for k, v in data.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], _CfgNode):
                    self[k] = _CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v


def _yaml_text(width, image_hw):
    h, w = image_hw
    c = [width, 2 * width, 4 * width, 8 * width]
    return f"""MODEL:
  NAME: pose_hrnet
  NUM_JOINTS: 17
  IMAGE_SIZE: [{w}, {h}]
  HEATMAP_SIZE: [{w // 4}, {h // 4}]
  SIGMA: 2
  EXTRA:
    PRETRAINED_LAYERS: ['*']
    FINAL_CONV_KERNEL: 1
    STAGE2: {{NUM_MODULES: 1, NUM_BRANCHES: 2, BLOCK: BASIC, NUM_BLOCKS: [4, 4], NUM_CHANNELS: [{c[0]}, {c[1]}], FUSE_METHOD: SUM}}
    STAGE3: {{NUM_MODULES: 4, NUM_BRANCHES: 3, BLOCK: BASIC, NUM_BLOCKS: [4, 4, 4], NUM_CHANNELS: [{c[0]}, {c[1]}, {c[2]}], FUSE_METHOD: SUM}}
    STAGE4: {{NUM_MODULES: 3, NUM_BRANCHES: 4, BLOCK: BASIC, NUM_BLOCKS: [4, 4, 4, 4], NUM_CHANNELS: [{c[0]}, {c[1]}, {c[2]}, {c[3]}], FUSE_METHOD: SUM}}
"""


_installed = False


def _install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF_SRC}")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    yacs = types.ModuleType("yacs")
    yacs_config = types.ModuleType("yacs.config")
    yacs_config.CfgNode = _CfgNode
    yacs.config = yacs_config
    sys.modules.setdefault("yacs", yacs)
    sys.modules.setdefault("yacs.config", yacs_config)
    # bypass the broken models/__init__.py (imports a module that is not in the repo)
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF_SRC, "models")]
    sys.modules["models"] = pkg
    _installed = True


def build_reference_hrnet(width=32, image_hw=(256, 192)):
    """Construct the reference PoseHighResolutionNet (HRnet.py:275-339) for the given width."""
    _install()
    tmp = tempfile.mkdtemp(prefix="stlpose_ref_")
    os.makedirs(os.path.join(tmp, "resources", "HRnet"))
    os.makedirs(os.path.join(tmp, "src"))
    with open(os.path.join(tmp, "resources", "HRnet", "cfg_hrnet_w32_256x192.yaml"), "w") as f:
        f.write(_yaml_text(width, image_hw))
    cwd = os.getcwd()
    os.chdir(os.path.join(tmp, "src"))
    try:
        hr = importlib.import_module("models.HRnet")
        model = hr.PoseHighResolutionNet(is_train=False)
    finally:
        os.chdir(cwd)
    return model


def lib():
    """Return the reference's lib.inference / lib.pose_parsing / lib.transforms / lib.loss."""
    _install()
    ns = types.SimpleNamespace()
    ns.inference = importlib.import_module("lib.inference")
    ns.pose_parsing = importlib.import_module("lib.pose_parsing")
    ns.transforms = importlib.import_module("lib.transforms")
    ns.loss = importlib.import_module("lib.loss")
    ns.CONSTANTS = importlib.import_module("CONSTANTS")
    return ns


def metrics_functions():
    """``calc_dists`` and ``dist_acc`` of the reference's lib/metrics.py, executed from their unmodified source text.

    The module itself cannot be imported (it needs pycocotools, and lines 355-356 inside ``accuracy`` are corrupted and
    do not parse), so the two self-contained helper functions are cut out by their ``def`` blocks and exec'd with NumPy
    in scope; ``accuracy`` is restated in oracle/pose_oracle.py around them."""
    import numpy as np
    path = os.path.join(REF_SRC, "lib", "metrics.py")
    lines = open(path).read().split("\n")
    ns = {"np": np}
    for name in ("calc_dists", "dist_acc"):
        start = next(i for i, l in enumerate(lines) if l.startswith(f"def {name}("))
        end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("def "))
        exec(compile("\n".join(lines[start:end]), f"{path}:{start + 1}", "exec"), ns)
    return types.SimpleNamespace(calc_dists=ns["calc_dists"], dist_acc=ns["dist_acc"])


def generate_target_function():
    """``JointsDataset.generate_target`` (data/JointsDataset.py:230-286) executed from its unmodified source text; the
    module itself does not import here (pycocotools, removed NumPy aliases).  Returns f(self_like, joints, joints_vis)
    where ``self_like`` carries num_joints, target_type, heatmap_size, image_size, sigma, use_different_joints_weight,
    joints_weight."""
    import textwrap
    import numpy as np
    path = os.path.join(REF_SRC, "data", "JointsDataset.py")
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("def generate_target("))
    indent = len(lines[start]) - len(lines[start].lstrip())
    end = start + 1
    while end < len(lines) and (not lines[end].strip() or len(lines[end]) - len(lines[end].lstrip()) > indent):
        end += 1
    ns = {"np": np}
    exec(compile(textwrap.dedent("\n".join(lines[start:end])), f"{path}:{start + 1}", "exec"), ns)
    return ns["generate_target"]


def submission_functions():
    """``generate_submission_hrnet`` (lib/metrics.py:192-265) and ``convert_keypoints_to_coco_format``
    (data/data_processing.py:52-82) executed from their unmodified source text, around the reference's importable
    ``lib.nms``.  Neither module imports here (pycocotools; ``REORDER_MAP``); the only stand-in is ``np.float`` (removed
    from NumPy >= 1.24, used at data_processing.py:62), supplied as an attribute of the namespace's ``np``."""
    import json
    from collections import defaultdict
    import numpy as np
    _install()
    nms_lib = importlib.import_module("lib.nms")

    def cut(path, name):
        lines = open(path).read().split("\n")
        start = next(i for i, l in enumerate(lines) if l.startswith(f"def {name}("))
        end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("def ")), len(lines))
        return compile("\n" * start + "\n".join(lines[start:end]), path, "exec")

    np_ns = types.SimpleNamespace(**{k: getattr(np, k) for k in ("array", "zeros", "concatenate")}, float=np.float64)
    dp_ns = {"np": np_ns}
    exec(cut(os.path.join(REF_SRC, "data", "data_processing.py"), "convert_keypoints_to_coco_format"), dp_ns)
    m_ns = {"np": np, "json": json, "defaultdict": defaultdict, "nms_lib": nms_lib,
            "data_processing": types.SimpleNamespace(convert_keypoints_to_coco_format=dp_ns["convert_keypoints_to_coco_format"])}
    exec(cut(os.path.join(REF_SRC, "lib", "metrics.py"), "generate_submission_hrnet"), m_ns)
    return types.SimpleNamespace(generate_submission_hrnet=m_ns["generate_submission_hrnet"], nms=nms_lib,
                                 convert_keypoints_to_coco_format=dp_ns["convert_keypoints_to_coco_format"])
