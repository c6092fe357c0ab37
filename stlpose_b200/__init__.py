"""stlpose_b200 -- B200-native HRNet keypoint hot path behind the STLPose Python interface.

    from stlpose_b200 import PoseHighResolutionNet, forward_pass, get_final_preds_hrnet, PersonMSELoss

mirrors ``models.PoseHighResolutionNet``, ``lib.inference.forward_pass``, ``lib.pose_parsing.get_*_preds_hrnet``,
``lib.transforms.flip_back`` and ``lib.loss.PersonMSELoss`` of angelvillar96/STLPose.  All arithmetic runs in
``libstlpose_b200.so`` (hand-written CUDA for sm_100a); importing this package without the built library, or
calling it without a CUDA device, raises -- there is no CPU path.
"""
from ._lib import StlError, lib  # noqa: F401
from .hrnet import PoseHighResolutionNet  # noqa: F401
from .inference import forward_pass  # noqa: F401
from .loss import PersonMSELoss  # noqa: F401
from .pose_parsing import get_final_preds_hrnet, get_max_preds_hrnet  # noqa: F401
from .train_step import TrainStep  # noqa: F401
from .transforms import FLIP_PAIRS, flip_back  # noqa: F401
