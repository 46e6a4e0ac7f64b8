"""rpe_b200 -- B200-native (sm_100a) per-frame frame-to-frame pose path of
aimi-lab/robust-pose-estimator, drop-in behind the reference's PoseNet / PoseEstimator /
infer_trajectory API.  Host code is Python/PyTorch (device memory, streams, torch.distributed);
the hot stages are hand-written CUDA reached through the C ABI declared in include/rpe_b200.h.

Module paths mirror the reference (core/pose/pose_net.py -> rpe_b200.core.pose.pose_net, ...).
There is NO CPU fallback: every operator raises if the CUDA library is missing or a tensor is not
on a CUDA device.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not load the shared library until first use)
