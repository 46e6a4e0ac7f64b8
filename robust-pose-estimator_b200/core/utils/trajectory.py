"""Trajectory wire format of the reference (/root/reference/core/utils/trajectory.py:17-23, 38-61):
``trajectory.freiburg`` lines "ts tx ty tz qx qy qz qw" with translations converted mm -> m."""
import os

import numpy as np
import torch

from ...lie import SE3


def save_trajectory(trajectory, path):
    with open(os.path.join(path, "trajectory.freiburg"), "w") as f:
        for tr in trajectory:
            assert isinstance(tr["camera-pose"], SE3)
            vec = tr["camera-pose"].vec().cpu().squeeze().numpy()
            t = (vec[0] / 1000.0, vec[1] / 1000.0, vec[2] / 1000.0)
            f.write(f"{tr['timestamp']} {t[0]} {t[1]} {t[2]} {vec[3]} {vec[4]} {vec[5]} {vec[6]}\n")


def save_trajectory_array(poses, timestamps, path):
    """``save_trajectory`` for the (n,7) float32 array ``PoseEstimator.infer_sequence`` returns (mm): the same bytes as the
    reference writer produces for the equivalent list of {'camera-pose', 'timestamp'} entries."""
    poses = np.asarray(poses.cpu() if isinstance(poses, torch.Tensor) else poses, dtype=np.float32)
    with open(os.path.join(path, "trajectory.freiburg"), "w") as f:
        for ts, vec in zip(timestamps, poses):
            t = (vec[0] / 1000.0, vec[1] / 1000.0, vec[2] / 1000.0)
            f.write(f"{ts} {t[0]} {t[1]} {t[2]} {vec[3]} {vec[4]} {vec[5]} {vec[6]}\n")


def read_freiburg(path, ret_stamps=False, no_stamp=False):
    with open(path, "r") as f:
        lines = f.read().replace(",", " ").replace("\t", " ").split("\n")
    rows = [[v.strip() for v in ln.split(" ") if v.strip() != ""] for ln in lines if len(ln) > 0 and ln[0] != "#"]
    rows = [r for r in rows if len(r) > 0]
    off = 0 if no_stamp else 1
    trans = torch.from_numpy(np.asarray([r[off:off + 3] for r in rows], dtype=float)) * 1000.0      # m -> mm
    quat = torch.from_numpy(np.asarray([r[off + 3:off + 7] for r in rows], dtype=float))
    poses = SE3.InitFromVec(torch.cat((trans, quat), dim=-1))
    if ret_stamps and not no_stamp:
        stamps = [r[0] for r in rows]
        try:
            ts = np.asarray([int(s.split(".")[0] + s.split(".")[1]) for s in stamps]) * 100
        except IndexError:
            ts = np.asarray([int(s) for s in stamps])
        return poses, ts
    return poses
