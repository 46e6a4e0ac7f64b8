"""GPU probe: BASELINE config 4 as a throughput figure -- only3d checkpoint (3-D residual only) on full-resolution 1280x1024 stereo,
through PoseEstimator.infer_sequence.  Frames: the two frames of oracle/_ref/golden_only3d_1280x1024.npz, alternating.
    python tools/config4_probe.py [pairs] [chunk]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200  # noqa: E402,F401
from rpe_b200.core.pose.pose_estimator import PoseEstimator  # noqa: E402
from rpe_b200.lie import SE3  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
g = np.load(os.path.join(ROOT, "oracle", "_ref", "golden_only3d_1280x1024.npz"))
W, H = [int(v) for v in g["size"]]
dev = torch.device("cuda:0")
idx = [k % 2 for k in range(pairs + 1)]
L = torch.from_numpy(g["imgs_l"])[idx].to(dev)
R = torch.from_numpy(g["imgs_r"])[idx].to(dev)
M = torch.from_numpy(np.stack([np.unpackbits(g["masks_in"][i])[:H * W].astype(bool).reshape(1, H, W) for i in range(2)]))[idx].to(dev)
cfg = {"frame2frame": True, "dist_thr": 0.05, "depth_clipping": [1, 250], "debug": False, "conf_weighing": True, "average_pts": False,
       "lbgfs_iters": 20, "precision": "fp16x3", "residuals": "3d"}
est = PoseEstimator(cfg, torch.tensor(g["K"]), float(g["bf"]), os.path.join(ROOT, "oracle", "_ref", "trained", "only3d_1a7ix98y.pth"), (W, H)).to(dev)
for rep in range(3):
    est.last_pose = SE3.Identity(1, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    traj, failed = est.infer_sequence(L, R, M, chunk=chunk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"only3d 1280x1024, {pairs} pairs, chunk {chunk}: {ms:8.1f} ms = {pairs / ms * 1e3:6.1f} pairs/s, failed {int(failed.sum())}, "
          f"peak memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")
