"""GPU probe: rpe_pose_solve group sizing on real pose-head inputs (one engine chunk of the bench sequence).
    python tools/pose_probe.py [pairs]"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200  # noqa: E402,F401
from rpe_b200 import _lib, ops  # noqa: E402
from rpe_b200.core.pose.pose_estimator import PoseEstimator  # noqa: E402
from rpe_b200.dataset.synthetic import bench_sequence  # noqa: E402
from rpe_b200.engine import F2FEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
seq = bench_sequence()
L, R, M = seq.frames_u8(cache_dir=tempfile.gettempdir())
dev = torch.device("cuda:0")
ck = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
cfg = {"frame2frame": True, "dist_thr": 0.05, "depth_clipping": [1, 250], "debug": False, "conf_weighing": True, "average_pts": False,
       "lbgfs_iters": 20, "precision": "fp16x3"}
est = PoseEstimator(cfg, torch.tensor(seq.calib["intrinsics"]["left"]), seq.calib["bf"], ck if os.path.isfile(ck) else None, (640, 512)).to(dev)
eng = F2FEngine(est, chunk=n)
eng.keep_solve_inputs = True
eng.infer_sequence(*(torch.from_numpy(x[:n + 1]).to(dev) for x in (L, R, M)) if False else (torch.from_numpy(L[:n + 1]).to(dev).float(), torch.from_numpy(R[:n + 1]).to(dev).float(), torch.from_numpy(M[:n + 1]).to(dev)))
a = eng.last_solve_inputs
lib = _lib.lib()
ref = None
for bpg in (8, 16, 32, 64, 128):
    lib.rpe_pose_set_group_size(bpg)
    for _ in range(2):
        sol = ops.pose_solve(*a, max_iter=20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        sol = ops.pose_solve(*a, max_iter=20)
    e1.record()
    torch.cuda.synchronize()
    ev = float(sol.n_evals.sum())
    ms = e0.elapsed_time(e1) / 5
    same = True if ref is None else bool(torch.equal(ref, sol.out[:, :13]))
    ref = sol.out[:, :13].clone() if ref is None else ref
    print(f"pairs {a[0].shape[0]} group size {bpg:3d}: {ms:7.3f} ms, {ev / a[0].shape[0]:.1f} evals/pair, {1e3 * ms / ev:6.2f} us per pair-evaluation, "
          f"{42 * 512 * 640 * ev / ms / 1e6:7.1f} GB/s, bit-identical to the first setting: {same}")
lib.rpe_pose_set_group_size(16)
