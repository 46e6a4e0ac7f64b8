"""Multi-GPU probe (torchrun): parallel.infer_sequence_sharded over NCCL with fewer pairs than ranks and with ragged shards; every
rank must return the single-process trajectory bit for bit.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_edge_probe.py"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200  # noqa: E402,F401
from rpe_b200 import parallel  # noqa: E402
from rpe_b200.core.pose.pose_estimator import PoseEstimator  # noqa: E402
from rpe_b200.dataset.synthetic import bench_sequence  # noqa: E402
from rpe_b200.lie import SE3  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
seq = bench_sequence()
L, R, M = seq.frames_u8(cache_dir=tempfile.gettempdir())
ck = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
cfg = {"frame2frame": True, "dist_thr": 0.05, "depth_clipping": [1, 250], "debug": False, "conf_weighing": True, "average_pts": False,
       "lbgfs_iters": 20, "precision": "fp16x3"}
est = PoseEstimator(cfg, torch.tensor(seq.calib["intrinsics"]["left"]), seq.calib["bf"], ck, (640, 512)).to(dev)
ok = True
for n_frames in (1, 2, 3, 6, 11):
    Lh, Rh, Mh = (torch.from_numpy(x[:n_frames]) for x in (L, R, M))
    est.last_pose = SE3.Identity(1, device=dev)
    traj, failed = parallel.infer_sequence_sharded(est, lambda a, b: (Lh[a:b].to(dev), Rh[a:b].to(dev), Mh[a:b].to(dev)), n_frames, chunk=4)
    est.last_pose = SE3.Identity(1, device=dev)
    ref, failed_ref = est.infer_sequence(Lh.to(dev), Rh.to(dev), Mh.to(dev), chunk=4)
    same = bool(torch.equal(traj, ref)) and bool(torch.equal(failed, failed_ref))
    ok = ok and same and tuple(traj.shape) == (n_frames, 7)
    print(f"rank {rank}/{world}: {n_frames:2d} frames -> trajectory {tuple(traj.shape)}, equal to the single-process run: {same}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED EDGE PROBE", "OK" if int(flag) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
