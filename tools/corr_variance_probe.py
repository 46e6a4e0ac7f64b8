"""GPU probe: per-launch duration of the correlation-volume kernel inside the batched engine, with the address of the pyramid it
writes and the free memory at that moment -- to see what the slow launches (4.7-5.6 ms against 3.1-3.5) have in common.
    python tools/corr_variance_probe.py [steps]"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200  # noqa: E402,F401
from rpe_b200 import ops  # noqa: E402
from rpe_b200.core.pose.pose_estimator import PoseEstimator  # noqa: E402
from rpe_b200.dataset.synthetic import bench_sequence  # noqa: E402
from rpe_b200.engine import F2FEngine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
seq = bench_sequence()
L, R, M = seq.frames_u8(cache_dir=tempfile.gettempdir())
dev = torch.device("cuda:0")
ck = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
cfg = {"frame2frame": True, "dist_thr": 0.05, "depth_clipping": [1, 250], "debug": False, "conf_weighing": True, "average_pts": False,
       "lbgfs_iters": 20, "precision": "fp16x3"}
est = PoseEstimator(cfg, torch.tensor(seq.calib["intrinsics"]["left"]), seq.calib["bf"], ck if os.path.isfile(ck) else None, (640, 512)).to(dev)
eng = F2FEngine(est, chunk=32)
Ld, Rd, Md = (torch.from_numpy(x).to(dev) for x in (L, R, M))

records = []
orig = ops.CorrPyramid.from_planes.__func__


def wrapped(cls, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pyr = orig(cls, *a, **k)
    e1.record()
    free, _ = torch.cuda.mem_get_info()
    records.append((e0, e1, pyr.pyramid.data_ptr(), pyr.pyramid.numel() * 4, free, torch.cuda.memory_reserved()))
    return pyr


ops.CorrPyramid.from_planes = classmethod(wrapped)
for s in range(steps):
    eng.reset()
    eng.infer_sequence(Ld, Rd, Md)
torch.cuda.synchronize()
for k, (e0, e1, ptr, nbytes, free, reserved) in enumerate(records):
    print(f"launch {k:2d}: {e0.elapsed_time(e1):7.3f} ms  pyramid {nbytes / 1e9:5.2f} GB at 0x{ptr:x} (offset in 2 MB pages {ptr % (1 << 21)}), "
          f"free {free / 1e9:6.1f} GB, reserved {reserved / 1e9:6.1f} GB")
