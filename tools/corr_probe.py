"""GPU probe: rpe_corr_build (NCHW fp32 in) and rpe_corr_build_planes (NHWC planes in) at the bench batch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200  # noqa: E402,F401
from rpe_b200 import ops, tc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
f1, f2 = torch.randn(B, 256, 64, 80, device=dev), torch.randn(B, 256, 64, 80, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


alg = B * (5120 * 5120 * 4 * (1 + 1 / 4 + 1 / 16 + 1 / 64))
for name, prec in (("tf32", ops.CORR_TF32), ("tf32x3", ops.CORR_TF32X3), ("fp16x3 (NCHW in)", ops.CORR_F16X3)):
    ms = timeit(lambda: ops.CorrPyramid(f1, f2, precision=prec))
    print(f"corr_build {name:18s} B={B}: {ms:7.3f} ms  {alg / ms / 1e6:7.1f} GB/s pyramid write  {2 * B * 5120 * 5120 * 256 / ms / 1e9:7.1f} TFLOP/s algorithmic")
feat = tc.Planes(2 * B, 64, 80, 256, dev)
feat.hi.normal_()
feat.lo.normal_(std=1e-3)
ms = timeit(lambda: ops.CorrPyramid.from_planes(feat, feat.view(B), B))
print(f"corr_build planes in       B={B}: {ms:7.3f} ms  {alg / ms / 1e6:7.1f} GB/s pyramid write  {2 * B * 5120 * 5120 * 256 / ms / 1e9:7.1f} TFLOP/s algorithmic")

# sustained: the same launch back to back for ~1.5 s (the bench runs under the power cap; a 20 ms burst does not)
def sustained(fn, seconds=1.5):
    fn(); torch.cuda.synchronize()
    import time
    t_end = time.time() + seconds
    n = 0
    while time.time() < t_end:
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        n += 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20, n


ms, n = sustained(lambda: ops.CorrPyramid.from_planes(feat, feat.view(B), B))
print(f"corr_build planes in, SUSTAINED after {n} back-to-back launches: {ms:7.3f} ms  {alg / ms / 1e6:7.1f} GB/s pyramid write")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
