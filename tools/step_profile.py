"""Kernel-level time breakdown of one engine step (torch.profiler / CUPTI; no ncu replay cost).
    python tools/step_profile.py [--precision fp16x3] [--pairs 16] [--chunk 8]"""
import argparse
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16x3")
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--top", type=int, default=45)
    args = ap.parse_args()
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    from rpe_b200.engine import F2FEngine
    dev = torch.device("cuda", 0)
    L, R, M, seq = bench.synthetic_sequence(args.pairs + 1)
    est = PoseEstimator(dict(bench.SLAM, precision=args.precision), torch.tensor(seq.calib["intrinsics"]["left"]), seq.calib["bf"],
                        bench.CKPT, (bench.W_IMG, bench.H_IMG)).to(dev)
    dL, dR, dM = (torch.from_numpy(a).to(dev) for a in (L, R, M))
    dL, dR = dL.float(), dR.float()
    eng = F2FEngine(est, chunk=args.chunk)
    for _ in range(2):
        eng.reset()
        eng.infer_sequence(dL, dR, dM)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.reset()
    eng.infer_sequence(dL, dR, dM)
    e1.record()
    torch.cuda.synchronize()
    print(f"step (unprofiled): {e0.elapsed_time(e1):.2f} ms for {args.pairs} pairs")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        eng.reset()
        eng.infer_sequence(dL, dR, dM)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"<.*", "", ev.name)[:80]
            agg[name][0] += 1
            agg[name][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"total device time {tot / 1e3:.2f} ms in {sum(v[0] for v in agg.values())} kernels/memcpys")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:5d}  avg={t / n:8.1f} us  {k}")


if __name__ == "__main__":
    main()
