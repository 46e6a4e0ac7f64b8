"""Development probe: CUDA-event timings of each stage kernel at the bench shape (640x512, batch 2)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpe_b200  # noqa: E402,F401
from rpe_b200 import ops, _lib  # noqa: E402


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0), "SMs", _lib.lib().rpe_device_sm_count())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    H, W = 512, 640
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.rand(*s, device=dev, generator=g)
    # ---- geometry / warp
    sflow = -30 * r(1, 2, H, W) - 5
    K = torch.tensor([[[400.0, 0, 320], [0, 400.0, 256], [0, 0, 1]]], device=dev)
    bf = torch.tensor([8.8], device=dev)
    mask = torch.ones(1, 1, H, W, dtype=torch.bool, device=dev)
    med, mn = timeit(lambda: ops.depth_proj(sflow, bf, K, mask), flush=flush)
    print(f"depth_proj        {med*1e3:8.1f} us (min {mn*1e3:.1f})  {23*H*W/med/1e6:7.1f} GB/s alg")
    depth, valid, pcl = ops.depth_proj(sflow, bf, K, mask)
    img = 255 * r(1, 3, H, W)
    flow = 10 * (r(1, 2, H, W) - 0.5)
    med, mn = timeit(lambda: ops.warp8_mask(pcl, img, sflow, mask, flow), flush=flush)
    print(f"warp8_mask        {med*1e3:8.1f} us (min {mn*1e3:.1f})  {74*H*W/med/1e6:7.1f} GB/s alg")
    # ---- pose
    for n in (1, 8, 32):
        fl = (4 * (r(n, 2, H, W) - 0.5)).contiguous()
        p1 = pcl.repeat(n, 1, 1, 1).contiguous()
        p2 = (p1 + 0.001 * (r(n, 3, H, W) - 0.5)).contiguous()
        w1, w2 = r(n, 1, H, W), r(n, 1, H, W)
        m = torch.ones(n, 1, H, W, dtype=torch.bool, device=dev)
        Kn = K.repeat(n, 1, 1).contiguous()
        lw = torch.tensor([[0.93, 1.0]], device=dev).repeat(n, 1).contiguous()
        for groups in ((1,) if n == 1 else (1, 4, 8, 16)):
            _lib.lib().rpe_pose_set_groups(groups)
            sol = ops.pose_solve(fl, p1, p2, w1, w2, m, m, Kn, lw, max_iter=20)
            ev = sol.n_evals.sum().item()
            med, mn = timeit(lambda: ops.pose_solve(fl, p1, p2, w1, w2, m, m, Kn, lw, max_iter=20), iters=10, flush=flush)
            print(f"pose_solve n={n:3d} groups={groups:2d}  {med*1e3:9.1f} us  evals={ev:.0f}  {med*1e3/ev:7.2f} us/eval  "
                  f"{42*H*W*ev/med/1e6:8.1f} GB/s alg  per pair {med*1e3/n:8.1f} us")
        sol = ops.pose_solve(fl, p1, p2, w1, w2, m, m, Kn, lw, mode=ops.SOLVER_GN, max_iter=10)
        med, mn = timeit(lambda: ops.pose_solve(fl, p1, p2, w1, w2, m, m, Kn, lw, mode=ops.SOLVER_GN, max_iter=10), iters=10)
        print(f"pose_gn    n={n:3d}            {med*1e3:9.1f} us  evals={sol.n_evals.sum().item():.0f}")
    _lib.lib().rpe_pose_set_groups(8)
    # ---- correlation
    B, C, h, w = 2, 256, 64, 80
    f1, f2 = r(B, C, h, w) - 0.5, r(B, C, h, w) - 0.5
    for prec, name in ((ops.CORR_TF32, "tf32"), (ops.CORR_TF32X3, "tf32x3")):
        med, mn = timeit(lambda: ops.CorrPyramid(f1, f2, precision=prec), flush=flush)
        Q = h * w
        fl_ = 2 * B * Q * Q * C * (3 if prec else 1)
        print(f"corr_build {name:7s} {med*1e3:8.1f} us (min {mn*1e3:.1f})  {fl_/med/1e9:7.1f} TFLOP/s  "
              f"{B*Q*Q*4*1.328/med/1e6:7.1f} GB/s pyramid-write")
    cp = ops.CorrPyramid(f1, f2)
    coords = torch.stack(torch.meshgrid(torch.arange(w, device=dev), torch.arange(h, device=dev), indexing="xy"), 0).float()
    coords = (coords[None].repeat(B, 1, 1, 1) + 8 * (r(B, 2, h, w) - 0.5)).contiguous()
    med, mn = timeit(lambda: cp(coords), flush=flush)
    print(f"corr_lookup       {med*1e3:8.1f} us (min {mn*1e3:.1f})  {B*14.87e6/med/1e6:7.1f} GB/s alg")
    med, mn = timeit(lambda: cp(coords))
    print(f"corr_lookup (L2 warm) {med*1e3:8.1f} us")
    # torch reference ops for context
    a = f1.view(B, C, -1)
    med, mn = timeit(lambda: torch.matmul(a.transpose(1, 2), f2.view(B, C, -1)), flush=flush)
    print(f"torch.matmul fp32 (cuBLAS, context only) {med*1e3:8.1f} us")
    mk = r(B, 576, h, w)
    fl8 = r(B, 2, h, w)
    med, mn = timeit(lambda: ops.convex_upsample8(fl8, mk), flush=flush)
    print(f"convex_upsample8  {med*1e3:8.1f} us")


if __name__ == "__main__":
    main()
