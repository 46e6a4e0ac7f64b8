"""CUDA-event timing of the tcgen05 convolution plans at the shapes of the RAFT trunk (one chunk = 16 RAFT samples at 64x80,
24 encoder images).  Prints effective TFLOP/s (useful flops x3 for the fp16x3 split) per shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpe_b200  # noqa: E402,F401
from rpe_b200 import tc  # noqa: E402

SHAPES = [
    # name, N, H, W, [cin per source], cout, kh, kw, stride
    ("convc1 1x1 324->256", 16, 64, 80, [324], 256, 1, 1, 1),
    ("convc2 3x3 256->192", 16, 64, 80, [256], 192, 3, 3, 1),
    ("convf2 3x3 128->64", 16, 64, 80, [128], 64, 3, 3, 1),
    ("conv   3x3 256->126", 16, 64, 80, [256], 126, 3, 3, 1),
    ("zr1    1x5 256->256", 16, 64, 80, [128, 128], 256, 1, 5, 1),
    ("q1     1x5 256->128", 16, 64, 80, [128, 128], 128, 1, 5, 1),
    ("zr2    5x1 256->256", 16, 64, 80, [128, 128], 256, 5, 1, 1),
    ("q2     5x1 256->128", 16, 64, 80, [128, 128], 128, 5, 1, 1),
    ("fh1    3x3 128->256", 16, 64, 80, [128], 256, 3, 3, 1),
    ("fh2    3x3 256->2", 16, 64, 80, [256], 2, 3, 3, 1),
    ("mask2  1x1 256->576", 16, 64, 80, [256], 576, 1, 1, 1),
    ("enc stem 1x1 160->64", 16, 256, 320, [160], 64, 1, 1, 1),
    ("enc l1 3x3 64->64", 16, 256, 320, [64], 64, 3, 3, 1),
    ("enc l2a 3x3/2 64->96", 16, 256, 320, [64], 96, 3, 3, 2),
    ("enc l2 3x3 96->96", 16, 128, 160, [96], 96, 3, 3, 1),
    ("enc l3a 3x3/2 96->128", 16, 128, 160, [96], 128, 3, 3, 2),
    ("enc l3 3x3 128->128", 16, 64, 80, [128], 128, 3, 3, 1),
    ("enc out 1x1 128->256", 16, 64, 80, [128], 256, 1, 1, 1),
]


def main():
    dev = torch.device("cuda")
    torch.manual_seed(0)
    single = "--single" in sys.argv
    only = [a[7:] for a in sys.argv if a.startswith("--only=")]
    reps = 1 if "--once" in sys.argv else 20
    n_over = [int(a[4:]) for a in sys.argv if a.startswith("--n=")]
    for name, n, h, w, cins, cout, kh, kw, stride in SHAPES:
        n = n_over[0] if n_over else n
        if only and not any(o in name for o in only):
            continue
        cout_pad = (cout + 15) // 16 * 16
        srcs = []
        for cin in cins:
            pl = tc.Planes(n, h, w, (cin + 15) // 16 * 16, dev)
            pl.hi.normal_()
            pl.lo.normal_(std=1e-3)
            wt = tc.pack_weight(torch.randn(cout, cin, kh, kw, device=dev) / (cin * kh * kw) ** 0.5, 0, cin, cout_pad, scale=1024.0)
            srcs.append((pl, 0, cin, wt))
        oh, ow = (h + 2 * (kh // 2) - kh) // stride + 1, (w + 2 * (kw // 2) - kw) // stride + 1
        outp = tc.Planes(n, oh, ow, cout_pad, dev)
        extra = {}
        act = "relu"
        if name.startswith("zr") and not single:        # fused GRU epilogues as used by the update operator
            hbuf = torch.randn(n, oh, ow, 128, device=dev)
            outp = tc.Planes(n, oh, ow, 128, dev)
            extra = dict(mode=1, aux=hbuf, out_f32=torch.zeros(n, oh, ow, 128, device=dev), pre=torch.randn(n, oh, ow, 256, device=dev))
            act = "sigmoid"
        elif name.startswith("q") and not single:
            extra = dict(mode=2, aux=torch.randn(n, oh, ow, 128, device=dev), aux2=torch.rand(n, oh, ow, 128, device=dev),
                         pre=torch.randn(n, oh, ow, 128, device=dev))
            act = "tanh"
        plan = tc.ConvPlan(name, srcs, (n, h, w), kh, kw, cout, act, bias=torch.zeros(cout, device=dev), stride=stride,
                           out_planes=outp, single_pass=single, **extra)
        for _ in range(0 if reps == 1 else 3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"{name:24s} {us:9.1f} us   {plan.flops / us / 1e6:8.1f} TFLOP/s (tensor)   {plan.flops / 1e9:7.1f} GFLOP")


if __name__ == "__main__":
    main()
