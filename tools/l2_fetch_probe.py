"""GPU probe: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 B) against the windowed gather of the correlation lookup (40-byte rows)
and two streaming kernels that must not suffer (norm_act, the correlation volume).  python tools/l2_fetch_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200  # noqa: E402,F401
from rpe_b200 import _lib, ops, tc  # noqa: E402
from rpe_b200.ops import _p, _stream, check  # noqa: E402

B = 64
dev = torch.device("cuda:0")
l = _lib.lib()
feat = tc.Planes(2 * B, 64, 80, 256, dev)
feat.hi.normal_()
feat.lo.normal_(std=1e-3)
pyr = ops.CorrPyramid.from_planes(feat, feat.view(B), B)
coords = torch.stack(torch.meshgrid(torch.arange(64.0), torch.arange(80.0), indexing="ij")[::-1]).to(dev)[None].repeat(B, 1, 1, 1)
coords = (coords + 6.0 * torch.randn_like(coords)).contiguous()
out = tc.Planes(B, 64, 80, 384, dev)
a = torch.randn(B, 128, 160, 96, device=dev)
stats = torch.rand(B, 96, 2, device=dev) + 0.5
pl = tc.Planes(B, 128, 160, 96, dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("granularity in force at start:", l.rpe_l2_fetch_granularity(0))
import time
t_end = time.time() + 3.0                                  # reach the steady (power-capped) state first
while time.time() < t_end:
    ops.CorrPyramid.from_planes(feat, feat.view(B), B)
    ops.corr_lookup_planes(pyr, coords, out)
    torch.cuda.synchronize()
for rep in range(5):
    for g in (64, 32):
        got = l.rpe_l2_fetch_granularity(g)
        t_lk = timeit(lambda: ops.corr_lookup_planes(pyr, coords, out), n=30)
        t_na = timeit(lambda: check(l.rpe_norm_act_split_res(_p(a), _p(stats), 1, None, None, None, None, 0, None, _p(pl.hi), _p(pl.lo), 96, B,
                                                             128 * 160, 96, _stream()), "norm_act"))
        t_cb = timeit(lambda: ops.CorrPyramid.from_planes(feat, feat.view(B), B), n=10)
        print(f"rep {rep} requested {g:3d} B -> in force {got:3d} B: lookup {t_lk:7.1f} us   norm_act {t_na:7.1f} us   corr_build {t_cb:8.1f} us", flush=True)
