"""GPU probe: does the duration of the correlation-volume kernel depend on WHERE its 8.9 GB pyramid lives?
One 40 GB arena, the pyramid placed at several offsets (alignments from 256 B to 1 GB), 6 launches each, plus separately
cudaMalloc'ed buffers.  python tools/corr_align_probe.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200  # noqa: E402,F401
from rpe_b200 import _lib, tc  # noqa: E402
from rpe_b200.ops import _stream, check  # noqa: E402

B = 64
dev = torch.device("cuda:0")
l = _lib.lib()
feat = tc.Planes(2 * B, 64, 80, 256, dev)
feat.hi.normal_()
feat.lo.normal_(std=1e-3)
f2 = feat.view(B)
nbytes = l.rpe_corr_pyramid_bytes(B, 64, 80, 4)


def run(ptr, n=6):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(l.rpe_corr_build_planes(C.c_void_p(feat.hi.data_ptr()), C.c_void_p(feat.lo.data_ptr()), C.c_void_p(f2.hi.data_ptr()),
                                      C.c_void_p(f2.lo.data_ptr()), C.c_void_p(ptr), B, 256, 64, 80, 4, 0, 0, _stream()), "corr")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts


arena = torch.empty(40 << 30, dtype=torch.uint8, device=dev)
base = arena.data_ptr()
print(f"arena at 0x{base:x} ({nbytes / 1e9:.2f} GB pyramid)")
run(base, 2)
G = 1 << 30
for off in (0, 256, 4096, 2 << 20, 64 << 20, 66 << 20, 512 << 20, G, G + (64 << 20), 5 * G, 20 * G, 30 * G):
    a = (base + off)
    ts = run(a)
    print(f"offset {off / (1 << 20):10.3f} MB  addr 0x{a:x}: " + " ".join(f"{t:6.3f}" for t in ts) + " ms")
del arena
torch.cuda.empty_cache()
bufs = [torch.empty(nbytes + (i << 20), dtype=torch.uint8, device=dev) for i in range(8)]          # eight separate cudaMallocs, all alive
for b in bufs:
    ts = run(b.data_ptr())
    print(f"cudaMalloc'ed buffer 0x{b.data_ptr():x}: " + " ".join(f"{t:6.3f}" for t in ts) + " ms")

print("round-robin over the eight buffers (3 launches each), 5 rounds: is a slow buffer always slow?")
for r in range(5):
    print(f"round {r}: " + "  ".join(f"{sum(run(b.data_ptr(), 3)) / 3:5.2f}" for b in bufs))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
