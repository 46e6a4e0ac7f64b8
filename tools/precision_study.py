"""CPU study: which operand precision of the trunk convolutions keeps the pose gate (1e-4 rad / 1e-4 rel. translation)?

Emulates the tensor-core arithmetic of csrc/conv.cu on the oracle port (oracle/pipeline_ref.py): every RAFT convolution
sees operands rounded to a given number of planes / formats (the products themselves are exact in the tensor core and the
accumulation is fp32, so rounding the operands IS the arithmetic).  Runs the 3-frame golden sequence and prints the pose
and flow deviation from the reference's own outputs (tests/golden/e2e_384x352.npz).  Test tooling, not product code.

    python tools/precision_study.py [--golden tests/golden/e2e_384x352.npz]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pipeline_ref, se3_np  # noqa: E402
from oracle.detrand import unpack  # noqa: E402


def planes(t, dtype, n):
    """t rounded to the sum of n planes of `dtype` (n = 0: untouched fp32)."""
    if n == 0:
        return t
    acc = torch.zeros_like(t)
    r = t
    for _ in range(n):
        p = r.to(dtype).float()
        acc = acc + p
        r = t - acc
    return acc


def run(golden, act, wgt):
    """act, wgt: (dtype, n_planes)."""
    g = np.load(golden)
    W, H = [int(v) for v in g["size"]]
    sd = torch.load(os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth"), map_location="cpu", weights_only=False)["state_dict"]
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sd:
        if k.startswith("flow.") and k.endswith(".weight") and sd[k].dim() == 4:
            sd[k] = planes(sd[k].float(), *wgt)
    orig = pipeline_ref._conv

    def conv(x, sd_, name, stride=1, padding=0):
        if name.startswith("flow."):
            x = planes(x, *act)
        return orig(x, sd_, name, stride, padding)

    pipeline_ref._conv = conv
    try:
        trk = pipeline_ref.RefTracker(sd, g["K"], float(g["bf"]))
        poses = [trk.step(torch.from_numpy(g["imgs_l"][i].astype(np.float32))[None], torch.from_numpy(g["imgs_r"][i].astype(np.float32))[None],
                          torch.from_numpy(unpack(g["masks_in"][i], (1, 1, H, W)))) for i in range(3)]
    finally:
        pipeline_ref._conv = orig
    out = []
    for k in (1, 2):
        a, b = poses[k], g["traj"][k]
        d = se3_np.mul(se3_np.inv(a.astype(np.float64)), b.astype(np.float64))
        out.append((np.linalg.norm(se3_np.log(d)[3:]), np.linalg.norm(a[:3] - b[:3]) / max(np.linalg.norm(b[:3]), 1e-12)))
    epe = np.sqrt(((trk.last["time_flow"][0].numpy() - g["s_time_flow"]) ** 2).sum(0))
    return out, epe.mean(), epe.max()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", default=os.path.join(ROOT, "tests", "golden", "e2e_384x352.npz"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    bf, fh, f32 = torch.bfloat16, torch.float16, torch.float32
    variants = {
        "fp32 (oracle itself)": ((f32, 0), (f32, 0)),
        "bf16x3: act 2 x bf16, wgt 2 x bf16": ((bf, 2), (bf, 2)),
        "fp16x2: act 2 x fp16, wgt 1 x fp16": ((fh, 2), (fh, 1)),
        "fp16x2': act 1 x fp16, wgt 2 x fp16": ((fh, 1), (fh, 2)),
        "fp16x1: act 1 x fp16, wgt 1 x fp16": ((fh, 1), (fh, 1)),
        "bf16x2: act 2 x bf16, wgt 1 x bf16": ((bf, 2), (bf, 1)),
        "bf16x1": ((bf, 1), (bf, 1)),
    }
    for name, (act, wgt) in variants.items():
        if a.only and a.only not in name:
            continue
        errs, em, ex = run(a.golden, act, wgt)
        print(f"{name:40s} pose err (rot rad, rel trans): " + "  ".join(f"({r:.2e}, {t:.2e})" for r, t in errs) + f"   flow EPE mean {em:.2e} max {ex:.2e}", flush=True)
