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
    for k in range(n):
        p = r.to(dtype[k] if isinstance(dtype, tuple) else dtype).float()
        acc = acc + p
        r = t - acc
    return acc


def run(golden, act, wgt, corr=None, only_layers=None):
    """act, wgt: (dtype, n_planes).  corr: (dtype, n_planes) rounding of the feature maps entering the all-pairs correlation
    (None = fp32 matmul).  only_layers: substring filter -- only convolutions whose name contains one of them are rounded."""
    g = dict(np.load(golden))
    W, H = [int(v) for v in g["size"]]
    if "order" in g:                                   # config-1 fixture golden: two distinct frames + the frame order, masks all true
        order = [int(k) for k in g["order"]]
        g["imgs_l"], g["imgs_r"] = g["imgs_l"][order], g["imgs_r"][order]
        g["masks_in"] = np.stack([np.packbits(np.ones(H * W, bool))] * len(order))
        g["s_time_flow"] = None
    sel = (lambda name: True) if not only_layers else (lambda name: any(t in name for t in only_layers))
    sd = torch.load(os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth"), map_location="cpu", weights_only=False)["state_dict"]
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sd:
        if k.startswith("flow.") and k.endswith(".weight") and sd[k].dim() == 4 and sel(k):
            sd[k] = planes(sd[k].float(), *wgt)
    orig = pipeline_ref._conv

    def conv(x, sd_, name, stride=1, padding=0):
        if name.startswith("flow.") and sel(name):
            x = planes(x, *act)
        return orig(x, sd_, name, stride, padding)

    orig_pyr = pipeline_ref.corr_pyramid

    def pyr(f1, f2, levels=4):
        return orig_pyr(planes(f1, *corr), planes(f2, *corr), levels)

    pipeline_ref._conv = conv
    if corr is not None:
        pipeline_ref.corr_pyramid = pyr
    try:
        trk = pipeline_ref.RefTracker(sd, g["K"], float(g["bf"]))
        poses = [trk.step(torch.from_numpy(g["imgs_l"][i].astype(np.float32))[None], torch.from_numpy(g["imgs_r"][i].astype(np.float32))[None],
                          torch.from_numpy(unpack(g["masks_in"][i], (1, 1, H, W)))) for i in range(3)]
    finally:
        pipeline_ref._conv = orig
        pipeline_ref.corr_pyramid = orig_pyr
    out = []
    for k in (1, 2):
        a, b = poses[k], g["traj"][k]
        d = se3_np.mul(se3_np.inv(a.astype(np.float64)), b.astype(np.float64))
        out.append((np.linalg.norm(se3_np.log(d)[3:]), np.linalg.norm(a[:3] - b[:3]) / max(np.linalg.norm(b[:3]), 1e-12)))
    if g["s_time_flow"] is None:
        epe = np.sqrt(((trk.last["time_flow"][0, :, ::4, ::4].numpy() - g["s_time_flow_ds4"]) ** 2).sum(0))
        # trajectory error of frame 2 is meaningless when the sequence returns to its start: report the per-pair relative pose
        a, b = trk.last["rel"], g["rel_pose"][1]
        d = se3_np.mul(se3_np.inv(a.astype(np.float64)), b.astype(np.float64))
        out[1] = (np.linalg.norm(se3_np.log(d)[3:]), np.linalg.norm(a[:3] - b[:3]) / max(np.linalg.norm(b[:3]), 1e-12))
    else:
        epe = np.sqrt(((trk.last["time_flow"][0].numpy() - g["s_time_flow"]) ** 2).sum(0))
    return out, epe.mean(), epe.max()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", default=os.path.join(ROOT, "tests", "golden", "e2e_384x352.npz"))
    ap.add_argument("--only", default="")
    ap.add_argument("--corr", action="store_true", help="also round the feature maps of the correlation to 2 bf16 planes")
    ap.add_argument("--layers", default="", help="comma-separated substrings: only these convolutions are rounded (the others stay fp32)")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    bf, fh, f32 = torch.bfloat16, torch.float16, torch.float32
    variants = {
        "fp32 (oracle itself)": ((f32, 0), (f32, 0)),
        "bf16x3: act 2 x bf16, wgt 2 x bf16": ((bf, 2), (bf, 2)),
        "f16+bf16 x3: fp16 hi + bf16 lo planes": (((fh, bf), 2), ((fh, bf), 2)),
        "fp16x3: act 2 x fp16, wgt 2 x fp16": ((fh, 2), (fh, 2)),
        "fp16x2: act 2 x fp16, wgt 1 x fp16": ((fh, 2), (fh, 1)),
        "fp16x2': act 1 x fp16, wgt 2 x fp16": ((fh, 1), (fh, 2)),
        "fp16x1: act 1 x fp16, wgt 1 x fp16": ((fh, 1), (fh, 1)),
        "bf16x2: act 2 x bf16, wgt 1 x bf16": ((bf, 2), (bf, 1)),
        "bf16x1": ((bf, 1), (bf, 1)),
    }
    for name, (act, wgt) in variants.items():
        if a.only and a.only not in name:
            continue
        errs, em, ex = run(a.golden, act, wgt, corr=(act[0], 2) if a.corr and act[1] else None,
                           only_layers=[t for t in a.layers.split(",") if t] or None)
        print(f"{name:40s} pose err (rot rad, rel trans): " + "  ".join(f"({r:.2e}, {t:.2e})" for r, t in errs) + f"   flow EPE mean {em:.2e} max {ex:.2e}", flush=True)
