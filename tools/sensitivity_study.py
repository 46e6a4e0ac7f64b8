"""CPU study: how strongly does the reference's own pipeline amplify rounding-level perturbations into the pose?

Runs the oracle port (oracle/pipeline_ref.py, fp32, pinned to the reference) on a golden sequence and multiplies the output of
every RAFT convolution by (1 + eps * N(0, 1)) -- a stand-in for "any other fp32 implementation" (different summation order,
fused multiply-adds, another BLAS).  eps = 6e-8 is one fp32 ulp.  Prints the per-pair pose deviation from the unperturbed run.
Test tooling (the product never imports it).

    python tools/sensitivity_study.py [--golden tests/golden/e2e_cfg1_tartan.npz]
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


def run(g, sd, eps, seed, lw_mask=None):
    W, H = [int(v) for v in g["size"]]
    if lw_mask is not None:                              # config 4: the 3-D-only objective (loss_weight[1] = 0)
        sd = dict(sd)
        sd["loss_weight"] = sd["loss_weight"] * torch.tensor(lw_mask, dtype=sd["loss_weight"].dtype)
    gen = torch.Generator().manual_seed(seed)
    orig = pipeline_ref._conv

    def conv(x, sd_, name, stride=1, padding=0):
        y = orig(x, sd_, name, stride, padding)
        if eps > 0 and name.startswith("flow."):
            y = y * (1.0 + eps * torch.randn(y.shape, generator=gen))
        return y

    pipeline_ref._conv = conv
    try:
        trk = pipeline_ref.RefTracker(sd, g["K"], float(g["bf"]))
        rels, evals = [], []
        for i in range(g["imgs_l"].shape[0]):
            trk.step(torch.from_numpy(g["imgs_l"][i].astype(np.float32))[None], torch.from_numpy(g["imgs_r"][i].astype(np.float32))[None],
                     torch.from_numpy(unpack(g["masks_in"][i], (1, 1, H, W))))
            if i > 0:
                rels.append(trk.last["rel"].astype(np.float64)), evals.append(trk.last["n_evals"])
    finally:
        pipeline_ref._conv = orig
    return rels, evals


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", default=os.path.join(ROOT, "tests", "golden", "e2e_cfg1_tartan.npz"))
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--ckpt", default="poseNet_2xf8up4b.pth")
    ap.add_argument("--only3d", action="store_true", help="zero the 2-D loss weight (BASELINE config 4)")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    g = dict(np.load(a.golden))
    W, H = [int(v) for v in g["size"]]
    if "order" in g:
        order = [int(k) for k in g["order"]]
        g["imgs_l"], g["imgs_r"] = g["imgs_l"][order], g["imgs_r"][order]
        g["masks_in"] = np.stack([np.packbits(np.ones(H * W, bool))] * len(order))
    sd = torch.load(os.path.join(ROOT, "oracle", "_ref", "trained", a.ckpt), map_location="cpu", weights_only=False)["state_dict"]
    lw = (1.0, 0.0) if a.only3d else None
    base, ev0 = run(g, sd, 0.0, 0, lw)
    print("unperturbed evaluations per pair:", ev0)
    for eps in (6e-8, 1e-6, 1e-5):
        for seed in range(a.seeds):
            rels, ev = run(g, sd, eps, seed, lw)
            out = []
            for r, b in zip(rels, base):
                d = se3_np.mul(se3_np.inv(r), b)
                out.append((np.linalg.norm(se3_np.log(d)[3:]), np.linalg.norm(r[:3] - b[:3]) / np.linalg.norm(b[:3])))
            print(f"eps {eps:.0e} seed {seed}: " + "  ".join(f"pair {k}: rot {ro:.2e} rad, rel. trans {tr:.2e}" for k, (ro, tr) in enumerate(out)) + f"  evals {ev}",
                  flush=True)
