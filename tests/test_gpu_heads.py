"""GPU parity of the confidence heads on the tcgen05 convolution kernels (core/unet/unet_tc.py, csrc/heads.cu) against the plain
fp32 torch evaluation of the same TinyUNets (core/unet/unet.py, reference /root/reference/core/unet/unet.py:8-82 +
pose_net.py:110-115): valid-region bookkeeping, BatchNorm folding on both sides of the ReLU, transposed convolutions as 1x1
convolutions, merged first layer with a zero-weight slot, bilinear resize + sigmoid."""
import os

import numpy as np
import pytest
import torch

from oracle.detrand import det_uniform

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")


def _model(H, W, trained):
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_net import PoseNet
    torch.manual_seed(3)
    m = PoseNet({"image_shape": (H, W), "use_weights": True, "lbgfs_iters": 20, "small": False, "dropout": 0.0, "precision": "fp16x3"})
    if trained and os.path.isfile(CKPT):
        m.load_state_dict(torch.load(CKPT, map_location="cpu", weights_only=False)["state_dict"])
    else:                                                  # random weights with non-trivial BatchNorm statistics
        with torch.no_grad():
            for k, v in m.state_dict().items():
                if "weight_head" in k and "running_mean" in k:
                    v.copy_(torch.randn_like(v) * 0.3)
                if "weight_head" in k and "running_var" in k:
                    v.copy_(torch.rand_like(v) + 0.5)
                if "weight_head" in k and "norm.weight" in k:
                    v.copy_(torch.rand_like(v) + 0.5)
                if "weight_head" in k and k.endswith(".bias"):
                    v.copy_(torch.randn_like(v) * 0.1)
    return m.cuda().eval()


@pytest.mark.parametrize("H,W,n,trained", [(512, 640, 2, True), (352, 384, 3, False), (1024, 1280, 1, False)])
def test_heads_match_fp32_torch(H, W, n, trained):
    from rpe_b200 import ops
    from rpe_b200.core.unet.unet import tiny_unet_forward
    m = _model(H, W, trained)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    h8, w8 = H // 8, W // 8
    # (randomly initialised heads are not trained for 0..255 images and -30 px flows: keep their inputs O(1) so that the logits
    # stay O(10) and the comparison is not dominated by cancellation at |activations| ~ 1e3)
    smax, imax = (30.0, 255.0) if trained else (1.0, 1.0)
    sflow1, img1, pcl1 = dev(det_uniform((n, 2, H, W), 1, -smax, 2)), dev(det_uniform((n, 3, H, W), 2, 0, imax)), dev(det_uniform((n, 3, H, W), 3, -1, 1))
    sflow2w, img2w, pcl2w = dev(det_uniform((n, 2, H, W), 4, -smax, 2)), dev(det_uniform((n, 3, H, W), 5, 0, imax)), dev(det_uniform((n, 3, H, W), 6, -1, 1))
    gru, ctx = dev(det_uniform((n, 128, h8, w8), 7, -1, 1)), dev(det_uniform((n, 128, h8, w8), 8, 0, 2))
    # ---- product path: get_weight_maps' head branch through the C ABI
    from rpe_b200.core.unet.unet_tc import HeadsTC
    from rpe_b200.tc import Planes, nchw_to_planes
    heads = HeadsTC(m._head_weights())
    gp, cp = Planes(n, h8, w8, 128, gru.device), Planes(n, h8, w8, 128, gru.device)
    nchw_to_planes(gru, gp), nchw_to_planes(ctx, cp)
    c1, c2 = heads.forward([sflow1, img1, pcl1], [sflow2w, img2w, pcl2w], gp, cp, n, (H, W))
    # ---- fp32 torch evaluation of the same graph (the reference's operator sequence)
    x3 = torch.empty((n, 272, h8, w8), device=gru.device)
    ops.downsample8_cat([sflow1, img1, pcl1], out=x3, ch_offset=0)
    ops.downsample8_cat([sflow2w, img2w, pcl2w], out=x3, ch_offset=8)
    x3[:, 16:144], x3[:, 144:] = gru, ctx
    x2 = torch.cat((x3[:, :8], x3[:, 16:]), 1)
    Wt = m._head_weights()
    with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
        l1 = tiny_unet_forward(x2, Wt, "weight_head_2d.0.", (H, W))
        l2 = tiny_unet_forward(x3, Wt, "weight_head_3d.0.", (H, W))
    r1, r2 = torch.sigmoid(l1), torch.sigmoid(l2)
    e1, e2 = float((c1 - r1).abs().max()), float((c2 - r2).abs().max())
    # logit-space error (the sigmoid flattens differences where it saturates)
    lg = lambda c: torch.log(c.clamp(1e-6, 1 - 1e-6) / (1 - c.clamp(1e-6, 1 - 1e-6)))
    sel = (r1 > 1e-3) & (r1 < 1 - 1e-3)
    el = float((lg(c1) - l1)[sel].abs().max()) if bool(sel.any()) else 0.0
    print(f"\nheads {H}x{W} n={n}: conf max abs err {e1:.2e} / {e2:.2e}, logit err {el:.2e} (|logit| max {float(l1.abs().max()):.1f})")
    if trained:
        assert e1 < 2e-5 and e2 < 2e-5
    else:
        # untrained heads push |activations| to ~1e2 .. 1e3 and the logits are differences of such numbers: the bound is relative to
        # the largest logit (fp32 evaluations of this graph in another summation order differ by the same amount)
        assert el <= 5e-5 * float(l1.abs().max()) and e1 < 2e-3 and e2 < 2e-3
    assert c1.shape == (n, 1, H, W) and c2.shape == (n, 1, H, W)
