import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200
from rpe_b200 import ops
from rpe_b200.core.pose.pose_net import PoseNet
from rpe_b200.core.RAFT.core.update import update_forward
from rpe_b200.core.RAFT.core.raft import coords_grid
CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
g = np.load(os.path.join(ROOT, "oracle", "_ref", "golden_full.npz"))
ck = torch.load(CKPT, map_location="cpu", weights_only=False)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
i1 = dev(g["imgs_l"][1:3].astype(np.float32))
i2 = torch.cat((dev(g["imgs_l"][2:3].astype(np.float32)), dev(g["imgs_r"][2:3].astype(np.float32))))
for flag in ("1", "0", "1"):
    os.environ["RPE_FUSED_FLOW_HEAD"] = flag
    cfg = dict(ck["config"]["model"], image_shape=(512, 640), lbgfs_iters=20, use_weights=True, precision="fp16x3")
    model = PoseNet(cfg); model.load_state_dict(ck["state_dict"]); model = model.cuda().eval()
    raft = model.flow
    with torch.no_grad():
        fm = raft.features(torch.cat((i1, i2), 0)); net, inp = raft.context(i1)
        f1, f2 = fm[:2].contiguous(), fm[2:].contiguous()
        # torch fp32 reference of ONE update iteration on the same inputs
        from rpe_b200.core.RAFT.core.corr import CorrBlock
        cb = CorrBlock(f1, f2, radius=4, precision=ops.CORR_F16X3)
        c0 = coords_grid(2, 64, 80, f1.device)
        corr = cb(c0)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            net_ref, _, delta_ref = update_forward(net, inp, corr, c0 - c0, raft.weights(), "update_block.", want_mask=False)
        for call in (1, 2):
            preds, n2, _, flo = raft.refine(f1, f2, net, inp, iters=1, upsample=False)
            print(f"flag {flag} call {call}: net vs torch {(n2 - net_ref).abs().max().item():.3e}   flow vs torch {(flo - delta_ref).abs().max().item():.3e}  (|delta_ref| max {delta_ref.abs().max().item():.2f})")
