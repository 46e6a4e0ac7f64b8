import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpe_b200
from rpe_b200 import ops, _lib
from rpe_b200.ops import _p, _stream, check
from rpe_b200.core.pose.pose_net import PoseNet
CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
ck = torch.load(CKPT, map_location="cpu", weights_only=False)
os.environ["RPE_FUSED_FLOW_HEAD"] = "1"
cfg = dict(ck["config"]["model"], image_shape=(512, 640), lbgfs_iters=20, use_weights=True, precision="fp16x3")
model = PoseNet(cfg); model.load_state_dict(ck["state_dict"]); model = model.cuda().eval()
raft = model.flow
utc = raft.update_tc()
B, h, w = 2, 64, 80
st = utc._state(B, h, w, torch.device("cuda:0"))
torch.manual_seed(0)
st["hp"].hi.normal_(); st["h"].normal_()
snap = {k: (v.clone() if torch.is_tensor(v) else (v.hi.clone(), v.lo.clone())) for k, v in st.items() if k not in ("plans", "dims", "grid")}
torch.cuda.synchronize()
st["plans"]["fh1"].run()
torch.cuda.synchronize()
for k, v in st.items():
    if k in ("plans", "dims", "grid"):
        continue
    if torch.is_tensor(v):
        ch = not torch.equal(v, snap[k])
    else:
        ch = not (torch.equal(v.hi, snap[k][0]) and torch.equal(v.lo, snap[k][1]))
    if ch:
        print("changed by fh1:", k)
l = _lib.lib()
check(l.rpe_tap_gather3x3(_p(st["fpart"]), 36, _p(utc._bias("flow_head.conv2")), _p(st["delta"]), 4, B, h, w, _stream()), "tap")
torch.cuda.synchronize()
snap2 = {k: (v.clone() if torch.is_tensor(v) else (v.hi.clone(), v.lo.clone())) for k, v in st.items() if k not in ("plans", "dims", "grid")}
print("done; fpart finite:", bool(torch.isfinite(st["fpart"]).all()), "delta max", float(st["delta"].abs().max()))
