"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's whole per-frame f2f path, used
 (a) as the end-to-end checker of the CUDA path where /root/reference is not available (the GPU box), and
 (b) as the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`; kind = "port").
Never imported by the product package.

Follows, op for op, the reference's CPU fp32 execution:
  RAFT.forward             /root/reference/core/RAFT/core/raft.py:79-137
  BasicEncoder/ResidualBlock  core/RAFT/core/extractor.py:6-56,118-192
  BasicUpdateBlock & parts core/RAFT/core/update.py:6-14,33-60,79-97,114-136
  CorrBlock                core/RAFT/core/corr.py:12-60 (torch.matmul, avg_pool2d, grid_sample)
  TinyUNet                 core/unet/unet.py:8-82
  PoseNet.infer            core/pose/pose_net.py:60-125
  PoseEstimator (f2f)      core/pose/pose_estimator.py:50-125
  DPoseSE3Head.solve       core/pose/pose_head.py:60-79 -> oracle/pose_np.lbfgs_solve (fp64)
Pinned by tests/test_oracle_pipeline.py against tests/golden/e2e_384x352.npz (outputs of the reference itself).

``RefTracker(device="cuda", autocast=True, solver="torch")`` is the same restatement executed the way the stock reference
runs on a GPU -- (c) the "reference on the GPU" arm of bench.py: cuDNN / cuBLAS / ATen kernels, fp16 autocast around fnet,
cnet and the update block (raft.py:92,100,117: ``with autocast():`` has no ``enabled=`` switch), batch 1 per frame, and
oracle/pose_torch.solve = torch.optim.LBFGS over the (pure-torch) lietorch stand-in with a host sync per iteration.
"""
import contextlib
import numpy as np
import torch
import torch.nn.functional as F

from . import pose_np, se3_np


# ------------------------------------------------------------------------------------------------------
# network pieces (weights: the checkpoint's state_dict, keys as shipped)
# ------------------------------------------------------------------------------------------------------
def _conv(x, sd, name, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride, padding)


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        training=False, eps=1e-5)


def basic_encoder(x, sd, pre, batch_norm):
    norm = (lambda t, n: _bn(t, sd, n)) if batch_norm else (lambda t, n: F.instance_norm(t, eps=1e-5))
    x = F.relu(norm(_conv(x, sd, pre + "conv1", 2, 3), pre + "norm1"))
    for layer, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        for blk in (0, 1):
            p = f"{pre}{layer}.{blk}."
            s = stride if blk == 0 else 1
            y = F.relu(norm(_conv(x, sd, p + "conv1", s, 1), p + "norm1"))
            y = F.relu(norm(_conv(y, sd, p + "conv2", 1, 1), p + "norm2"))
            if s != 1:
                x = norm(_conv(x, sd, p + "downsample.0", s, 0), p + "norm3")
            x = F.relu(x + y)
    return _conv(x, sd, pre + "conv2")


def update_block(net, inp, corr, flow, sd, pre="flow.update_block."):
    e = pre + "encoder."
    cor = F.relu(_conv(corr, sd, e + "convc1"))
    cor = F.relu(_conv(cor, sd, e + "convc2", 1, 1))
    flo = F.relu(_conv(flow, sd, e + "convf1", 1, 3))
    flo = F.relu(_conv(flo, sd, e + "convf2", 1, 1))
    out = F.relu(_conv(torch.cat([cor, flo], 1), sd, e + "conv", 1, 1))
    x = torch.cat([inp, out, flow], 1)
    g = pre + "gru."
    for half, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([net, x], 1)
        z = torch.sigmoid(_conv(hx, sd, g + "convz" + half, 1, pad))
        r = torch.sigmoid(_conv(hx, sd, g + "convr" + half, 1, pad))
        q = torch.tanh(_conv(torch.cat([r * net, x], 1), sd, g + "convq" + half, 1, pad))
        net = (1 - z) * net + z * q
    delta = _conv(F.relu(_conv(net, sd, pre + "flow_head.conv1", 1, 1)), sd, pre + "flow_head.conv2", 1, 1)
    mask = 0.25 * _conv(F.relu(_conv(net, sd, pre + "mask.0", 1, 1)), sd, pre + "mask.2")
    return net, mask, delta


def corr_pyramid(f1, f2, levels=4):
    B, C, h, w = f1.shape
    corr = torch.matmul(f1.view(B, C, h * w).transpose(1, 2), f2.view(B, C, h * w))
    corr = (corr.view(B, h, w, 1, h, w) / torch.sqrt(torch.tensor(C).float())).reshape(B * h * w, 1, h, w)
    pyr = [corr]
    for _ in range(levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def corr_lookup(pyr, coords, r=4):
    B, _, h, w = coords.shape
    coords = coords.permute(0, 2, 3, 1)
    out = []
    d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
    for i, corr in enumerate(pyr):
        c = coords.reshape(B * h * w, 1, 1, 2) / 2 ** i + delta
        H, W = corr.shape[-2:]
        gx = 2 * c[..., 0:1] / (W - 1) - 1
        gy = 2 * c[..., 1:2] / (H - 1) - 1
        s = F.grid_sample(corr, torch.cat([gx, gy], -1), align_corners=True)
        out.append(s.view(B, h, w, -1))
    return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous().float()


def convex_upsample(flow, mask):
    N, _, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 2, 8 * H, 8 * W)


def raft_forward(sd, image1, image2, iters=12, autocast=False):
    """-> (flow_up (B,2,H,W), net, inp) like RAFT.forward(...)[0][-1], [1], [2].  ``autocast``: the fp16 autocast regions of
    the reference's CUDA run (raft.py:92,100,117); on CPU the reference's autocast is a no-op."""
    amp = (lambda: torch.autocast("cuda", dtype=torch.float16)) if autocast else contextlib.nullcontext
    dev = image1.device
    image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
    image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
    B = image1.shape[0]
    with amp():
        fm = basic_encoder(torch.cat([image1, image2], 0), sd, "flow.fnet.", False)
    pyr = corr_pyramid(fm[:B].float(), fm[B:].float())
    with amp():
        cnet = basic_encoder(image1, sd, "flow.cnet.", True)
        net, inp = torch.split(cnet, [128, 128], dim=1)
        net, inp = torch.tanh(net), torch.relu(inp)
    h, w = image1.shape[-2] // 8, image1.shape[-1] // 8
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    coords0 = torch.stack([xs, ys], 0).float()[None].repeat(B, 1, 1, 1)
    coords1 = coords0.clone()
    flow_up = None
    for _ in range(iters):
        corr = corr_lookup(pyr, coords1)
        with amp():
            net, up_mask, delta = update_block(net, inp, corr, coords1 - coords0, sd)
        coords1 = coords1 + delta
        flow_up = convex_upsample(coords1 - coords0, up_mask)      # the reference up-samples every iteration
    return flow_up, net, inp


def tiny_unet(x, sd, pre, out_size):
    feats = []
    for i in range(3):
        p = f"{pre}encoder.enc_blocks.{i}."
        x = _conv(F.relu(_bn(_conv(x, sd, p + "conv1"), sd, p + "norm")), sd, p + "conv2")
        feats.append(x)
        x = F.max_pool2d(x, 2)
    x = feats[2]
    for i in range(2):
        x = F.conv_transpose2d(x, sd[f"{pre}decoder.upconvs.{i}.weight"], sd[f"{pre}decoder.upconvs.{i}.bias"], stride=2)
        e = feats[1 - i]
        dh, dw = (e.shape[-2] - x.shape[-2]) // 2, (e.shape[-1] - x.shape[-1]) // 2
        e = e[..., dh:e.shape[-2] - dh, dw:e.shape[-1] - dw]
        p = f"{pre}decoder.dec_blocks.{i}."
        x = _conv(_bn(F.relu(_conv(torch.cat([x, e], 1), sd, p + "conv1")), sd, p + "norm"), sd, p + "conv2")
    return F.interpolate(_conv(x, sd, pre + "head"), out_size, mode="bilinear")


# ------------------------------------------------------------------------------------------------------
# PoseNet.infer / PoseEstimator restated
# ------------------------------------------------------------------------------------------------------
def _remap(x, flow, mode="bilinear"):
    n, _, h, w = flow.shape
    rows, cols = torch.meshgrid(torch.arange(h, device=flow.device), torch.arange(w, device=flow.device), indexing="ij")
    g = torch.empty_like(flow)
    g[:, 1] = 2 * (flow[:, 1] + rows) / (h - 1) - 1
    g[:, 0] = 2 * (flow[:, 0] + cols) / (w - 1) - 1
    return F.grid_sample(x, g.permute(0, 2, 3, 1), align_corners=True, mode=mode)


def _proj(depth, K):
    n, _, H, W = depth.shape
    dev = depth.device
    xs = torch.linspace(0, W - 1, W, device=dev).repeat(1, H, 1) + 0.5
    ys = torch.linspace(0, H - 1, H, device=dev).repeat(1, W, 1).transpose(1, 2) + 0.5
    ic = torch.vstack([xs.flatten(), ys.flatten(), torch.ones(H * W, device=dev)])
    return (depth.view(n, 1, -1) * (torch.linalg.inv(K) @ ic.view(1, 3, -1))).view(n, 3, H, W)


class RefTracker:
    """CPU port of PoseEstimator(frame2frame=True) + PoseNet.infer; ``step`` mirrors ``forward``."""

    def __init__(self, state_dict, K, bf, lbgfs_iters=20, conf_weighing=True, depth_clip=250.0, device="cpu", autocast=False,
                 solver="numpy"):
        self.device = torch.device(device)
        self.autocast = bool(autocast)
        self.solver = solver            # "numpy": oracle/pose_np (fp64 restatement); "torch": torch.optim.LBFGS like the reference
        self.sd = {k.replace("module.", ""): (v.float() if v.is_floating_point() else v).to(self.device) for k, v in state_dict.items()}
        self.K = torch.as_tensor(K).float()[None].to(self.device)
        self.scale = torch.tensor(1 / depth_clip, device=self.device)
        self.baseline = torch.tensor(bf).unsqueeze(0).float().to(self.device)
        self.iters = lbgfs_iters
        self.use_weights = conf_weighing
        self.last_pose = se3_np.identity().astype(np.float32)
        self.frame = None
        self.timing = {"raft": 0.0, "heads": 0.0, "solve": 0.0}
        self.last = {}

    def _flow2depth(self, limg, rimg):
        flow = raft_forward(self.sd, limg, rimg, autocast=self.autocast)[0].float()
        depth = (self.baseline * self.scale)[:, None, None] / -flow[:, 0]
        valid = (depth > 0) & (depth <= 1.0)
        depth[~valid] = 1.0
        return depth.unsqueeze(1), flow, valid.unsqueeze(1)

    def _infer(self, last, cur):
        import time
        sd = self.sd
        t0 = time.perf_counter()
        flow, net, inp = raft_forward(sd, torch.cat([last["img"], cur["img"]]), torch.cat([cur["img"], cur["rimg"]]),
                                      autocast=self.autocast)
        t1 = time.perf_counter()
        flow = flow.float()
        time_flow, sflow2 = flow[0:1], flow[1:2]
        gru, ctx = net[0:1].float(), inp[0:1].float()
        depth2 = (self.baseline * self.scale)[:, None, None] / -sflow2[:, 0]
        valid = (depth2 > 0) & (depth2 <= 1.0)
        depth2[~valid] = 1.0
        depth2 = depth2.unsqueeze(1)
        cur["mask"] &= valid.unsqueeze(1)
        depth1 = last["depth"] * self.scale
        pcl1, pcl2 = _proj(depth1, self.K), _proj(depth2, self.K)
        pcl2w = _remap(pcl2, time_flow)
        img2w = _remap(cur["img"], time_flow)
        sflow2w = _remap(sflow2, time_flow)
        m2w = _remap(cur["mask"].float(), time_flow, mode="nearest")
        mask2w = (m2w > 0).any(dim=1).unsqueeze(1) & m2w.to(bool)
        H, W = time_flow.shape[-2:]
        if self.use_weights:
            inp1 = F.interpolate(torch.cat((last["flow"], last["img"], pcl1), 1), scale_factor=0.125, mode="bilinear")
            inp2 = F.interpolate(torch.cat((sflow2w, img2w, pcl2w), 1), scale_factor=0.125, mode="bilinear")
            conf1 = torch.sigmoid(tiny_unet(torch.cat((inp1, gru, ctx), 1), sd, "weight_head_2d.0.", (H, W)))
            conf2 = torch.sigmoid(tiny_unet(torch.cat((inp1, inp2, gru, ctx), 1), sd, "weight_head_3d.0.", (H, W)))
        else:
            conf1 = torch.ones((1, 1, H, W), device=self.device)
            conf2 = torch.ones((1, 1, H, W), device=self.device)
        t2 = time.perf_counter()
        if self.solver == "torch":
            from . import pose_torch
            Xt, lgt, n_evals = pose_torch.solve(time_flow, pcl1, pcl2w, conf1, conf2, last["mask"], mask2w, self.K,
                                                sd["loss_weight"][None], lbgfs_iters=self.iters)
            X, lg = Xt[0].float().cpu().numpy().astype(np.float64), lgt[0].float().cpu().numpy().astype(np.float64)
        else:
            X, lg, n_evals = pose_np.lbfgs_solve(time_flow[0].cpu().numpy(), pcl1[0].cpu().numpy(), pcl2w[0].cpu().numpy(),
                                                 conf1[0, 0].cpu().numpy(), conf2[0, 0].cpu().numpy(), last["mask"][0, 0].cpu().numpy(),
                                                 mask2w[0, 0].cpu().numpy(), self.K[0].cpu().numpy(), sd["loss_weight"].cpu().numpy(),
                                                 max_iter=self.iters)
        t3 = time.perf_counter()
        self.timing["raft"] += t1 - t0
        self.timing["heads"] += t2 - t1
        self.timing["solve"] += t3 - t2
        cur["depth"] = depth2 / self.scale
        cur["flow"] = sflow2
        self.last = dict(time_flow=time_flow, stereo_flow2=sflow2, conf1=conf1, conf2=conf2, mask2w=mask2w,
                         mask2_valid=cur["mask"].clone(), n_evals=n_evals, rel=X.astype(np.float32), log=lg.astype(np.float32))
        return X.astype(np.float32), lg.astype(np.float32)

    def step(self, limg, rimg, mask):
        """limg, rimg (1,3,H,W) float tensors 0..255, mask (1,1,H,W) bool -> absolute pose (7,) float32 in mm."""
        with torch.no_grad():
            last = self.frame
            limg, rimg, mask = limg.to(self.device), rimg.to(self.device), mask.to(self.device)
            cur = {"img": limg.contiguous(), "rimg": rimg.contiguous(), "mask": mask.bool().clone()}
            if last is None:
                depth, sflow, _ = self._flow2depth(limg, rimg)
                cur["depth"], cur["flow"] = depth / self.scale, sflow
                rel, lg = se3_np.identity().astype(np.float32), np.zeros(6, np.float32)
            else:
                rel, lg = self._infer(last, cur)
            self.frame = cur
        if np.isnan(rel).any() or (np.abs(lg) > 0.1).any():
            rel = se3_np.identity().astype(np.float32)
        inv_scale = np.float32(1.0) / np.float32(self.scale.item())
        rel = se3_np.scale(rel.astype(np.float64), float(inv_scale))
        self.last_pose = se3_np.mul(self.last_pose.astype(np.float64), se3_np.inv(rel)).astype(np.float32)
        return self.last_pose
