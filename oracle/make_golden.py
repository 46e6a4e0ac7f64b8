"""ORACLE / TEST INFRASTRUCTURE ONLY -- generates golden vectors by RUNNING THE UNMODIFIED REFERENCE.

The reference (/root/reference, read-only) is imported as-is, with oracle/lietorch standing in for
the missing third-party ``lietorch``.  Runs only in the build container (the GPU box has no
/root/reference); its outputs travel:

  tests/golden/*.npz            small, committed
  oracle/_ref/golden_full.npz   full 640x512 stage dump, git-ignored, shipped by gpurun
  oracle/_ref/trained/*.pth     the reference's shipped checkpoints (data, not source), git-ignored

Reference entry points exercised:
  PoseEstimator.forward                 core/pose/pose_estimator.py:50-96   (3 frames -> 2 pairs)
  PoseNet.infer                         core/pose/pose_net.py:60-85
  RAFT.forward / CorrBlock              core/RAFT/core/raft.py:79-137, core/RAFT/core/corr.py:12-60
  remap_from_flow(_nearest)             core/interpol/flow_utils.py:4-26
  DPoseSE3Head.objective / solve        core/pose/pose_head.py:53-79

usage: python oracle/make_golden.py [--full]
"""
import argparse
import os
import shutil
import sys
import importlib.util

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("RPE_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import torch  # noqa: E402
import yaml  # noqa: E402


def _load_synth():
    spec = importlib.util.spec_from_file_location(
        "rpe_synth", os.path.join(ROOT, "robust-pose-estimator_b200", "dataset", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


from detrand import det_uniform  # noqa: E402


def pack(mask):
    return np.packbits(np.asarray(mask).astype(bool).reshape(-1))


class Recorder:
    def __init__(self):
        self.lbfgs = []       # per objective evaluation: (pose7, grad6 before clipping)
        self.corr = []
        self.stage = {}


class FixedFrames:
    """Minimal stand-in for SyntheticStereoSequence over given uint8 frames (config 1: the reference's own fixtures)."""

    def __init__(self, L, R, M, K, bf):
        self.L, self.R, self.M = L, R, M
        self.calib = {"intrinsics": {"left": np.asarray(K, np.float64)}, "bf": float(bf)}
        self.rel_xi = np.zeros((len(L) - 1, 6))

    def __getitem__(self, i):
        return self.L[i].astype(np.float32), self.R[i].astype(np.float32), self.M[i].copy(), i


def run_sequence(size, seed, n_frames, ckpt, rec_corr=False, holes=2, conf=True, loss_weight_mask=None, seq=None, quiet=False):
    """Run the reference tracker over a synthetic sequence and record every stage of the last pair."""
    from core.pose.pose_estimator import PoseEstimator
    import core.pose.pose_net as pose_net_mod
    import core.RAFT.core.raft as raft_mod
    from core.RAFT.core.corr import CorrBlock
    from lietorch import SE3

    if seq is None:
        synth = _load_synth()
        seq = synth.SyntheticStereoSequence(n_frames, size, seed=seed, holes=holes)
    with open(os.path.join(REF, "configuration", "infer_f2f.yaml")) as f:
        config = yaml.load(f, Loader=yaml.SafeLoader)
    config["slam"]["conf_weighing"] = conf
    config["img_size"] = list(size)
    est = PoseEstimator(config["slam"], torch.tensor(seq.calib["intrinsics"]["left"]), baseline=seq.calib["bf"],
                        checkpoint=ckpt, img_shape=config["img_size"], init_pose=SE3.Identity(1))
    model = est.model
    if loss_weight_mask is not None:
        # the shipped code has no switch to drop a residual term; "3D-only" = the reference objective run with
        # loss_weight[1] (the 2D term, pose_head.py:53-58) zeroed
        with torch.no_grad():
            model.loss_weight.mul_(torch.tensor(loss_weight_mask, dtype=model.loss_weight.dtype))
    rec = Recorder()

    # ---- hooks (wrap, never modify, the reference callables) ------------------------------
    class RecCorrBlock(CorrBlock):
        def __init__(self, fmap1, fmap2, **kw):
            super().__init__(fmap1, fmap2, **kw)
            self._rec = {"fmap1": fmap1.clone(), "fmap2": fmap2.clone(), "lookups": []}
            rec.corr.append(self)

        def __call__(self, coords):
            out = super().__call__(coords)
            if rec_corr:
                self._rec["lookups"].append((coords.clone(), out.clone()))
            return out

    raft_mod.CorrBlock = RecCorrBlock

    orig_clip = torch.nn.utils.clip_grad_norm_

    def clip_hook(params, max_norm, *a, **k):
        p = params
        rec.lbfgs.append((p.group.data.detach().clone().reshape(-1), p.grad.detach().clone().reshape(-1)))
        return orig_clip(params, max_norm, *a, **k)

    torch.nn.utils.clip_grad_norm_ = clip_hook

    orig_gwm = model.get_weight_maps

    def gwm_hook(pcl1, pcl2, image1l, image2l, mask2, time_flow, stereo_flow1, stereo_flow2, gru, ctx):
        rec.stage.update(pcl1=pcl1.clone(), pcl2=pcl2.clone(), mask2_valid=mask2.clone(),
                         time_flow=time_flow.clone(), stereo_flow1=stereo_flow1.clone(),
                         stereo_flow2=stereo_flow2.clone(), gru=gru.clone(), ctx=ctx.clone())
        out = orig_gwm(pcl1, pcl2, image1l, image2l, mask2, time_flow, stereo_flow1, stereo_flow2, gru, ctx)
        conf1, conf2, pcl2w, mask2w = out
        rec.stage.update(conf1=conf1.clone(), conf2=conf2.clone(), pcl2w=pcl2w.clone(), mask2w=mask2w.clone())
        return out

    model.get_weight_maps = gwm_hook

    orig_remap = pose_net_mod.remap_from_flow

    def remap_hook(x, flow):
        y, valid = orig_remap(x, flow)
        rec.stage.setdefault("remaps", []).append(y.clone())
        return y, valid

    pose_net_mod.remap_from_flow = remap_hook

    orig_solve = model.pose_head.problem.solve

    def solve_hook(*xs):
        rec.stage["xs"] = [x.detach().clone() for x in xs]
        rec.lbfgs.clear()
        return orig_solve(*xs)

    model.pose_head.problem.solve = solve_hook

    out = {"K": seq.calib["intrinsics"]["left"], "bf": np.float64(seq.calib["bf"]), "size": np.array(size),
           "seed": np.int64(seed), "gt_rel_xi": seq.rel_xi}
    frames = []
    poses = []
    rel = []
    rel_pose = []                       # PoseNet.infer's own return value per pair (normalised units, before the guard)
    orig_infer = model.infer

    def infer_hook(*a, **k):
        r = orig_infer(*a, **k)
        rel_pose.append((r[0] if isinstance(r, tuple) else r).vec().detach().clone().numpy().reshape(7))
        return r

    model.infer = infer_hook
    try:
        with torch.no_grad():
            for i in range(n_frames):
                limg, rimg, mask, _ = seq[i]
                frames.append((limg.astype(np.uint8), rimg.astype(np.uint8), mask.copy()))
                rec.corr.clear()
                pose, _, flow, weights = est(torch.from_numpy(limg)[None], torch.from_numpy(rimg)[None],
                                             torch.from_numpy(mask)[None])
                poses.append(pose.vec().detach().clone().numpy().reshape(7))
                if i > 0:
                    xs = rec.stage["xs"]
                    rel.append(dict(evals=[(p.numpy().copy(), g.numpy().copy()) for p, g in rec.lbfgs]))
                if not quiet or i % 8 == 0:
                    print(f"  frame {i}: pose {poses[-1]}  evals {len(rec.lbfgs)}", flush=True)
    finally:
        raft_mod.CorrBlock = CorrBlock
        torch.nn.utils.clip_grad_norm_ = orig_clip
        pose_net_mod.remap_from_flow = orig_remap

    out["imgs_l"] = np.stack([f[0] for f in frames])
    out["imgs_r"] = np.stack([f[1] for f in frames])
    out["masks_in"] = np.stack([pack(f[2]) for f in frames])
    out["traj"] = np.stack(poses)
    out["rel_pose"] = np.stack(rel_pose) if rel_pose else np.zeros((0, 7), np.float32)
    for k, r in enumerate(rel):
        out[f"pair{k}_eval_pose"] = np.stack([e[0] for e in r["evals"]])
        out[f"pair{k}_eval_grad"] = np.stack([e[1] for e in r["evals"]])
    return out, rec, est


def stage_dict(rec, est, full):
    """Stage tensors of the LAST pair.  `full` keeps every tensor; otherwise floats are sub-sampled."""
    st = rec.stage
    xs = st["xs"]
    flow, pcl1, pcl2w, w1, w2, m1, m2w, K, lw = xs
    head = est.model.pose_head.problem
    d = {}
    f32 = lambda t: t.detach().numpy().astype(np.float32)
    d["time_flow"] = f32(st["time_flow"][0])
    d["stereo_flow1"] = f32(st["stereo_flow1"][0]) if full else None
    d["stereo_flow2"] = f32(st["stereo_flow2"][0])
    d["mask1"] = pack(m1)
    d["mask2_valid"] = pack(st["mask2_valid"])          # mask2 &= stereo-valid   (pose_net.py:77)
    d["mask2w"] = pack(m2w)
    d["loss_weight"] = f32(lw[0])
    d["depth1_norm"] = f32(est.last_frame.depth[0, 0] * est.scale) if est.last_frame is not None else None
    keep = (lambda a: a) if full else (lambda a: a.reshape(a.shape[0], -1)[:, ::7].copy())
    d["pcl1"] = keep(f32(st["pcl1"][0]))
    d["pcl2"] = keep(f32(st["pcl2"][0]))
    d["pcl2w"] = keep(f32(pcl2w[0]))
    d["img2w"] = keep(f32(st["remaps"][-2][0]))
    d["sflow2w"] = keep(f32(st["remaps"][-1][0]))
    d["conf1"] = f32(w1[0]) if full else f32(w1[0]).astype(np.float16)
    d["conf2"] = f32(w2[0]) if full else f32(w2[0]).astype(np.float16)
    d["gru"] = f32(st["gru"][0]) if full else None
    d["ctx"] = f32(st["ctx"][0]) if full else None
    d = {k: v for k, v in d.items() if v is not None}
    # objective + autograd gradient of the reference at a few poses (fp64 like solve())
    from lietorch import SE3, LieGroupParameter
    xs64 = [x.double() if x.dtype == torch.float32 else x for x in xs]
    probe = [np.zeros(6), np.array([0.01, -0.02, 0.015, 0.004, -0.003, 0.002]),
             np.array([-0.03, 0.01, 0.02, -0.01, 0.008, 0.012])]
    vals = []
    for xi in probe:
        with torch.enable_grad():
            G = SE3.exp(torch.tensor(xi, dtype=torch.float64).view(1, 1, 6))
            y = LieGroupParameter(G)
            loss = head.objective(*xs64, y=(y,)).sum()
            loss.backward()
            l3 = head.depth_objective(xs64[1], xs64[2], xs64[4], xs64[5], xs64[6], y)
            l2 = head.reprojection_objective(xs64[0], xs64[1], xs64[3], xs64[5], xs64[7], y)
        vals.append(np.concatenate((G.data.numpy().reshape(7), [loss.item(), l2.item(), l3.item()],
                                    y.grad.numpy().reshape(6))))
    d["objective_probe"] = np.stack(vals)                # [pose7 | f, L2d, L3d | grad6]
    return d


def small_stage_goldens():
    """Reference operators on platform-exact pseudo-random tensors (no network, no checkpoint)."""
    from core.RAFT.core.corr import CorrBlock
    from core.interpol.flow_utils import remap_from_flow, remap_from_flow_nearest
    from core.geometry.pinhole_transforms import create_img_coords_t
    import torch.nn.functional as F
    g = {}
    # --- CorrBlock (corr.py:12-60): B=2, C=32, 16x16 grid (levels 16x16, 8x8, 4x4, 2x2)
    B, C, h, w = 2, 32, 16, 16
    f1 = torch.from_numpy(det_uniform((B, C, h, w), 11))
    f2 = torch.from_numpy(det_uniform((B, C, h, w), 12))
    cb = CorrBlock(f1, f2, num_levels=4, radius=4)
    for l, lvl in enumerate(cb.corr_pyramid):
        lv = lvl.numpy().reshape(B, h * w, *lvl.shape[-2:])
        g[f"corr_l{l}"] = lv[:1] if l == 0 else lv          # level 0 of batch 0 only (size)
    coords = torch.stack(torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")[::-1], 0).float()[None]
    coords = coords.repeat(B, 1, 1, 1) + torch.from_numpy(det_uniform((B, 2, h, w), 13, -9.0, 9.0))
    coords[0, :, 0, 0] = torch.tensor([3.0, 5.0])           # exactly-integer coordinates
    coords[0, :, 0, 1] = torch.tensor([-7.5, 40.0])         # far out of range
    g["corr_coords"] = coords.numpy()
    g["corr_lookup"] = cb(coords).numpy()
    # --- warps (flow_utils.py:4-26) on a 40x56 image, flow up to +-12 px, with exact .5 cases
    H, W = 40, 56
    x8 = torch.from_numpy(det_uniform((1, 8, H, W), 21, -2.0, 2.0))
    flow = torch.from_numpy(det_uniform((1, 2, H, W), 22, -12.0, 12.0))
    flow[0, :, 1, :8] = torch.tensor([0.5, 1.5, 2.5, -0.5, 3.5, 4.5, 0.0, -1.5])[None]
    flow[0, :, 2, :4] = torch.tensor([[60.0, -60.0, 0.25, 0.75], [0.0, 0.0, 45.0, -45.0]])
    m = torch.from_numpy(det_uniform((1, 1, H, W), 23, 0.0, 1.0) > 0.2)
    g["warp_flow"] = flow.numpy()
    g["warp_bilinear"] = remap_from_flow(x8, flow)[0].numpy()
    wm, valid = remap_from_flow_nearest(m, flow)
    g["warp_mask_in"] = pack(m)
    g["warp_mask_out"] = pack(valid & wm.to(bool))           # pose_net.py:107-108
    # --- 1/8 bilinear down-sampling (pose_net.py:110-113)
    g["down8"] = F.interpolate(x8, scale_factor=0.125, mode="bilinear").numpy()
    # --- depth from stereo flow + back-projection (pose_net.py:73-79,121-125)
    K = torch.tensor([[[35.0, 0.0, 28.0], [0.0, 35.0, 20.0], [0.0, 0.0, 1.0]]])
    sflow = torch.from_numpy(det_uniform((1, 2, H, W), 24, -30.0, 4.0))
    sflow[0, 0, 0, :3] = torch.tensor([0.0, -8.8, -8.8000001])
    bfs = torch.tensor([8.8])
    depth = bfs[:, None, None] / -sflow[:, 0]
    valid = (depth > 0) & (depth <= 1.0)
    depth[~valid] = 1.0
    ic = create_img_coords_t(y=H, x=W)
    pcl = (depth.view(1, 1, -1) * (torch.linalg.inv(K) @ ic.view(1, 3, -1))).view(1, 3, H, W)
    g["geom_K"] = K.numpy()
    g["geom_sflow"] = sflow.numpy()
    g["geom_depth"] = depth.numpy()
    g["geom_valid"] = pack(valid)
    g["geom_pcl"] = pcl.numpy()
    return g


def posehead_unit_golden():
    """The reference's own pose-head recipe (tests/unit_test_pose_head.py:13-50) on exact inputs, n=2."""
    from lietorch import SE3
    from core.geometry.pinhole_transforms import transform, project, reproject, create_img_coords_t
    from core.pose.pose_head import DPoseSE3Head
    n, R = 2, 128
    kmat = torch.diag(torch.tensor([107.0, 107, 1]))
    kmat[0, -1] = R // 2
    kmat[1, -1] = R // 2
    kmat = kmat.repeat((n, 1, 1))
    depth = 100 * torch.clamp(torch.from_numpy(det_uniform((n, 1, R, R), 31, 0.0, 1.0)), 0.01, 1)
    ic = create_img_coords_t(R, R)
    pcl = reproject(depth, kmat, ic)[:, :3].view(n, 3, R, R)
    xi = torch.from_numpy(det_uniform((n, 1, 6), 32, -0.02, 0.02))
    poses = SE3.exp(xi)
    flow_off = project(pcl.view(n, 3, -1), kmat, poses)[:, :2].reshape(n, 2, R, R)
    valid = ((flow_off[:, 0] >= 0) & (flow_off[:, 0] < R) & (flow_off[:, 1] >= 0) & (flow_off[:, 1] < R)).unsqueeze(1)
    flow = flow_off - ic[:2].reshape(1, 2, R, R)
    pcl_t = transform(pcl.view(n, 3, -1), poses).view(n, 3, R, R)
    ones = torch.ones((n, 1, R, R))
    masks = torch.ones((n, 1, R, R), dtype=torch.bool)
    g = {"xi_gt": xi.numpy().reshape(n, 6), "depth": depth.numpy(), "K": kmat[0].numpy(),
         "flow": flow.numpy(), "pcl": pcl.numpy(), "pcl_t": pcl_t.numpy(), "valid": pack(valid)}
    vec, log, nev = [], [], []
    for i in range(n):                                    # per sample: batch>1 couples L-BFGS
        head = DPoseSE3Head(ic, lbgfs_iters=100)
        evals = []
        orig_clip = torch.nn.utils.clip_grad_norm_
        torch.nn.utils.clip_grad_norm_ = lambda p, m, *a, **k: (evals.append(1), orig_clip(p, m, *a, **k))[1]
        try:
            xs = (flow[i:i + 1], pcl[i:i + 1], pcl_t[i:i + 1], ones[i:i + 1], ones[i:i + 1], valid[i:i + 1],
                  masks[i:i + 1], kmat[i:i + 1], torch.tensor([[0.001, 1.0]]))
            y = head.solve(*xs)[0]
        finally:
            torch.nn.utils.clip_grad_norm_ = orig_clip
        vec.append(y.group.vec().detach().numpy().reshape(7))
        log.append(y.log().detach().numpy().reshape(6))
        nev.append(len(evals))
    g["sol_vec"] = np.stack(vec)
    g["sol_log"] = np.stack(log)
    g["n_evals"] = np.array(nev)
    return g


def posehead_real_golden(rec, est, step=3):
    """Pose-head inputs of a real network run, decimated by `step`, fed to the reference's DPoseSE3Head
    as a smaller problem (pins objective, autograd gradient, L-BFGS trajectory on realistic
    residuals, masks and confidence weights)."""
    from lietorch import SE3, LieGroupParameter
    from core.geometry.pinhole_transforms import create_img_coords_t
    from core.pose.pose_head import DPoseSE3Head
    flow, pcl1, pcl2w, w1, w2, m1, m2w, K, lw = [x.clone() for x in rec.stage["xs"]]
    sub = lambda t: t[..., ::step, ::step].contiguous()
    flow, pcl1, pcl2w, w1, w2, m1, m2w = map(sub, (flow, pcl1, pcl2w, w1, w2, m1, m2w))
    flow = flow / step
    K = K.clone()
    K[:, :2] = K[:, :2] / step
    h, w = flow.shape[-2:]
    head = DPoseSE3Head(create_img_coords_t(h, w), lbgfs_iters=20)
    xs = (flow, pcl1, pcl2w, w1, w2, m1, m2w, K, lw)
    evals = []
    orig_clip = torch.nn.utils.clip_grad_norm_

    def hook(p, m, *a, **k):
        evals.append((p.group.data.detach().clone().reshape(-1).numpy(), p.grad.detach().clone().reshape(-1).numpy()))
        return orig_clip(p, m, *a, **k)

    torch.nn.utils.clip_grad_norm_ = hook
    try:
        y = head.solve(*xs)[0]
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    g = {"flow": flow[0].numpy(), "pcl1": pcl1[0].numpy(), "pcl2w": pcl2w[0].numpy(), "w1": w1[0, 0].numpy(),
         "w2": w2[0, 0].numpy(), "m1": pack(m1), "m2w": pack(m2w), "K": K[0].numpy(), "lw": lw[0].numpy(),
         "shape": np.array([h, w]),
         "eval_pose": np.stack([e[0] for e in evals]), "eval_grad": np.stack([e[1] for e in evals]),
         "sol_vec": y.group.vec().detach().numpy().reshape(7), "sol_log": y.log().detach().numpy().reshape(6)}
    xs64 = [x.double() if x.dtype == torch.float32 else x for x in xs]
    vals = []
    for xi in (np.zeros(6), np.array([0.01, -0.02, 0.015, 0.004, -0.003, 0.002]), np.array([-0.03, 0.01, 0.02, -0.01, 0.008, 0.012])):
        with torch.enable_grad():
            G = SE3.exp(torch.tensor(xi, dtype=torch.float64).view(1, 1, 6))
            yy = LieGroupParameter(G)
            loss = head.objective(*xs64, y=(yy,)).sum()
            loss.backward()
            l3 = head.depth_objective(xs64[1], xs64[2], xs64[4], xs64[5], xs64[6], yy)
            l2 = head.reprojection_objective(xs64[0], xs64[1], xs64[3], xs64[5], xs64[7], yy)
        vals.append(np.concatenate((G.data.numpy().reshape(7), [loss.item(), l2.item(), l3.item()], yy.grad.numpy().reshape(6))))
    g["objective_probe"] = np.stack(vals)
    return g


def mask_spec_golden():
    """The reference's own mask_specularities (dataset/stereo_dataset.py:12-16, cv2.erode) on deterministic frames with
    saturated blobs touching the borders; sizes that are not multiples of the CUDA tile."""
    from dataset.stereo_dataset import mask_specularities                  # the unmodified reference function
    from detrand import det_uniform                                      # oracle/ is on sys.path (HERE)
    g = {}
    for k, (H, W) in enumerate(((70, 90), (33, 129), (64, 64))):
        img = det_uniform((H, W, 3), 900 + k, 0.0, 255.0).astype(np.uint8)
        yy, xx = np.mgrid[0:H, 0:W]
        for (cy, cx, rad) in ((0, 0, 9), (H - 1, W // 2, 7), (H // 2, W - 1, 12), (H // 3, W // 3, 5), (H // 2, W // 2, 1)):
            img[(yy - cy) ** 2 + (xx - cx) ** 2 <= rad * rad] = 250 + (k % 3)        # saturated highlights (sum >= 750)
        img[5, 7] = (245, 245, 244)                                                  # sum 734: just below the threshold
        img[9, 11] = (245, 245, 245)                                                 # sum 735: just above
        mask = det_uniform((H, W), 950 + k, 0.0, 1.0) > 0.002                        # a few isolated invalid pixels
        mask[H - 4:, :6] = False
        g[f"img{k}"] = img
        g[f"mask{k}"] = np.packbits(mask)
        g[f"shape{k}"] = np.array([H, W])
        g[f"out{k}"] = np.packbits(mask_specularities(img, mask.copy()).astype(bool))
        g[f"out_nomask{k}"] = np.packbits(mask_specularities(img).astype(bool))
    return g


def tartan_frames(size=(640, 512), mm_per_unit=60.0, bf=2200.0):
    """BASELINE config 1: stereo frames from the reference's OWN fixtures tests/test_data/tartan_air
    (000000/000001 _left.png + _left_depth.npy; real rendered texture).  The left image and its depth go through the
    reference's ResizeStereo (dataset/transforms.py:20-39) to 640x512; the fixtures hold no right view, so it is synthesised
    from the depth: disparity = bf / depth_mm, forward-splatted into the right view with a z-buffer (nearest surface wins,
    holes take the farther neighbour), then the left image is sampled at x + disparity.  Images are quantised to uint8."""
    import cv2
    from dataset.transforms import ResizeStereo
    d = os.path.join(REF, "tests", "test_data", "tartan_air")
    tr = ResizeStereo(size)
    L, R = [], []
    for name in ("000000", "000001"):
        left = cv2.cvtColor(cv2.imread(os.path.join(d, name + "_left.png")), cv2.COLOR_BGR2RGB)
        depth = np.load(os.path.join(d, name + "_left_depth.npy")).astype(np.float32)
        lt, dt, _ = tr(torch.from_numpy(left).permute(2, 0, 1).float(), torch.from_numpy(depth)[None], None)
        left_u8 = np.clip(np.rint(lt.numpy()), 0, 255).astype(np.uint8)                # (3,H,W)
        disp = bf / (dt.numpy()[0].astype(np.float64) * mm_per_unit)                   # (H,W) pixels
        H, W = disp.shape
        cols = np.arange(W)[None, :].repeat(H, 0)
        xr = np.rint(cols - disp).astype(np.int64)
        ok = (xr >= 0) & (xr < W)
        order = np.argsort(disp, axis=None, kind="stable")                             # ascending: larger disparity written last
        order = order[ok.ravel()[order]]
        disp_r = np.full(H * W, np.nan)
        rows = (order // W)
        disp_r[rows * W + xr.ravel()[order]] = disp.ravel()[order]
        disp_r = disp_r.reshape(H, W)
        fl, fr = disp_r.copy(), disp_r.copy()
        for x in range(1, W):                                                          # nearest valid neighbour from each side
            m = np.isnan(fl[:, x])
            fl[m, x] = fl[m, x - 1]
        for x in range(W - 2, -1, -1):
            m = np.isnan(fr[:, x])
            fr[m, x] = fr[m, x + 1]
        disp_r = np.fmin(fl, fr)                                                       # the farther surface fills a hole
        disp_r[np.isnan(disp_r)] = np.nanmedian(disp)
        xs = np.clip(cols + disp_r, 0, W - 1)
        x0 = np.clip(np.floor(xs).astype(np.int64), 0, W - 2)
        a = (xs - x0)[None]
        rr = np.arange(H)[:, None]
        lf = left_u8.astype(np.float64)
        right = (1 - a) * lf[:, rr, x0] + a * lf[:, rr, x0 + 1]
        L.append(left_u8)
        R.append(np.clip(np.rint(right), 0, 255).astype(np.uint8))
    K = np.array([[320.0, 0, 320], [0, 320.0, 256], [0, 0, 1]])
    return np.stack(L), np.stack(R), K, bf


def metrics_golden():
    """The reference's own evaluate_ate_freiburg.eval (evaluation/evaluate_ate_freiburg.py:6-33 -> core/metrics/trajectory_metrics.py)
    on a deterministic ground-truth / estimate pair written with the reference's save_trajectory."""
    import tempfile
    from lietorch import SE3
    from core.utils.trajectory import save_trajectory
    from evaluation.evaluate_ate_freiburg import eval as ref_eval
    n = 40
    xi = det_uniform((n, 6), 77, -0.05, 0.05).astype(np.float64)
    xi[:, :3] *= 40.0                                                       # millimetres
    poses_gt, poses_pr = [], []
    T_gt = SE3.Identity(1).double()
    T_pr = SE3.exp(torch.tensor([[3.0, -2.0, 1.0, 0.02, -0.01, 0.03]], dtype=torch.float64))      # a rigid offset the alignment removes
    noise = det_uniform((n, 6), 78, -1.0, 1.0).astype(np.float64) * np.array([0.3, 0.3, 0.3, 2e-3, 2e-3, 2e-3])
    for k in range(n):
        step = SE3.exp(torch.tensor(xi[k:k + 1]))
        T_gt = T_gt * step
        T_pr = T_pr * SE3.exp(torch.tensor(xi[k:k + 1] + (noise[k:k + 1] if k % 7 else 0.0)))
        poses_gt.append(T_gt), poses_pr.append(T_pr)
    poses_pr[20] = poses_pr[19]                                              # one "failed" pair: the pose is repeated
    out = {}
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        save_trajectory([{"camera-pose": p.float(), "timestamp": k} for k, p in enumerate(poses_gt)], a)
        save_trajectory([{"camera-pose": p.float(), "timestamp": k} for k, p in enumerate(poses_pr)], b)
        fa, fb = os.path.join(a, "trajectory.freiburg"), os.path.join(b, "trajectory.freiburg")
        out["gt_file"] = np.frombuffer(open(fa, "rb").read(), dtype=np.uint8)
        out["pred_file"] = np.frombuffer(open(fb, "rb").read(), dtype=np.uint8)
        for name, kw in (("plain", {}), ("delta3", {"delta": 3}), ("offset", {"offset": -4}), ("ignore", {"ignore_failed_pos": True})):
            r = ref_eval(fa, fb, **kw)
            out[name + "_scalars"] = np.array([float(r[0]), float(r[1]), float(r[2])])
            out[name + "_trans_error"] = np.asarray(r[3], dtype=np.float64)
            out[name + "_rpe_trans"] = np.asarray(r[4], dtype=np.float64)
            out[name + "_rpe_rot"] = np.asarray(r[5], dtype=np.float64)
    return out


def resize_golden():
    """The reference's own ResizeStereo (dataset/transforms.py:20-39, torchvision resize + center_crop) on deterministic frames:
    up-scaling with crop, exact 2:1 down-scaling, a non-integer down-scaling, and the nearest-neighbour mask path."""
    from dataset.transforms import ResizeStereo
    g = {}
    for k, ((Hi, Wi), (W, H)) in enumerate((((120, 160), (160, 128)), ((256, 320), (160, 128)), ((135, 240), (160, 128)), ((90, 100), (64, 96)))):
        left = torch.from_numpy(det_uniform((3, Hi, Wi), 500 + k, 0.0, 255.0)).floor()           # integer-valued like decoded frames
        right = torch.from_numpy(det_uniform((3, Hi, Wi), 520 + k, 0.0, 255.0)).floor()
        mask = torch.from_numpy((det_uniform((1, Hi, Wi), 540 + k, 0.0, 1.0) > 0.3).astype(np.uint8))
        l, r, m = ResizeStereo((W, H))(left, right, mask)
        g[f"in_shape{k}"], g[f"size{k}"] = np.array([Hi, Wi]), np.array([W, H])
        g[f"left{k}"], g[f"right{k}"], g[f"mask{k}"] = l.numpy(), r.numpy(), m.numpy()
    return g


def filedata_golden():
    """The reference's own ``get_data`` (dataset/dataset_utils.py:10-35) on the on-disk fixtures of oracle/file_fixture.py: the
    StereoDataset frame folder with an .ini calibration, the StereoVideoDataset folder with a .yaml calibration, ground truth
    and time stamps, and the rectified calibration of a camcal.json."""
    import hashlib
    import tempfile
    import cv2
    import file_fixture as ff
    from dataset.dataset_utils import get_data
    from dataset.rectification import StereoRectifier
    g = {}

    def put_calib(prefix, calib):
        g[prefix + "K_left"], g[prefix + "K_right"] = np.asarray(calib["intrinsics"]["left"]), np.asarray(calib["intrinsics"]["right"])
        g[prefix + "extrinsics"], g[prefix + "bf"] = np.asarray(calib["extrinsics"]), np.float64(calib["bf"])
        g[prefix + "bf_orig"], g[prefix + "img_size"] = np.float64(calib["bf_orig"]), np.asarray(calib["img_size"])

    with tempfile.TemporaryDirectory() as tmp:
        folder = ff.write_frame_folder(os.path.join(tmp, "frames"))
        dataset, calib = get_data(folder, ff.IMG_SIZE)
        assert type(dataset).__name__ == "StereoDataset" and len(dataset) == ff.N_FRAMES
        put_calib("frames_", calib)
        items = [dataset[k] for k in range(len(dataset))]
        # images: every third row / column only (file size); masks in full, bit-packed
        g["frames_left"] = np.stack([it[0].numpy() for it in items])[..., 1::3, 2::3]
        g["frames_right"] = np.stack([it[1].numpy() for it in items])[..., 1::3, 2::3]
        g["frames_mask"] = pack(np.stack([it[2].numpy() for it in items]))
        g["frames_number"] = np.array([it[3] for it in items])
        for mode in ("conventional", "pseudo"):
            rect = StereoRectifier(ff.write_json_calibration(os.path.join(tmp, "json")), img_size_new=ff.IMG_SIZE, mode=mode)
            put_calib(f"json_{mode}_", rect.get_rectified_calib())

    vdir = os.path.join(ROOT, "tests", "golden", "file_video")
    ff.write_video_folder(vdir, encode=not os.path.isfile(os.path.join(vdir, "video.mp4")))
    cap = cv2.VideoCapture(os.path.join(vdir, "video.mp4"))
    sha = hashlib.sha1()
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        sha.update(fr.tobytes())
    g["video_decoded_sha1"] = np.array(sha.hexdigest())
    for mode in ("conventional", "pseudo"):
        dataset, calib = get_data(vdir, ff.IMG_SIZE, rect_mode=mode)
        assert type(dataset).__name__ == "StereoVideoDataset"
        put_calib(f"video_{mode}_", calib)
        items = list(dataset)
        g[f"video_{mode}_left"] = np.stack([np.asarray(it[0]) for it in items])[..., 1::3, 2::3]
        g[f"video_{mode}_right"] = np.stack([np.asarray(it[1]) for it in items])[..., 1::3, 2::3]
        g[f"video_{mode}_mask"] = pack(np.stack([np.asarray(it[2]) for it in items]))
        g[f"video_{mode}_pose"] = np.stack([np.asarray(it[3]) for it in items])
        g[f"video_{mode}_number"] = np.array([it[4] for it in items])
        g[f"video_{mode}_len"] = np.array(len(dataset))
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filedata", action="store_true",
                    help="only tests/golden/file_dataset.npz (+ tests/golden/file_video/): the reference's get_data on file fixtures")
    ap.add_argument("--resize", action="store_true", help="only tests/golden/resize_stereo.npz (reference dataset/transforms.ResizeStereo)")
    ap.add_argument("--metrics", action="store_true", help="only tests/golden/metrics.npz (reference evaluate_ate_freiburg.eval)")
    ap.add_argument("--config1", action="store_true",
                    help="only tests/golden/e2e_cfg1_tartan.npz: the reference run on its own tests/test_data/tartan_air fixtures")
    ap.add_argument("--bench64", action="store_true",
                    help="only tests/golden/bench64_poses.npz: the reference's poses of the 64 frame pairs bench.py times")
    ap.add_argument("--full", action="store_true", help="also write the 640x512 dump to oracle/_ref")
    ap.add_argument("--skip-small", action="store_true")
    ap.add_argument("--variants", action="store_true",
                    help="BASELINE configs 4/5: infer_f2f_nw (no confidence heads) trajectory -> tests/golden, only3d 1280x1024 -> oracle/_ref")
    ap.add_argument("--skip-nw", action="store_true", help="with --variants: only the (git-ignored) only3d 1280x1024 golden")
    ap.add_argument("--mask-spec", action="store_true", help="only tests/golden/mask_specularities.npz (reference dataset function)")
    args = ap.parse_args()
    assert os.path.isdir(REF), "reference not mounted"
    if args.filedata:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "file_dataset.npz"), **filedata_golden())
        print("tests/golden/file_dataset.npz written")
        return
    if args.resize:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_stereo.npz"), **resize_golden())
        print("tests/golden/resize_stereo.npz written")
        return
    if args.metrics:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **metrics_golden())
        print("tests/golden/metrics.npz written")
        return
    if args.mask_spec:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mask_specularities.npz"), **mask_spec_golden())
        print("tests/golden/mask_specularities.npz written")
        return
    torch.manual_seed(0)
    torch.set_num_threads(8)
    gold = os.path.join(ROOT, "tests", "golden")
    refdir = os.path.join(HERE, "_ref")
    os.makedirs(gold, exist_ok=True)
    os.makedirs(os.path.join(refdir, "trained"), exist_ok=True)
    for name in ("poseNet_2xf8up4b.pth", "only3d_1a7ix98y.pth", "only2d_18bl5o77.pth"):
        dst = os.path.join(refdir, "trained", name)
        if not os.path.isfile(dst):
            shutil.copyfile(os.path.join(REF, "trained", name), dst)
    ckpt = os.path.join(REF, "trained", "poseNet_2xf8up4b.pth")

    if args.config1:
        print("config 1: tartan_air fixtures, 640x512, frames [0, 1, 0] -> 2 pairs, conf heads on")
        L, R, K, bf = tartan_frames()
        order = [0, 1, 0]
        M = np.ones((3, 1, 512, 640), bool)
        seq = FixedFrames(L[order], R[order], M, K, bf)
        out, rec, est = run_sequence((640, 512), seed=0, n_frames=3, ckpt=ckpt, seq=seq)
        out.update({f"s_{k}": v for k, v in stage_dict(rec, est, full=False).items()})
        out["imgs_l"], out["imgs_r"], out["order"] = L, R, np.array(order)             # the two distinct frames only
        out["n_evals"] = np.array([out[f"pair{k}_eval_pose"].shape[0] for k in range(2)])
        # full-resolution fp32 flows (identical-input mask gate) go to the git-ignored oracle/_ref; the committed file keeps the 1/4 grid
        np.savez_compressed(os.path.join(refdir, "golden_cfg1_flows.npz"), time_flow=out.pop("s_time_flow"),
                            stereo_flow2=out.pop("s_stereo_flow2"))
        out["s_time_flow_ds4"] = rec.stage["time_flow"][0, :, ::4, ::4].numpy()
        out["s_stereo_flow2_ds4"] = rec.stage["stereo_flow2"][0, :, ::4, ::4].numpy()
        np.savez_compressed(os.path.join(gold, "e2e_cfg1_tartan.npz"), **out)
        print("tests/golden/e2e_cfg1_tartan.npz written")
        return
    if args.bench64:
        import hashlib
        print("bench inputs: 65-frame bench_sequence() -> 64 reference poses")
        synth = _load_synth()
        seq = synth.bench_sequence()
        out, rec, est = run_sequence((640, 512), seed=0, n_frames=synth.BENCH_FRAMES, ckpt=ckpt, seq=seq, quiet=True)
        n = synth.BENCH_FRAMES - 1
        keep = {"rel_pose": out["rel_pose"], "traj": out["traj"], "K": out["K"], "bf": out["bf"],
                "n_evals": np.array([out[f"pair{k}_eval_pose"].shape[0] for k in range(n)]),
                "frames_sha1": np.frombuffer(hashlib.sha1(out["imgs_l"].tobytes() + out["imgs_r"].tobytes()
                                                          + out["masks_in"].tobytes()).digest(), dtype=np.uint8)}
        np.savez_compressed(os.path.join(gold, "bench64_poses.npz"), **keep)
        print("tests/golden/bench64_poses.npz written")
        return
    if not args.skip_small:
        print("stage goldens (reference operators on exact inputs)")
        np.savez_compressed(os.path.join(gold, "stages_small.npz"), **small_stage_goldens())
        print("pose-head unit recipe")
        np.savez_compressed(os.path.join(gold, "posehead_unit.npz"), **posehead_unit_golden())
        print("e2e 384x352, 3 frames, conf heads on")
        out, rec, est = run_sequence((384, 352), seed=1, n_frames=3, ckpt=ckpt)
        out.update({f"s_{k}": v for k, v in stage_dict(rec, est, full=False).items()})
        np.savez_compressed(os.path.join(gold, "e2e_384x352.npz"), **out)
        np.savez_compressed(os.path.join(gold, "posehead_real.npz"), **posehead_real_golden(rec, est))
    if args.full:
        print("e2e 640x512, 3 frames (full dump)")
        out, rec, est = run_sequence((640, 512), seed=0, n_frames=3, ckpt=ckpt, rec_corr=True)
        out.update({f"s_{k}": v for k, v in stage_dict(rec, est, full=True).items()})
        cb = rec.corr[-1]                                   # the batch-2 RAFT pass of the last pair
        out["c_fmap1"] = cb._rec["fmap1"].numpy()
        out["c_fmap2"] = cb._rec["fmap2"].numpy()
        for it in (0, 11):
            out[f"c_coords{it}"] = cb._rec["lookups"][it][0].numpy()
            out[f"c_lookup{it}"] = cb._rec["lookups"][it][1].numpy()
        np.savez(os.path.join(refdir, "golden_full.npz"), **out)
    if args.variants:
        from core.utils.trajectory import save_trajectory
        from lietorch import SE3
        import tempfile
    if args.variants and not args.skip_nw:
        print("infer_f2f_nw: 384x352, 5 frames, conf_weighing False, trajectory.freiburg")
        out, rec, est = run_sequence((384, 352), seed=3, n_frames=5, ckpt=ckpt, conf=False)
        traj = [{"camera-pose": SE3(torch.from_numpy(p)[None]), "timestamp": i} for i, p in enumerate(out["traj"])]
        with tempfile.TemporaryDirectory() as td:
            save_trajectory(traj, td)                                  # the reference's writer (core/utils/trajectory.py:17-23)
            out["freiburg"] = np.frombuffer(open(os.path.join(td, "trajectory.freiburg"), "rb").read(), dtype=np.uint8)
        out["n_evals"] = np.array([out[f"pair{k}_eval_pose"].shape[0] for k in range(4)])
        keep = ("K", "bf", "size", "seed", "imgs_l", "imgs_r", "masks_in", "traj", "freiburg", "n_evals")
        np.savez_compressed(os.path.join(gold, "e2e_nw_384x352.npz"), **{k: out[k] for k in keep})
    if args.variants:
        print("only3d_1a7ix98y.pth, loss_weight[1] = 0, 1280x1024, 2 frames")
        out, rec, est = run_sequence((1280, 1024), seed=4, n_frames=2, ckpt=os.path.join(REF, "trained", "only3d_1a7ix98y.pth"),
                                     loss_weight_mask=(1.0, 0.0))
        st = rec.stage
        out["n_evals"] = np.array([out["pair0_eval_pose"].shape[0]])
        out["s_time_flow_ds4"] = st["time_flow"][0, :, ::4, ::4].numpy()
        out["s_stereo_flow2_ds4"] = st["stereo_flow2"][0, :, ::4, ::4].numpy()
        out["s_mask2w"] = pack(st["mask2w"][0, 0].numpy())
        np.savez_compressed(os.path.join(refdir, "golden_only3d_1280x1024.npz"), **out)
    print("done")


if __name__ == "__main__":
    main()
