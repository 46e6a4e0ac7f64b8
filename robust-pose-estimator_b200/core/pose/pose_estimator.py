"""Frame-to-frame tracker with the reference's interface
(/root/reference/core/pose/pose_estimator.py:11-160).  Only the f2f branch exists here: the
frame-to-model branch (SurfelMap) is out of scope.

Differences that do not change results:
  * the failure guard (NaN or |log| > 0.1 -> identity, pose_estimator.py:81-85) is evaluated on the
    device without a host synchronisation; the flags are kept in ``failure_flags`` and the reference's
    RuntimeWarning is raised lazily by ``check_failures()`` (or immediately with config['sync_guard']).
"""
import warnings
from collections import OrderedDict

import torch

from ...lie import SE3
from ..utils.frame_class import Frame
from .pose_net import PoseNet


class PoseEstimator(torch.nn.Module):
    def __init__(self, config, intrinsics, baseline, checkpoint, img_shape, init_pose=None):
        """
        :param config: the ``slam`` section of configuration/infer_f2f.yaml (+ optional keys
                       ``precision`` fp32|fp16x3|tf32|bf16|fp16, ``solver`` lbfgs_ref|gn, ``residuals`` 2d3d|3d|2d,
                       ``sync_guard``, ``cuda_graph``: replay the whole per-frame device work of ``forward`` as one CUDA
                       graph -- the latency path; ``flow`` / ``weights`` are then not returned)
        :param intrinsics: rectified camera intrinsics (3,3)
        :param baseline: stereo baseline x focal length in pixel * mm ("bf")
        :param checkpoint: path of a reference checkpoint (trained/*.pth), or a dict {state_dict, config},
                           or None for random initialisation of the same architecture
        :param img_shape: (W, H) like the reference's ``img_size``
        :param init_pose: SE3 of shape (1,), defaults to identity
        """
        super().__init__()
        if not config.get("frame2frame", True):
            raise NotImplementedError("only the frame-to-frame configuration is implemented on this path")
        if checkpoint is None:
            checkp = {"config": {"model": {"iters": 12, "dropout": 0.0, "small": False}}, "state_dict": None}
        elif isinstance(checkpoint, dict):
            checkp = checkpoint
        else:
            checkp = torch.load(checkpoint, map_location="cpu", weights_only=False)
        mcfg = dict(checkp["config"]["model"])
        mcfg["image_shape"] = (img_shape[1], img_shape[0])
        mcfg["lbgfs_iters"] = config["lbgfs_iters"]
        mcfg["use_weights"] = config["conf_weighing"]
        mcfg["precision"] = config.get("precision", "fp32")
        mcfg["solver"] = config.get("solver", "lbfgs_ref")
        model = PoseNet(mcfg)
        if checkp["state_dict"] is not None:
            model.load_state_dict(OrderedDict((k.replace("module.", ""), v) for k, v in checkp["state_dict"].items()))
        # the shipped code always adds both residual terms (pose_head.py:53-58); '3d' / '2d' zero the other
        # term's loss weight (index 1 = 2-D reprojection, index 0 = 3-D point-to-point)
        residuals = config.get("residuals", "2d3d")
        if residuals not in ("2d3d", "3d", "2d"):
            raise ValueError(f"residuals must be '2d3d', '3d' or '2d', got {residuals!r}")
        if residuals != "2d3d":
            with torch.no_grad():
                model.loss_weight[1 if residuals == "3d" else 0] = 0.0
        model.eval()
        self.model = model
        self.intrinsics = intrinsics.unsqueeze(0).float()
        self.register_buffer("scale", torch.tensor(1 / config["depth_clipping"][1]), persistent=False)
        self.register_buffer("baseline", torch.tensor(baseline).unsqueeze(0).float(), persistent=False)
        self.last_pose = (SE3.Identity(1) if init_pose is None else init_pose).float()
        self.last_frame = None
        self.frame = None
        self.frame2frame = True
        self.scene = None
        self.config = config
        self.failure_flags = []

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.intrinsics = fn(self.intrinsics)
        self.last_pose = SE3(fn(self.last_pose.data))
        return out

    def forward(self, limg, rimg, mask):
        """limg, rimg (1,3,H,W) float 0..255, mask (1,1,H,W) bool -> (pose SE3 in mm, None, flow, weights)."""
        self.last_pose = self.last_pose.to(limg.device)
        if self.config.get("cuda_graph", False):
            return self._forward_graphed(limg, rimg, mask)
        self.last_frame = self.frame
        self.frame = Frame(limg, rimg, mask=mask)
        rel_pose, ret_frame, flow, weights = self.get_pose_f2f()

        # pose_estimator.py:81-87, evaluated on the device
        bad = torch.isnan(rel_pose.vec()).any() | (torch.abs(rel_pose.log()) > 1.0e-1).any()
        ident = SE3.IdentityLike(self.last_pose)
        rel_pose = SE3(torch.where(bad, ident.data, rel_pose.vec().to(ident.dtype)))
        self.failure_flags.append(bad)
        if self.config.get("sync_guard", False) and bool(bad):
            warnings.warn("pose estimation not converged, skip.", RuntimeWarning)
        self.last_frame = ret_frame
        rel_pose = rel_pose.scale(1 / self.scale)                  # de-normalise depth scaling
        self.last_pose = self.last_pose * rel_pose.inv()           # chain transforms
        return self.last_pose, self.scene, flow, weights

    def _forward_graphed(self, limg, rimg, mask):
        """Latency path: the same per-frame step through ``F2FEngine`` with one frame per chunk, captured once and replayed
        as a single CUDA graph (features / context / stereo depth of the previous frame are the carried state, as in
        ``Frame``).  Guard, de-normalisation and chaining as in ``forward`` (pose_estimator.py:81-91)."""
        from ...engine import F2FEngine
        eng = getattr(self, "_stream_engine", None)
        if eng is None:
            eng = self._stream_engine = F2FEngine(self, 1, True)
        if self.frame is None:                                     # new sequence
            eng.reset()
        self.last_frame = self.frame
        self.frame = Frame(limg, rimg, mask=mask)
        rel, log, _ = eng.infer_sequence(limg, rimg, mask)
        ident = SE3.IdentityLike(self.last_pose)
        if rel.shape[0] == 0:                                      # first frame: stereo depth only
            return self.last_pose, self.scene, None, None
        bad = torch.isnan(rel).any() | (torch.abs(log) > 1.0e-1).any()
        rel_pose = SE3(torch.where(bad, ident.data, rel.to(ident.dtype)))
        self.failure_flags.append(bad)
        if self.config.get("sync_guard", False) and bool(bad):
            warnings.warn("pose estimation not converged, skip.", RuntimeWarning)
        rel_pose = rel_pose.scale(1 / self.scale)
        self.last_pose = self.last_pose * rel_pose.inv()
        return self.last_pose, self.scene, None, None

    def get_pose_f2f(self):
        flow = None
        if self.last_frame is None:
            rel = SE3.IdentityLike(self.last_pose)
            depth, stereo_flow, valid = self.model.flow2depth(self.frame.img, self.frame.rimg, self.baseline * self.scale)
            self.frame.depth = depth / self.scale
            self.frame.flow = stereo_flow                            # `valid` is discarded
            return rel, None, None, None
        rel, depth1, depth2, weights, flow, stereo_flow = self.model.infer(
            self.last_frame.img, self.frame.img, self.intrinsics, self.baseline * self.scale,
            depth1=self.last_frame.depth * self.scale, image2r=self.frame.rimg, mask1=self.last_frame.mask,
            mask2=self.frame.mask, stereo_flow1=self.last_frame.flow, ret_details=True)
        self.frame.depth = depth2 / self.scale
        self.frame.flow = stereo_flow
        return rel, self.last_frame, flow, weights

    def infer_sequence(self, limgs, rimgs, masks, chunk=8, use_graphs=False):
        """Throughput path: a whole sequence (T,3,H,W) / (T,1,H,W) on the device -- or in (pinned) host memory, then uploaded
        chunk by chunk behind the compute of the previous chunk -- -> absolute poses
        (T,7) float32 on the HOST (mm; row 0 = initial pose, like the reference's trajectory list) and the
        per-pair failure flags (T-1,).  Pairs are solved in chunks by ``F2FEngine``; the trajectory is composed
        on the host with rpe_compose_trajectory_host (one device->host copy of (T-1) x 13 floats)."""
        import ctypes as C

        from ... import _lib
        rel, log, evals = self.infer_pairs(limgs, rimgs, masks, chunk, use_graphs)
        n = rel.shape[0]
        host = torch.cat((rel, log), 1).cpu().contiguous()                      # the only device->host transfer
        rel_h, log_h = host[:, :7].contiguous(), host[:, 7:].contiguous()
        init = self.last_pose.data.reshape(7).float().cpu().contiguous()
        out = torch.empty((n + 1, 7), dtype=torch.float32)
        failed = torch.zeros((max(n, 1),), dtype=torch.uint8)
        inv_scale = float((1 / self.scale).float().cpu())
        _lib.check(_lib.lib().rpe_compose_trajectory_host(C.c_void_p(rel_h.data_ptr()), C.c_void_p(log_h.data_ptr()), n,
                                                          C.c_void_p(init.data_ptr()), inv_scale,
                                                          C.c_void_p(out.data_ptr()), C.c_void_p(failed.data_ptr())),
                   "rpe_compose_trajectory_host")
        self.last_pose = SE3(out[-1:].clone().to(self.last_pose.device))
        self.last_evals = evals
        return out, failed[:n].bool()

    def infer_pairs(self, limgs, rimgs, masks, chunk=8, use_graphs=False, sequence_start=True):
        """Relative poses of the consecutive frame pairs of a (shard of a) sequence, device-resident: (rel (T-1,7) float32 in
        normalised units, log (T-1,6), evals (T-1,)).  Frames on the device or in (pinned) host memory.  ``sequence_start``:
        frame 0 is the first frame of the whole sequence (it alone keeps its un-and-ed mask); False for the halo
        frame of a later shard (``parallel.infer_sequence_sharded``)."""
        from ...engine import F2FEngine
        key = (chunk, use_graphs)
        if getattr(self, "_engine_key", None) != key:
            self._engine, self._engine_key = F2FEngine(self, chunk, use_graphs), key
        self._engine.reset()
        if limgs.device.type == "cpu":
            return self._infer_from_host(limgs, rimgs, masks, chunk, sequence_start)
        return self._engine.infer_sequence(limgs, rimgs, masks, sequence_start)

    def _infer_from_host(self, limgs, rimgs, masks, chunk, sequence_start=True):
        """``infer_sequence`` fed from HOST tensors (pinned uint8 / float frames, bool masks): the frames of engine chunk k+1
        are uploaded on a copy stream while chunk k is being solved, so the host->device transfer is hidden behind compute."""
        dev = self.baseline.device
        T = limgs.shape[0]
        bounds, a = [], 0
        while a < T:
            b = min(a + chunk + (1 if a == 0 else 0), T)              # the first chunk carries the first frame as well
            bounds.append((a, b))
            a = b
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy, main = self._copy_stream, torch.cuda.current_stream(dev)

        def upload(a, b):
            copy.wait_stream(main)                                    # buffers freed on the main stream may be reused here
            with torch.cuda.stream(copy):
                t = [x[a:b].to(dev, non_blocking=True) for x in (limgs, rimgs, masks)]
                ev = torch.cuda.Event()
                ev.record(copy)
            return t, ev

        nxt = upload(*bounds[0])
        rel, log, evals = [], [], []
        for i in range(len(bounds)):
            (l, r, m), ev = nxt
            if i + 1 < len(bounds):
                nxt = upload(*bounds[i + 1])
            main.wait_event(ev)
            for x in (l, r, m):
                x.record_stream(main)
            p, lg, e = self._engine.infer_sequence(l, r, m.bool(), sequence_start)        # uint8 frames are consumed as they are
            rel.append(p), log.append(lg), evals.append(e)
        return torch.cat(rel), torch.cat(log), torch.cat(evals)

    def check_failures(self):
        """Synchronise and raise the reference's warning for every failed pair; returns their indices."""
        if not self.failure_flags:
            return []
        idx = torch.stack(self.failure_flags).nonzero().flatten().tolist()
        for _ in idx:
            warnings.warn("pose estimation not converged, skip.", RuntimeWarning)
        return idx

    @property
    def device(self):
        return self.last_pose.device

    def get_last_frame(self):
        return self.last_frame

    def get_frame(self):
        return self.frame
