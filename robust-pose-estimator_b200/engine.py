"""Batched frame-to-frame engine: the throughput path behind ``PoseEstimator.infer_sequence``.

Frame pairs are independent in the f2f configuration, so a sequence is processed in
chunks of C consecutive frames.  Per chunk (all on one CUDA stream, optionally replayed as one CUDA graph):

  fnet over the 2C left/right images, cnet over the C left images          (once per image: the reference
                                                                             recomputes the features of
                                                                             frame k twice)
  one RAFT refinement over 2C samples: C temporal pairs (k-1 -> k) + C stereo pairs (left k -> right k)
      rpe_corr_build / 12 x rpe_corr_lookup / rpe_convex_upsample8          (sm_100a kernels)
  rpe_depth_proj, rpe_proj, rpe_warp8_mask, rpe_downsample8_cat, confidence heads, rpe_pose_solve (n = C)

State carried between chunks = the last frame's features, context, normalised depth, stereo flow, mask and
image, exactly what the reference keeps in ``Frame`` (pose_estimator.py:62-63,115-122).  The arithmetic per
pair is the same as the per-frame tracker; only batch composition differs (eval-mode BatchNorm and
InstanceNorm are per-sample, so results do not depend on how frames are grouped into chunks).  The first frame of a
sequence rides in the first chunk (its stereo pair is one more RAFT sample)."""
import torch

from . import ops
from .core.unet.unet import tiny_unet_forward


class _FrameState:
    """What the tracker carries from a frame to the next pair (the reference's ``Frame``): image, normalised depth, stereo flow,
    mask, and the frame's encoder outputs -- NCHW fp32 (``fmap``, ``net``, ``inp``) on the generic path, or on the tensor-core
    path the same as NHWC tensors ready for the kernels: ``fmap_hi/lo`` (feature planes), ``net_h`` (fp32) + ``net_hi/lo`` and
    ``inp_hi/lo`` (planes)."""
    __slots__ = ("img", "fmap", "net", "inp", "depth", "sflow", "mask", "fmap_hi", "fmap_lo", "net_h", "net_hi", "net_lo", "inp_hi", "inp_lo")

    def __init__(self):
        for k in self.__slots__:
            setattr(self, k, None)

    def fields(self):
        return [k for k in self.__slots__ if getattr(self, k) is not None]


class F2FEngine:
    def __init__(self, estimator, chunk=8, use_graphs=False):
        self.est = estimator
        self.model = estimator.model
        self.chunk = int(chunk)
        self.use_graphs = use_graphs
        self._graphs = {}
        self._scale = float(estimator.scale)        # host copy: no device sync inside (graph-captured) chunks
        self.keep_solve_inputs = False              # bench probe: keep the last chunk's rpe_pose_solve arguments alive
        self.last_solve_inputs = None
        self.reset()

    def reset(self):
        self.prev = None

    # ------------------------------------------------------------------------------------------------
    def _first_frame(self, limg, rimg, mask, sequence_start=True):
        """First frame of a (shard of a) sequence: stereo depth only.  For the first frame of the SEQUENCE the
        stereo validity is NOT and-ed into the mask; for the halo frame of a later shard it is."""
        raft = self.model.flow
        if raft.precision == "fp16x3":
            return self._first_frame_tc(limg, rimg, mask, sequence_start)
        fl, fr, net, inp = raft.encode(limg, rimg)
        f = torch.cat((fl, fr), 0)
        preds, _, _, _ = raft.refine(f[0:1].contiguous(), f[1:2].contiguous(), net, inp)
        bl = (self.est.baseline * self.est.scale).float().reshape(1)
        eye = torch.eye(3, device=limg.device)[None]
        m = mask.clone()
        depth, _, _ = ops.depth_proj(preds[-1], bl, eye, None if sequence_start else m, want_pcl=False)
        st = _FrameState()
        st.img, st.fmap, st.net, st.inp = limg, f[0:1].contiguous(), net, inp
        st.depth, st.sflow, st.mask = depth, preds[-1], m
        return st

    def _first_frame_tc(self, limg, rimg, mask, sequence_start=True):
        """``_first_frame`` on the tensor-core trunk: one stereo sample through the same buffers as ``_chunk_body_tc``."""
        from .tc import Planes
        raft, est = self.model.flow, self.est
        dev = limg.device
        H, W = limg.shape[-2:]
        h8, w8 = H // 8, W // 8
        utc = raft.update_tc()
        st = utc.state(1, h8, w8, dev)
        feat = raft.feature_list(("first", h8, w8, dev.index), 3, h8, w8, dev)
        imgs = torch.cat((limg, rimg), 0)
        imgs = imgs.contiguous() if imgs.dtype == torch.uint8 else imgs.float().contiguous()
        raft.encode_into(imgs, 1, feat.view(0, 2), st["h"], st["hp"], st["inp"])
        new = _FrameState()
        new.fmap_hi, new.fmap_lo = feat.hi[0:1].clone(), feat.lo[0:1].clone()
        new.net_h = st["h"][0:1].clone()
        new.net_hi, new.net_lo = st["hp"].hi[0:1].clone(), st["hp"].lo[0:1].clone()
        new.inp_hi, new.inp_lo = st["inp"].hi[0:1].clone(), st["inp"].lo[0:1].clone()
        pyr = ops.CorrPyramid.from_planes(feat, feat.view(1), 1, radius=raft.config["corr_radius"])
        flow_up, _ = utc.refine_state(pyr, 1, h8, w8, dev)
        bl = (est.baseline * est.scale).float().reshape(1)
        eye = torch.eye(3, device=dev)[None]
        m = mask.clone()
        depth, _, _ = ops.depth_proj(flow_up, bl, eye, None if sequence_start else m, want_pcl=False)
        new.img, new.depth, new.sflow, new.mask = limg, depth, flow_up, m
        return new

    def _chunk_body(self, prev, limg, rimg, mask, sequence_start=True):
        """C new frames given the state of the frame before them -> (pose (C,7), log (C,6), evals (C,), new state).
        ``prev is None``: ``limg[0]`` is the first frame of the (shard of the) sequence and rides along in the same batch --
        its stereo pair is one more RAFT sample of this chunk instead of a batch-1 pass of its own; C = len(limg) - 1."""
        if self.model.flow.precision == "fp16x3" and (prev is None or prev.fmap_hi is not None):
            return self._chunk_body_tc(prev, limg, rimg, mask, sequence_start)
        first = prev is None
        n_img = limg.shape[0]
        C = n_img - 1 if first else n_img
        raft, model, est = self.model.flow, self.model, self.est
        fL, fR, net0, inp = raft.encode(limg, rimg)
        if first:
            fL_prev, net_prev, inp_prev, fL_new = fL[:-1], net0[:-1], inp[:-1], fL[1:]
        else:
            fL_prev = torch.cat((prev.fmap, fL[:-1]), 0)
            net_prev = torch.cat((prev.net, net0[:-1]), 0)
            inp_prev = torch.cat((prev.inp, inp[:-1]), 0)
            fL_new = fL
        # samples [0, C): temporal pairs (k-1 -> k); samples [C, C + n_img): stereo pairs of the frames of this batch
        preds, gru, ctx, _ = raft.refine(torch.cat((fL_prev, fL), 0).contiguous(), torch.cat((fL_new, fR), 0).contiguous(),
                                         torch.cat((net_prev, net0), 0), torch.cat((inp_prev, inp), 0))
        time_flow = preds[-1][:C].contiguous()
        sflow_all = preds[-1][C:].contiguous()
        K = est.intrinsics.float().expand(n_img, 3, 3).contiguous()
        bl = (est.baseline * est.scale).float().reshape(1).expand(n_img).contiguous()
        mask_all = mask.clone()
        keep0 = mask_all[0:1].clone() if (first and sequence_start) else None
        depth_all, _, pcl_all = ops.depth_proj(sflow_all, bl, K, mask_all)          # mask &= stereo validity
        if keep0 is not None:
            mask_all[0:1] = keep0              # the first frame of the SEQUENCE keeps its input mask
        if first:
            limg_new, mask2, depth2, pcl2, sflow = limg[1:], mask_all[1:].contiguous(), depth_all[1:], pcl_all[1:].contiguous(), sflow_all[1:].contiguous()
            depth_prev, img_prev = depth_all[:-1].contiguous(), limg[:-1].contiguous()
            sflow_prev, mask1 = sflow_all[:-1].contiguous(), mask_all[:-1].contiguous()
        else:
            limg_new, mask2, depth2, pcl2, sflow = limg, mask_all, depth_all, pcl_all, sflow_all
            depth_prev = torch.cat((prev.depth, depth2[:-1]), 0).contiguous()
            img_prev = torch.cat((prev.img, limg[:-1]), 0).contiguous()
            sflow_prev = torch.cat((prev.sflow, sflow[:-1]), 0).contiguous()
            mask1 = torch.cat((prev.mask, mask2[:-1]), 0).contiguous()
        Kc = K[:C].contiguous()
        pcl1 = ops.proj(depth_prev, Kc, rescale=self._scale)                 # (d / scale) * scale round trip
        conf1, conf2, pcl2w, mask2w = model.get_weight_maps(pcl1, pcl2, img_prev, limg_new.contiguous(), mask2, time_flow, sflow_prev,
                                                            sflow, gru[:C], ctx[:C])
        lw = model.loss_weight[None, :].float().expand(C, 2).contiguous()
        head = model.pose_head.problem
        mode = ops.SOLVER_GN if head.solver == "gn" else ops.SOLVER_LBFGS_REF
        iters = head.gn_iters if head.solver == "gn" else head.lbgfs_iters
        sol = ops.pose_solve(time_flow, pcl1, pcl2w, conf1, conf2, mask1, mask2w, Kc, lw, mode=mode, max_iter=iters)
        if self.keep_solve_inputs:
            self.last_solve_inputs = (time_flow, pcl1, pcl2w, conf1, conf2, mask1, mask2w, Kc, lw)
        st = _FrameState()
        st.img, st.fmap, st.net, st.inp = limg[-1:], fL[-1:].contiguous(), net0[-1:], inp[-1:]
        st.depth, st.sflow, st.mask = depth_all[-1:], sflow_all[-1:], mask_all[-1:]
        return sol.pose, sol.log, sol.n_evals, st

    # ------------------------------------------------------------------------------------------------
    def _chunk_body_tc(self, prev, limg, rimg, mask, sequence_start=True):
        """``_chunk_body`` on the tensor-core trunk without layout round trips: the feature encoder writes NHWC split planes
        into ONE image list [previous left | left 0..C-1 | right 0..C-1] from which a single rpe_corr_build_planes launch forms the
        C temporal and the C (+1) stereo volumes; the context encoder writes tanh(net) / relu(inp) straight into the update
        operator's state buffers; the confidence heads read the final GRU state and the context from those buffers in place."""
        from .tc import Planes
        first = prev is None
        n_img = limg.shape[0]
        C = n_img - 1 if first else n_img
        off = 0 if first else 1                                   # image-list slot of left image 0
        raft, model, est = self.model.flow, self.model, self.est
        dev = limg.device
        H, W = limg.shape[-2:]
        h8, w8 = H // 8, W // 8
        B = C + n_img                                             # samples [0, C): temporal pairs, [C, B): stereo pairs
        utc = raft.update_tc()
        st = utc.state(B, h8, w8, dev)
        feat = raft.feature_list(("chunk", n_img, first, h8, w8, dev.index), off + 2 * n_img + 1, h8, w8, dev)
        imgs = torch.cat((limg, rimg), 0)                         # uint8 frames stay uint8: the kernels convert on load
        imgs = imgs.contiguous() if imgs.dtype == torch.uint8 else imgs.float().contiguous()
        raft.encode_into(imgs, n_img, feat.view(off, off + 2 * n_img), st["h"][C:], st["hp"].view(C, B), st["inp"].view(C, B))
        # temporal sample k reads the context of frame k-1: copies of the stereo slots (and the carried frame on a later chunk)
        n_shift = C - off
        if n_shift > 0:
            st["h"][off:C].copy_(st["h"][C:C + n_shift])
            st["hp"].view(off, C).copy_(st["hp"].view(C, C + n_shift))
            st["inp"].view(off, C).copy_(st["inp"].view(C, C + n_shift))
        if not first:
            feat.hi[0:1].copy_(prev.fmap_hi), feat.lo[0:1].copy_(prev.fmap_lo)
            st["h"][0:1].copy_(prev.net_h)
            st["hp"].hi[0:1].copy_(prev.net_hi), st["hp"].lo[0:1].copy_(prev.net_lo)
            st["inp"].hi[0:1].copy_(prev.inp_hi), st["inp"].lo[0:1].copy_(prev.inp_lo)
        new = _FrameState()                                       # encoder outputs of the last frame, before the GRU evolves them
        new.fmap_hi, new.fmap_lo = feat.hi[off + n_img - 1:off + n_img].clone(), feat.lo[off + n_img - 1:off + n_img].clone()
        new.net_h = st["h"][B - 1:B].clone()
        new.net_hi, new.net_lo = st["hp"].hi[B - 1:B].clone(), st["hp"].lo[B - 1:B].clone()
        new.inp_hi, new.inp_lo = st["inp"].hi[B - 1:B].clone(), st["inp"].lo[B - 1:B].clone()
        pyr = ops.CorrPyramid.from_planes(feat, feat.view(1), B, radius=raft.config["corr_radius"], f1_wrap=C, f1_sub=C - off)
        flow_up, _ = utc.refine_state(pyr, B, h8, w8, dev)
        del pyr
        time_flow = flow_up[:C]
        sflow_all = flow_up[C:]
        K = est.intrinsics.float().expand(n_img, 3, 3).contiguous()
        bl = (est.baseline * est.scale).float().reshape(1).expand(n_img).contiguous()
        mask_all = mask.clone()
        keep0 = mask_all[0:1].clone() if (first and sequence_start) else None
        depth_all, _, pcl_all = ops.depth_proj(sflow_all, bl, K, mask_all)          # mask &= stereo validity
        if keep0 is not None:
            mask_all[0:1] = keep0              # the first frame of the SEQUENCE keeps its input mask
        if first:
            limg_new, mask2, pcl2, sflow = limg[1:], mask_all[1:], pcl_all[1:], sflow_all[1:]
            depth_prev, img_prev = depth_all[:-1], limg[:-1]
            sflow_prev, mask1 = sflow_all[:-1], mask_all[:-1]
        else:
            limg_new, mask2, pcl2, sflow = limg, mask_all, pcl_all, sflow_all
            depth_prev = torch.cat((prev.depth, depth_all[:-1]), 0)
            img_prev = torch.cat((prev.img, limg[:-1]), 0)
            sflow_prev = torch.cat((prev.sflow, sflow_all[:-1]), 0)
            mask1 = torch.cat((prev.mask, mask_all[:-1]), 0)
        Kc = K[:C]
        pcl1 = ops.proj(depth_prev.contiguous(), Kc, rescale=self._scale)        # (d / scale) * scale round trip
        conf1, conf2, pcl2w, mask2w = model.get_weight_maps(pcl1, pcl2.contiguous(), img_prev.contiguous(), limg_new.contiguous(),
                                                            mask2.contiguous(), time_flow, sflow_prev.contiguous(), sflow.contiguous(),
                                                            None, None, state_planes=(st["hp"], st["inp"]))
        lw = model.loss_weight[None, :].float().expand(C, 2).contiguous()
        head = model.pose_head.problem
        mode = ops.SOLVER_GN if head.solver == "gn" else ops.SOLVER_LBFGS_REF
        iters = head.gn_iters if head.solver == "gn" else head.lbgfs_iters
        mask1 = mask1.contiguous()
        sol = ops.pose_solve(time_flow, pcl1, pcl2w, conf1, conf2, mask1, mask2w, Kc, lw, mode=mode, max_iter=iters)
        if self.keep_solve_inputs:
            self.last_solve_inputs = (time_flow, pcl1, pcl2w, conf1, conf2, mask1, mask2w, Kc, lw)
        new.img = limg[-1:]
        new.depth, new.sflow, new.mask = depth_all[-1:], sflow_all[-1:], mask_all[-1:]
        return sol.pose, sol.log, sol.n_evals, new

    # ------------------------------------------------------------------------------------------------
    def _graphed_chunk(self, prev, limg, rimg, mask):
        """Replay a captured CUDA graph of ``_chunk_body`` (static input / state / output buffers per shape)."""
        key = (tuple(limg.shape), limg.dtype)
        g = self._graphs.get(key)
        if g is None:
            static = {"limg": limg.clone(), "rimg": rimg.clone(), "mask": mask.clone(), "prev": _FrameState()}
            for k in prev.fields():
                setattr(static["prev"], k, getattr(prev, k).clone())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                                    # warm-up outside capture
                for _ in range(2):
                    self._chunk_body(static["prev"], static["limg"], static["rimg"], static["mask"])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static["out"] = self._chunk_body(static["prev"], static["limg"], static["rimg"], static["mask"])
            g = (graph, static)
            self._graphs[key] = g
        graph, static = g
        static["limg"].copy_(limg)
        static["rimg"].copy_(rimg)
        static["mask"].copy_(mask)
        for k in static["prev"].fields():
            getattr(static["prev"], k).copy_(getattr(prev, k))
        graph.replay()
        pose, log, evals, st = static["out"]
        new = _FrameState()
        for k in st.fields():
            setattr(new, k, getattr(st, k).clone())
        return pose.clone(), log.clone(), evals.clone(), new

    # ------------------------------------------------------------------------------------------------
    def infer_sequence(self, limgs, rimgs, masks, sequence_start=True):
        """limgs, rimgs (T,3,H,W) 0..255 on the device -- float like the reference's tensors, or uint8 as decoded (the tensor-core
        path reads uint8 directly; the other precisions convert) --, masks (T,1,H,W) bool.
        -> relative poses (T-1,7) f32 (normalised units, frame k-1 -> k), tangents (T-1,6), evals (T-1,).
        Continues from the previous call's last frame if ``reset()`` was not called."""
        T = limgs.shape[0]
        poses, logs, evals = [], [], []
        if limgs.dtype == torch.uint8 and self.model.flow.precision != "fp16x3":
            limgs, rimgs = limgs.float(), rimgs.float()
        with torch.no_grad():
            start = 0
            if self.prev is None:
                if T == 1:
                    self.prev = self._first_frame(limgs[0:1], rimgs[0:1], masks[0:1], sequence_start)
                    start = 1
                else:                      # the first frame joins the first chunk (no batch-1 stereo pass of its own)
                    start = min(1 + self.chunk, T)
                    p, l, e, self.prev = self._chunk_body(None, limgs[0:start], rimgs[0:start], masks[0:start], sequence_start)
                    poses.append(p), logs.append(l), evals.append(e)
            for a in range(start, T, self.chunk):
                b = min(a + self.chunk, T)
                args = (self.prev, limgs[a:b], rimgs[a:b], masks[a:b])
                if self.use_graphs and (b - a) == self.chunk and ops._timers is None:   # events cannot be recorded in a capture
                    p, l, e, self.prev = self._graphed_chunk(*args)
                else:
                    p, l, e, self.prev = self._chunk_body(*args)
                poses.append(p), logs.append(l), evals.append(e)
        if not poses:
            dev = limgs.device
            return torch.zeros((0, 7), device=dev), torch.zeros((0, 6), device=dev), torch.zeros((0,), device=dev)
        return torch.cat(poses), torch.cat(logs), torch.cat(evals)
