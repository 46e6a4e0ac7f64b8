"""PoseNet with the reference's interface (/root/reference/core/pose/pose_net.py:13-164) on the
B200-native path: RAFT trunk (tcgen05 convolution kernels with precision 'fp16x3', a cuDNN trunk otherwise),
correlation / lookup / up-sampling, stereo-depth lifting, back-projection, flow warping, the 1/8 down-sampling and
the SE(3) solve as sm_100a kernels; only the two small confidence heads run through cuDNN.

``state_dict`` layout equals the reference's (loss_weight, flow.*, weight_head_2d.0.*, weight_head_3d.0.*),
so trained/*.pth load unchanged.  CUDA only: CPU tensors raise (no fallback path)."""
from collections import OrderedDict

import torch
from torch import nn

from ... import ops
from ...lie import SE3
from ..RAFT.core.raft import RAFT
from ..unet.unet import tiny_unet_entries, tiny_unet_forward
from ..unet.unet_tc import HeadsTC
from ..utils.param_tree import ParamTree, build_tree
from .pose_head import DeclarativeLayerLie, DPoseSE3Head


class PoseNet(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.loss_weight = nn.Parameter(torch.tensor([1.0, 1.0]), requires_grad=False)
        H, W = config["image_shape"]
        self.image_shape = (H, W)
        self.use_weights = config["use_weights"]
        self.flow = RAFT(config)
        self.pose_head = DeclarativeLayerLie(DPoseSE3Head(None, config["lbgfs_iters"], solver=config.get("solver", "lbfgs_ref")))
        # nn.Sequential(TinyUNet, Sigmoid) in the reference -> keys "weight_head_2d.0.*"
        self.weight_head_2d = ParamTree()
        self.weight_head_2d.add_module("0", build_tree(tiny_unet_entries("", 128 + 128 + 8)))
        self.weight_head_3d = ParamTree()
        self.weight_head_3d.add_module("0", build_tree(tiny_unet_entries("", 128 + 128 + 8 + 8)))
        self._Wh = None
        self._heads_tc = None
        self._head_planes = {}

    # ---- parameter plumbing ------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._Wh = self._heads_tc = None
        self._head_planes = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = OrderedDict((k.replace("module.", ""), v) for k, v in state_dict.items())
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._Wh = self._heads_tc = None
        self._head_planes = {}
        self.flow.invalidate()          # nn.Module.load_state_dict does not call the child's override: drop every packed-weight cache
        return out

    def _head_weights(self):
        if self._Wh is None:
            W = self.weight_head_2d.table("weight_head_2d.")
            W.update(self.weight_head_3d.table("weight_head_3d."))
            self._Wh = W
        return self._Wh

    # ---- reference API -------------------------------------------------------------------------------
    def proj(self, depth, intrinsics):
        """depth (n,1,H,W), intrinsics (n,3,3) -> (n,3,H,W)   (pose_net.py:121-125)."""
        return ops.proj(depth.float().contiguous(), intrinsics.float().contiguous())

    def flow2depth(self, imagel, imager, baseline, upsample=True):
        """-> depth (n,1,H,W), stereo flow (n,2,H,W), valid (n,1,H,W)   (pose_net.py:127-135)."""
        flow = self.flow(imagel, imager, upsample=upsample)[0][-1]
        n = flow.shape[0]
        K = torch.eye(3, device=flow.device)[None].repeat(n, 1, 1)
        bl = baseline.float().reshape(-1)
        if not upsample:
            bl = bl / 8                 # the 1/8-resolution branch (pose_net.py:130-131): disparities are 8x smaller
        if bl.numel() == 1 and n > 1:
            bl = bl.expand(n)           # the reference broadcasts a (1,) baseline over the batch
        depth, valid, _ = ops.depth_proj(flow.contiguous(), bl.contiguous(), K, None, want_pcl=False)
        return depth, flow, valid

    def get_weight_maps(self, pcl1, pcl2, image1l, image2l, mask2, time_flow, stereo_flow1, stereo_flow2,
                        gru_hidden_state, context, state_planes=None):
        """-> conf1, conf2, pcl2 warped, mask2 warped   (pose_net.py:102-119).  ``state_planes``: (gru, ctx) as the update
        operator's NHWC split planes (batched engine); then ``gru_hidden_state`` / ``context`` may be None."""
        as_img = lambda t: t.contiguous() if t.dtype == torch.uint8 else t.float().contiguous()      # uint8 frames are read in place
        pcl2w, img2w, sflow2w, mask2w = ops.warp8_mask(pcl2, as_img(image2l), stereo_flow2, mask2.bool().contiguous(), time_flow)
        if self.use_weights and self.flow.precision == "fp16x3":
            # both TinyUNets on the tcgen05 convolution kernels (core/unet/unet_tc.py); the 264 / 272-channel inputs are never
            # assembled: the first convolution reads the down-sampled geometry, the GRU state and the context as separate sources
            n = pcl1.shape[0]
            if self._heads_tc is None:
                self._heads_tc = HeadsTC(self._head_weights())
            if state_planes is not None:
                gru_p, ctx_p = state_planes               # the update operator's own planes (first n samples), read in place
            else:
                from ...tc import Planes, nchw_to_planes
                h8, w8 = gru_hidden_state.shape[-2:]
                key = (n, h8, w8, pcl1.device.index)
                if key not in self._head_planes:
                    self._head_planes[key] = (Planes(n, h8, w8, 128, pcl1.device), Planes(n, h8, w8, 128, pcl1.device))
                gru_p, ctx_p = self._head_planes[key]
                nchw_to_planes(gru_hidden_state.float().contiguous(), gru_p)
                nchw_to_planes(context.float().contiguous(), ctx_p)
            conf1, conf2 = self._heads_tc.forward([stereo_flow1.float().contiguous(), as_img(image1l), pcl1],
                                                  [sflow2w, img2w, pcl2w], gru_p, ctx_p, n, self.image_shape)
        elif self.use_weights:
            n = pcl1.shape[0]
            H, W = self.image_shape
            h8, w8 = H // 8, W // 8
            # channel layout of the heads: [down(sflow1,img1,pcl1) 8 | (down(sflow2w,img2w,pcl2w) 8) | gru 128 | ctx 128]
            x3 = torch.empty((n, 272, h8, w8), device=pcl1.device, dtype=torch.float32)
            ops.downsample8_cat([stereo_flow1.float().contiguous(), image1l.float().contiguous(), pcl1], out=x3, ch_offset=0)
            ops.downsample8_cat([sflow2w, img2w, pcl2w], out=x3, ch_offset=8)
            x3[:, 16:144] = gru_hidden_state.float()
            x3[:, 144:272] = context.float()
            x2 = torch.cat((x3[:, :8], x3[:, 16:]), 1)
            W_ = self._head_weights()
            with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False,
                                            allow_tf32=self.flow.precision not in ("fp32", "fp16x3")):
                conf1 = torch.sigmoid(tiny_unet_forward(x2, W_, "weight_head_2d.0.", (H, W)))
                conf2 = torch.sigmoid(tiny_unet_forward(x3, W_, "weight_head_3d.0.", (H, W)))
        else:
            conf1 = torch.ones_like(mask2w, dtype=torch.float32)
            conf2 = torch.ones_like(mask2w, dtype=torch.float32)
        return conf1, conf2, pcl2w, mask2w

    def infer(self, image1l, image2l, intrinsics, baseline, depth1, image2r, mask1, mask2, stereo_flow1,
              ret_details=False):
        """One f2f pair (pose_net.py:60-85).  ``mask2`` is and-ed in place with the stereo validity."""
        with torch.no_grad():
            ref_imgs = torch.cat((image1l, image2l), dim=0)
            trg_imgs = torch.cat((image2l, image2r), dim=0)
            preds, gru, ctx = self.flow(ref_imgs, trg_imgs, upsample=True)
            time_flow = preds[-1][0:1].contiguous()
            stereo_flow2 = preds[-1][1:2].contiguous()
            K = intrinsics.float().contiguous()
            depth2, valid, pcl2 = ops.depth_proj(stereo_flow2, baseline.float().reshape(-1).contiguous(), K, mask2)
            pcl1 = ops.proj(depth1.float().contiguous(), K)
            conf1, conf2, pcl2w, mask2w = self.get_weight_maps(pcl1, pcl2, image1l, image2l, mask2, time_flow,
                                                               stereo_flow1, stereo_flow2, gru[0:1], ctx[0:1])
            pose_vec, pose_tan = self.pose_head(time_flow, pcl1, pcl2w, conf1, conf2, mask1.bool(), mask2w,
                                                K, self.loss_weight[None, :])
        pose_se3 = SE3(pose_vec)
        if ret_details:
            return pose_se3[0], depth1, depth2, [conf1, conf2], time_flow, stereo_flow2
        return pose_se3[0]

    def forward(self, image1l, image2l, intrinsics, baseline, image1r, image2r, mask1=None, mask2=None, ret_confmap=False):
        """Inference forward of pose_net.py:29-58 (three RAFT passes): -> pose tangent (n,6), depth1, depth2[, maps]."""
        with torch.no_grad():
            K = intrinsics.float().contiguous()
            bl = baseline.float().reshape(-1).contiguous()
            sflow1 = self.flow(image1l, image1r)[0][-1]
            sflow2 = self.flow(image2l, image2r)[0][-1]
            n, _, H, W = sflow1.shape
            m1 = torch.ones((n, 1, H, W), dtype=torch.bool, device=K.device) if mask1 is None else mask1.bool().clone()
            m2 = torch.ones((n, 1, H, W), dtype=torch.bool, device=K.device) if mask2 is None else mask2.bool().clone()
            depth1, _, pcl1 = ops.depth_proj(sflow1, bl, K, m1)
            depth2, _, pcl2 = ops.depth_proj(sflow2, bl, K, m2)
            preds, gru, ctx = self.flow(image1l, image2l)
            time_flow = preds[-1]
            conf1, conf2, pcl2w, m2w = self.get_weight_maps(pcl1, pcl2, image1l, image2l, m2, time_flow, sflow1, sflow2, gru, ctx)
            _, tan = self.pose_head(time_flow, pcl1, pcl2w, conf1, conf2, m1, m2w, K,
                                    self.loss_weight.repeat((n, 1)))
        tan = tan.squeeze(1)
        if ret_confmap:
            return tan, depth1, depth2, [conf1, conf2]
        return tan, depth1, depth2

    def init_from_raft(self, raft_ckp):
        sd = torch.load(raft_ckp, map_location="cpu")
        self.flow.load_state_dict(OrderedDict((k.replace("module.", ""), v) for k, v in sd.items()))
        return self

    def freeze_flow(self, freeze=True):
        return self

    def train(self, mode=True):
        return super().train(False)                                # inference-only implementation
