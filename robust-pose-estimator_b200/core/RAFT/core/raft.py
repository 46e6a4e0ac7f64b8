"""RAFT optical flow with the reference's interface (/root/reference/core/RAFT/core/raft.py:24-137).

With ``precision='fp16x3'`` the whole network runs on the hand-written sm_100a kernels: encoders and update
operator on the tcgen05 implicit-GEMM convolution (encoder_tc.py, update_tc.py), correlation volume, lookup and
convex up-sampling (csrc/corr.cu, upsample.cu).  The other precisions keep a torch (cuDNN) trunk built from the
same functional weight table.  Deviations that do not change the consumed result:
  * only the final flow prediction is up-sampled unless ``all_predictions=True`` (the pose path
    reads ``flow_predictions[-1]`` only, pose_net.py:66-67);
  * ``precision``: "fp32" (cuDNN fp32, TF32 off, correlation TF32x3 split) |
    "fp16x3" (parity-grade fast mode: encoders and the 12-iteration update operator on the tcgen05 fp16x3
    convolution kernels, encoder_tc.py / update_tc.py; no cuDNN on the path) | "tf32" | "bf16" / "fp16" (autocast like the reference's CUDA run,
    raft.py:92,100,117)."""
import contextlib

import torch
from torch import nn

from .... import ops
from ...utils.param_tree import build_tree
from .corr import CorrBlock
from .extractor import encoder_entries, encoder_forward
from .update import prepare_update_weights, update_entries, update_forward
from .encoder_tc import EncoderTC, stem_planes
from .update_tc import UpdateTC

_AUTOCAST = {"bf16": torch.bfloat16, "fp16": torch.float16}


def coords_grid(batch, ht, wd, device):
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack((xs, ys), 0).float()[None].repeat(batch, 1, 1, 1)


class RAFT(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.get("small", False):
            raise NotImplementedError("only the 'basic' RAFT of the shipped checkpoints is on the f2f path")
        self.config = config
        config["corr_levels"], config["corr_radius"] = 4, 4          # raft.py:37-38
        self.hidden_dim = self.context_dim = 128
        self.precision = config.get("precision", "fp32")
        if self.precision == "bf16x3":                     # round-1 name of the split mode (the planes were bf16 then)
            self.precision = "fp16x3"
        entries = (encoder_entries("fnet.", "instance", 256) + encoder_entries("cnet.", "batch", 256)
                   + update_entries("update_block."))
        tree = build_tree(entries)
        self.fnet, self.cnet, self.update_block = tree.fnet, tree.cnet, tree.update_block
        self._W = None
        self._tc = None
        self._enc_tc = None
        self._col = {}
        self._feat = {}

    # ---- weight table ------------------------------------------------------------------------
    def invalidate(self):
        """Drop everything derived from the parameters: weight table, packed / BN-folded tensor-core weights, im2col buffers."""
        self._W = self._tc = self._enc_tc = None
        self._col = {}
        self._feat = {}

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate()
        return super().load_state_dict(*a, **k)

    def weights(self):
        if self._W is None:
            W = {}
            for name in ("fnet", "cnet", "update_block"):
                W.update(getattr(self, name).table(name + "."))
            self._W = prepare_update_weights(W, "update_block.")
        return self._W

    def freeze_bn(self):
        return self                                               # BatchNorm is always evaluated in eval mode here

    # ---- building blocks reused by the batched tracker -----------------------------------------
    def _ctx(self):
        """Precision context of the convolutional trunk: fp32 = cuDNN fp32 with TF32 off (parity mode)."""
        stack = contextlib.ExitStack()
        dt = _AUTOCAST.get(self.precision)
        exact = self.precision in ("fp32", "fp16x3")
        stack.enter_context(torch.backends.cudnn.flags(enabled=True, benchmark=not exact, deterministic=False, allow_tf32=not exact))
        stack.enter_context(torch.autocast("cuda", dtype=dt) if dt is not None else torch.autocast("cuda", enabled=False))
        return stack

    def _encoders(self):
        if self._enc_tc is None:
            W = self.weights()
            self._enc_tc = (EncoderTC(W, "fnet.", "instance", [(0, 256, "none")]),
                            EncoderTC(W, "cnet.", "batch", [(0, self.hidden_dim, "tanh"), (self.hidden_dim, 256, "relu")]))
        return self._enc_tc

    def encode(self, limg, rimg):
        """fnet over the left and right images, cnet over the left ones -> (fmap_l, fmap_r, net, inp), NCHW float32.
        On the tensor-core path the stem im2col of the left images is computed once and shared by both encoders."""
        C = limg.shape[0]
        imgs = torch.cat((limg, rimg), 0).float().contiguous()
        if self.precision != "fp16x3":
            f = self.features(imgs)
            net, inp = self.context(limg)
            return f[:C], f[C:], net, inp
        fe, ce = self._encoders()
        key = tuple(imgs.shape) + (imgs.device.index,)
        col = self._col[key] = stem_planes(imgs, self._col.get(key))
        f = fe.forward_nchw(imgs, col)[0]
        net, inp = ce.forward_nchw(imgs[:C], col)
        return f[:C], f[C:], net, inp

    def feature_list(self, key, n, h, w, device):
        """Per-shape image list of feature planes (n, h, w, 256) the feature encoder writes into and the correlation reads.
        Owned by the model, so every engine over it binds the encoder plans to the same buffers."""
        from ....tc import Planes
        buf = self._feat.get(key)
        if buf is None:
            buf = self._feat[key] = Planes(n, h, w, 256, device)
        return buf

    def update_tc(self):
        if self._tc is None:
            self._tc = UpdateTC(self.weights())
        return self._tc

    def encode_into(self, imgs, n_left, feat_dst, h_dst, hp_dst, inp_dst):
        """Tensor-core path of the batched tracker: fnet over ``imgs`` = cat(left, right) (2 n_left images) writes its features as
        split planes into ``feat_dst``; cnet over the n_left left images writes tanh(net) into ``h_dst`` (fp32 NHWC) / ``hp_dst``
        (planes) and relu(inp) into ``inp_dst`` (planes) -- the update operator's own state buffers.  No NCHW round trip."""
        assert self.precision == "fp16x3"
        fe, ce = self._encoders()
        raw = imgs.dtype == torch.uint8                # raw single-plane stem: the normalisation lives in the stem weights
        key = tuple(imgs.shape) + (imgs.device.index, raw)
        col = self._col[key] = stem_planes(imgs, self._col.get(key), raw=raw)
        fe.forward(imgs, col, dests=[{"out_planes": feat_dst}], raw_stem=raw)
        ce.forward(imgs[:n_left], col, dests=[{"out_f32": h_dst, "out_planes": hp_dst}, {"out_planes": inp_dst}], raw_stem=raw)

    def features(self, images):
        """fnet over (N,3,H,W) images in 0..255 -> (N,256,H/8,W/8) float32."""
        if self.precision == "fp16x3":
            return self._encoders()[0].forward_nchw(images.float().contiguous())[0]
        x = (2 * (images / 255.0) - 1.0).contiguous()
        with self._ctx():
            return encoder_forward(x, self.weights(), "fnet.", "instance").float()

    def context(self, images):
        """cnet -> (net, inp) = (tanh, relu) halves, each (N,128,H/8,W/8)."""
        if self.precision == "fp16x3":
            net, inp = self._encoders()[1].forward_nchw(images.float().contiguous())
            return net, inp
        x = (2 * (images / 255.0) - 1.0).contiguous()
        with self._ctx():
            c = encoder_forward(x, self.weights(), "cnet.", "batch")
            net, inp = torch.split(c, [self.hidden_dim, self.context_dim], dim=1)
            return torch.tanh(net), torch.relu(inp)

    def refine(self, fmap1, fmap2, net, inp, iters=12, flow_init=None, upsample=True, all_predictions=False):
        """Correlation pyramid + `iters` GRU updates + convex up-sampling."""
        B, _, h, w = fmap1.shape
        # fp32: 3xTF32 split; fp16x3: the same split on fp16 planes (full-rate kind::f16 MMAs, the trunk's own arithmetic)
        corr_prec = {"fp32": ops.CORR_TF32X3, "fp16x3": ops.CORR_F16X3}.get(self.precision, ops.CORR_TF32)
        if fmap1.shape[1] % 64 != 0 and corr_prec == ops.CORR_F16X3:
            corr_prec = ops.CORR_TF32X3
        corr_fn = CorrBlock(fmap1, fmap2, radius=self.config["corr_radius"], precision=corr_prec)
        if self.precision == "fp16x3" and not all_predictions:
            flow_up, net, flow_lo = self.update_tc().refine(corr_fn._pyr, net, inp, iters, flow_init, want_mask=upsample)
            return [flow_up if upsample else flow_lo], net, inp, flow_lo
        coords0 = coords_grid(B, h, w, fmap1.device)
        coords1 = coords0.clone() if flow_init is None else coords0 + flow_init
        W = self.weights()
        preds = []
        for itr in range(iters):
            corr = corr_fn(coords1)
            flow = coords1 - coords0
            last = itr == iters - 1
            with self._ctx():
                net, up_mask, delta = update_forward(net, inp, corr, flow, W, "update_block.",
                                                     want_mask=upsample and (last or all_predictions))
            coords1 = coords1 + delta.float()
            if last or all_predictions:
                lo = (coords1 - coords0).contiguous()
                preds.append(ops.convex_upsample8(lo, up_mask.float().contiguous()) if upsample else lo)
        return preds, net, inp, coords1 - coords0

    # ---- reference interface -------------------------------------------------------------------
    def forward(self, image1, image2, iters=12, flow_init=None, test_mode=False, upsample=True, all_predictions=False):
        with torch.no_grad():
            B = image1.shape[0]
            fmaps = self.features(torch.cat((image1, image2), 0))
            net, inp = self.context(image1)
            preds, net, inp, flow_lo = self.refine(fmaps[:B].contiguous(), fmaps[B:].contiguous(), net, inp, iters,
                                                   flow_init, upsample, all_predictions)
        if test_mode:
            return flow_lo, preds[-1]
        return preds, net, inp
