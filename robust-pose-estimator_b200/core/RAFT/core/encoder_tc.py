"""RAFT feature / context encoders on the tcgen05 convolution kernels (csrc/conv.cu, csrc/encoder_ops.cu).

Same arithmetic graph as ``extractor.encoder_forward`` (reference: /root/reference/core/RAFT/core/extractor.py:118-192
``BasicEncoder``, :6-56 ``ResidualBlock``) with every convolution evaluated as an error-compensated fp16x3 implicit GEMM:
  * the 7x7 / 2 stem is a 1x1 convolution over an im2col of the normalised image (rpe_im2col7s2_split);
  * stride-2 convolutions read their input through strided TMA boxes;
  * fnet (InstanceNorm2d): convolution -> raw fp32 + partial sums, rpe_instnorm_stats_from_partials, rpe_norm_act_split_res
    (normalise / relu / residual);
  * cnet (eval-mode BatchNorm2d): the affine map is folded into the convolution weights and the relu / residual run in the
    convolution epilogue;
  * the residual stream exists as split planes only: block outputs overwrite them in place and the skip connections read them.
All buffers and plans are created once per input shape."""
import os

import torch

from .... import _lib
from ....ops import _p, _stream, _timed, check
from ....tc import ConvPlan, Planes, nhwc_to_nchw, pack_weight

_STAGES = (("layer1", 64, 1), ("layer2", 96, 2), ("layer3", 128, 2))
STEM_K = 168            # 7 filter rows x 24 (21 real taps*channels + 3 zeros)
STEM_LD = 176


def stem_planes(images, out=None, raw=False):
    """im2col of the 7x7 / 2 stem for NCHW images in 0..255 (fp32, or uint8 as the camera delivers them) -> split planes
    (n, H/2, W/2, STEM_LD) of the normalised image.  ``raw`` (uint8 only): ONE plane of the raw pixel values (exact in fp16), for
    the stem plans whose weights carry the normalisation (``EncoderTC(raw_stem=True)``)."""
    import torch
    n, _, H, W = images.shape
    if out is None:
        out = Planes(n, (H - 1) // 2 + 1, (W - 1) // 2 + 1, STEM_LD, images.device)
    u8 = images.dtype == torch.uint8
    assert u8 or not raw, "the raw stem form needs uint8 frames"
    fn = _lib.lib().rpe_im2col7s2_split_u8 if u8 else _lib.lib().rpe_im2col7s2_split
    with _timed("im2col_stem", n):
        check(fn(_p(images), _p(out.hi), _p(None if raw else out.lo), n, H, W, STEM_LD, _stream()), "rpe_im2col7s2_split")
    return out


class EncoderTC:
    def __init__(self, weights, prefix, norm, heads):
        """weights: flat {key: tensor} table; norm 'instance' | 'batch'; heads: [(c_lo, c_hi, activation)] slices of the
        final 1x1 convolution (fnet: one linear head of 256 channels; cnet: tanh / relu halves)."""
        self.W, self.prefix, self.norm, self.heads = weights, prefix, norm, heads
        self._packed = {}
        self._shapes = {}

    # ---- weights ---------------------------------------------------------------------------------------
    def _folded(self, conv, bn):
        """(weight, bias) of `conv` with the eval-mode BatchNorm `bn` folded in (identity for fnet)."""
        W, p = self.W, self.prefix
        w, b = W[p + conv + ".weight"].float(), W[p + conv + ".bias"].float()
        if self.norm == "batch":
            s = W[p + bn + ".weight"].float() / torch.sqrt(W[p + bn + ".running_var"].float() + 1e-5)
            w = w * s[:, None, None, None]
            b = (b - W[p + bn + ".running_mean"].float()) * s + W[p + bn + ".bias"].float()
        return w, b

    def _wb(self, conv, bn, c_lo=None, c_hi=None, raw=False):
        key = (conv, c_lo, c_hi, raw)
        if key not in self._packed:
            w, b = self._folded(conv, bn) if bn else (self.W[self.prefix + conv + ".weight"].float(), self.W[self.prefix + conv + ".bias"].float())
            if raw:                                               # conv(2 v / 255 - 1) = conv'(v): w' = 2 w / 255, b' = b - sum w (raft.py:82-83)
                b = b - w.double().sum((1, 2, 3)).float()
                w = w * (2.0 / 255.0)
            if conv == "conv1":                                   # stem: K = ky*24 + kx*3 + c
                wk = torch.zeros((w.shape[0], 7, 24), dtype=torch.float32, device=w.device)
                wk[:, :, :21] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 7, 21)
                w = wk.reshape(w.shape[0], STEM_K, 1, 1)
            if c_lo is not None:
                w, b = w[c_lo:c_hi], b[c_lo:c_hi]
            cout = w.shape[0]
            self._packed[key] = (pack_weight(w, 0, w.shape[1], (cout + 15) // 16 * 16), b.contiguous())
        return self._packed[key]

    # ---- per-shape state -------------------------------------------------------------------------------
    def _state(self, n, H, W, device, shared_col=None, dests=None, raw_stem=False):
        """dests: per output head a dict(out_f32=fp32 NHWC tensor or None, out_planes=Planes or None) the head convolution writes
        to directly (the batched tracker points them into the correlation / update-operator buffers); default: own fp32 tensors."""
        dkey = None if dests is None else tuple((None if d.get("out_f32") is None else d["out_f32"].data_ptr(),
                                                 None if d.get("out_planes") is None else d["out_planes"].hi.data_ptr()) for d in dests)
        key = (n, H, W, device.index, None if shared_col is None else shared_col.hi.data_ptr(), dkey, raw_stem)
        st = self._shapes.get(key)
        if st is not None:
            return st
        inst = self.norm == "instance"
        f32 = lambda h, w, c: torch.zeros((n, h, w, c), dtype=torch.float32, device=device)
        st = {"steps": [], "keep": []}
        h, w = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        l = _lib.lib()
        ws = torch.empty(l.rpe_instnorm_workspace_bytes(n, 128), dtype=torch.uint8, device=device) if inst else None
        st["keep"].append(ws)
        # instance-norm statistics from partial sums written by the convolution epilogue (conv.cu kind 6) instead of a second
        # pass over the raw tensor; one buffer sized for the largest layer (stem / layer1).  RPE_FUSED_INSTNORM=0: two-pass path.
        fused_stats = inst and os.environ.get("RPE_FUSED_INSTNORM", "1") != "0"
        tiles_of = lambda oh, ow: ((ow + 7) // 8) * ((oh + 15) // 16)
        part = torch.zeros((n * tiles_of(h, w) * 4 * 64 * 2,), dtype=torch.float32, device=device) if fused_stats else None
        st["keep"].append(part)
        pending = {}                                            # raw tensor -> the plan that fills `part` for it

        def stats_of(raw, c, hw):
            s = torch.empty((n, c, 2), dtype=torch.float32, device=device)
            plan = pending.pop(raw.data_ptr(), None)
            if plan is not None:
                st["steps"].append(("stats_tiles", part, s, 4 * plan.tiles_per_image, hw, c, ws))
            else:
                st["steps"].append(("stats", raw, s, hw, c, ws))
            return s

        def norm_act(a, sa, relu_a, b, sb, planes, hw, c, b_planes=None):
            st["steps"].append(("norm", a, sa, relu_a, b, sb, b_planes, planes, hw, c))

        def conv(name, bn, src, dims, k, cout, act, stride=1, out_f32=None, out_planes=None, res_planes=None):
            raw = raw_stem and name == "conv1"
            (wts, bias) = self._wb(name, bn, raw=raw)
            cin = src.c if name != "conv1" else STEM_K
            oh, ow = (dims[1] - 1) // stride + 1, (dims[2] - 1) // stride + 1
            want_stats = (fused_stats and out_f32 is not None and out_planes is None and res_planes is None and act == "none" and cout % 16 == 0
                          and n * tiles_of(oh, ow) * 4 * cout * 2 <= part.numel())
            plan = ConvPlan(self.prefix + name, [(src, 0, min(cin, wts[0].shape[-1]), wts)], dims, k, k, cout, act, bias=bias, stride=stride,
                            out_f32=out_f32, out_planes=out_planes, res_planes=res_planes, stat_partials=part if want_stats else None,
                            act_single=raw)
            if want_stats:
                assert plan.tiles_per_image == tiles_of(oh, ow), (plan.tiles_per_image, oh, ow)
                pending[out_f32.data_ptr()] = plan
            st["steps"].append(("conv", plan))
            return plan

        # The residual stream lives in its split planes only (hi + lo = 22 mantissa bits): block outputs are written as planes, in
        # place, and the skip connection of the next block reads them back -- no fp32 copy of the stream is written or re-read.
        # ---- stem
        col = shared_col if shared_col is not None else Planes(n, h, w, STEM_LD, device)
        st["col"] = col
        xp = Planes(n, h, w, 64, device)
        raw = f32(h, w, 64)
        if inst:
            conv("conv1", None, col, (n, h, w), 1, 64, "none", out_f32=raw)
            s = stats_of(raw, 64, h * w)
            norm_act(raw, s, 1, None, None, xp, h * w, 64)
        else:
            conv("conv1", "norm1", col, (n, h, w), 1, 64, "relu", out_planes=xp)
        cin = 64
        for layer, dim, stride in _STAGES:
            for blk in (0, 1):
                p = f"{layer}.{blk}."
                s_ = stride if blk == 0 else 1
                ih, iw = h, w
                if s_ != 1:
                    h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                hw = h * w
                yp = Planes(n, h, w, dim, device)
                if inst and (s_ != 1 or raw.shape[-1] != dim):
                    raw = f32(h, w, dim)
                xp_in = xp
                if s_ != 1:
                    xp = Planes(n, h, w, dim, device)
                if inst:
                    conv(p + "conv1", None, xp_in, (n, ih, iw), 3, dim, "none", stride=s_, out_f32=raw)
                    s1 = stats_of(raw, dim, hw)
                    norm_act(raw, s1, 1, None, None, yp, hw, dim)
                    raw2 = f32(h, w, dim) if s_ != 1 else raw
                    if s_ != 1:                                     # projection of the skip path first (raw2 is reused below)
                        rawd = f32(h, w, dim)
                        conv(p + "downsample.0", None, xp_in, (n, ih, iw), 1, dim, "none", stride=s_, out_f32=rawd)
                        sd = stats_of(rawd, dim, hw)
                    conv(p + "conv2", None, yp, (n, h, w), 3, dim, "none", out_f32=raw2)
                    s2 = stats_of(raw2, dim, hw)
                    if s_ != 1:
                        norm_act(raw2, s2, 1, rawd, sd, xp, hw, dim)
                    else:
                        norm_act(raw2, s2, 1, None, None, xp, hw, dim, b_planes=xp_in)
                else:
                    res = xp_in
                    if s_ != 1:
                        res = Planes(n, h, w, dim, device)
                        conv(p + "downsample.0", p + "downsample.1", xp_in, (n, ih, iw), 1, dim, "none", stride=s_, out_planes=res)
                    conv(p + "conv1", p + "norm1", xp_in, (n, ih, iw), 3, dim, "relu", stride=s_, out_planes=yp)
                    conv(p + "conv2", p + "norm2", yp, (n, h, w), 3, dim, "relu", out_planes=xp, res_planes=res)
                cin = dim
        # ---- output heads (slices of the final 1x1 convolution)
        outs = []
        for k, (c_lo, c_hi, act) in enumerate(self.heads):
            (wts, bias) = self._wb("conv2", None, c_lo, c_hi)
            if dests is None:
                o = f32(h, w, c_hi - c_lo)
                plan = ConvPlan(self.prefix + "conv2", [(xp, 0, 128, wts)], (n, h, w), 1, 1, c_hi - c_lo, act, bias=bias, out_f32=o)
            else:
                o = None
                plan = ConvPlan(self.prefix + "conv2", [(xp, 0, 128, wts)], (n, h, w), 1, 1, c_hi - c_lo, act, bias=bias,
                                out_f32=dests[k].get("out_f32"), out_planes=dests[k].get("out_planes"))
            st["steps"].append(("conv", plan))
            outs.append(o)
        st["outs"] = outs
        st["out_hw"] = (h, w)
        self._shapes[key] = st
        return st

    # ---- execution ---------------------------------------------------------------------------------------
    def forward(self, images, col=None, dests=None, raw_stem=False):
        """images (n,3,H,W) fp32 in 0..255 -> list of fp32 NHWC outputs, one per head (buffers reused by the next call).
        ``col``: already filled im2col planes from ``stem_planes`` whose first n images are these images (the context encoder
        reads the same left images as the feature encoder); the plans are then bound to that buffer.  ``dests``: see ``_state``."""
        n, _, H, W = images.shape
        st = self._state(n, H, W, images.device, col, dests, raw_stem)
        l = _lib.lib()
        s = _stream()
        if col is None:
            stem_planes(images, st["col"], raw=raw_stem)
        for step in st["steps"]:
            kind = step[0]
            if kind == "conv":
                step[1].run("conv_tc_enc")
            elif kind == "stats_tiles":
                _, part, stats, slots, hw, c, ws = step
                with _timed("instnorm_stats", n):
                    check(l.rpe_instnorm_stats_from_partials(_p(part), _p(stats), n, slots, c, c, hw, 1e-5, _p(ws), ws.numel(), s),
                          "rpe_instnorm_stats_from_partials")
            elif kind == "stats":
                _, raw, stats, hw, c, ws = step
                with _timed("instnorm_stats", n):
                    check(l.rpe_instnorm_stats(_p(raw), _p(stats), n, hw, c, 1e-5, _p(ws), ws.numel(), s), "rpe_instnorm_stats")
            else:
                _, a, sa, relu_a, b, sb, bp, planes, hw, c = step
                with _timed("norm_act", n):
                    check(l.rpe_norm_act_split_res(_p(a), _p(sa), relu_a, _p(b), _p(sb), _p(None if bp is None else bp.hi),
                                                   _p(None if bp is None else bp.lo), 0 if bp is None else bp.c, None, _p(planes.hi),
                                                   _p(planes.lo), planes.c, n, hw, c, s), "rpe_norm_act_split_res")
        return st["outs"]

    def forward_nchw(self, images, col=None):
        return [nhwc_to_nchw(o, o.shape[-1]) for o in self.forward(images, col)]
