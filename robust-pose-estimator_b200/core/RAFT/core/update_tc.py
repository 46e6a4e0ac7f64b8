"""RAFT update operator on the tcgen05 convolution kernels (csrc/conv.cu, csrc/update_ops.cu).

Same arithmetic graph as ``update.update_forward`` (reference: /root/reference/core/RAFT/core/update.py:79-136 and the
loop body of core/RAFT/core/raft.py:112-132) with every convolution evaluated as an error-compensated fp16x3 implicit
GEMM on the tensor cores, NHWC activations, no materialised concatenations and the whole 12-iteration loop resident in
pre-allocated device buffers (one set per batch shape)."""
import os

import torch

from .... import _lib, ops
from ....ops import _p, _stream, _timed, check
from ....tc import ConvPlan, Planes, pack_weight

_Planes, _pack_weight = Planes, pack_weight          # (names used by ops.conv2d_f16x3)


class UpdateTC:
    def __init__(self, weights, prefix="update_block."):
        """weights: flat {key: tensor} table of the RAFT module (fp32, on the device)."""
        self.W = weights
        self.prefix = prefix
        self._packed = {}
        self._shapes = {}
        self.fused_flow_head = os.environ.get("RPE_FUSED_FLOW_HEAD", "1") != "0"     # 0: plain conv1 -> conv2 plans (debug / A-B)

    # ---- weights ---------------------------------------------------------------------------------------
    def _w(self, name, c_lo, c_hi, cout_pad, transform=None):
        key = (name, c_lo, c_hi, cout_pad)
        if key not in self._packed:
            w = self.W[self.prefix + name + ".weight"] if transform is None else transform()
            self._packed[key] = pack_weight(w.float(), c_lo, c_hi, cout_pad)
        return self._packed[key]

    def _bias(self, name, transform=None):
        key = ("bias", name)
        if key not in self._packed:
            b = self.W[self.prefix + name + ".bias"] if transform is None else transform()
            self._packed[key] = b.float().contiguous()
        return self._packed[key]

    # ---- plans -------------------------------------------------------------------------------------------
    def _plan(self, st, name, inputs, kh, kw, cout, act, out_f32=None, f32_off=0, out_planes=None, bf_off=0, scale=1.0,
              weight=None, bias=None, use_bias=True, **kw_extra):
        """inputs: [(planes, c_offset, c_count_real, weight channel window start)]"""
        cout_pad = (cout + 15) // 16 * 16
        srcs = [(planes, c_off, c_cnt, self._w(name, w_lo, w_lo + c_cnt, cout_pad, weight)) for planes, c_off, c_cnt, w_lo in inputs]
        return ConvPlan(name, srcs, st["dims"], kh, kw, cout, act, bias=self._bias(name, bias) if use_bias else None, out_f32=out_f32,
                        f32_off=f32_off, out_planes=out_planes, bf_off=bf_off, scale=scale, **kw_extra)

    def _state(self, n, h, w, device):
        key = (n, h, w, device.index)
        st = self._shapes.get(key)
        if st is not None:
            return st
        P = lambda c: Planes(n, h, w, c, device)
        f32 = lambda c: torch.zeros((n, h, w, c), dtype=torch.float32, device=device)
        st = {"dims": (n, h, w)}
        # GRU input x = cat(inp, motion features, flow) (update.py:96,134): `inp` is constant over the iterations, so its
        # contribution to the six gate convolutions is evaluated once per refinement (pzr*/pq*) and enters as an addend
        st.update(corr=P(384), cor1=P(256), cf=P(256), col=P(128), flo1=P(128), inp=P(128), mot=P(128), hp=P(128), rh=P(128),
                  fh=P(256), mk=P(256), h=f32(128), z=f32(128), pzr1=f32(256), pzr2=f32(256), pq1=f32(128), pq2=f32(128),
                  delta=f32(4), fpart=f32(36), mask=f32(576), coords1=torch.zeros((n, 2, h, w), dtype=torch.float32, device=device))
        W, pre = self.W, self.prefix
        zr_w = lambda half: (lambda: torch.cat((W[pre + f"gru.convz{half}.weight"], W[pre + f"gru.convr{half}.weight"]), 0))
        zr_b = lambda half: (lambda: torch.cat((W[pre + f"gru.convz{half}.bias"], W[pre + f"gru.convr{half}.bias"]), 0))
        f1_w = lambda: W[pre + "encoder.convf1.weight"].permute(0, 2, 3, 1).reshape(128, 98, 1, 1)     # im2col order
        pl = {}
        pl["convc1"] = self._plan(st, "encoder.convc1", [(st["corr"], 0, 324, 0)], 1, 1, 256, "relu", out_planes=st["cor1"])
        pl["convc2"] = self._plan(st, "encoder.convc2", [(st["cor1"], 0, 256, 0)], 3, 3, 192, "relu", out_planes=st["cf"], bf_off=0)
        pl["convf1"] = self._plan(st, "encoder.convf1", [(st["col"], 0, 98, 0)], 1, 1, 128, "relu", out_planes=st["flo1"], weight=f1_w)
        pl["convf2"] = self._plan(st, "encoder.convf2", [(st["flo1"], 0, 128, 0)], 3, 3, 64, "relu", out_planes=st["cf"], bf_off=192)
        pl["conv"] = self._plan(st, "encoder.conv", [(st["cf"], 0, 256, 0)], 3, 3, 126, "relu", out_planes=st["mot"], bf_off=0)
        for half, (kh, kw) in (("1", (1, 5)), ("2", (5, 1))):
            # hx = cat(h, inp, motion, flow): weight input channels [0,128) h, [128,256) inp, [256,384) motion + flow
            pl["pzr" + half] = self._plan(st, "gru.convzr" + half, [(st["inp"], 0, 128, 128)], kh, kw, 256, "none", out_f32=st["pzr" + half],
                                          weight=zr_w(half), bias=zr_b(half))
            pl["pq" + half] = self._plan(st, "gru.convq" + half, [(st["inp"], 0, 128, 128)], kh, kw, 128, "none", out_f32=st["pq" + half])
            pl["zr" + half] = self._plan(st, "gru.convzr" + half, [(st["hp"], 0, 128, 0), (st["mot"], 0, 128, 256)], kh, kw, 256, "sigmoid",
                                         out_f32=st["z"], out_planes=st["rh"], weight=zr_w(half), use_bias=False, pre=st["pzr" + half],
                                         mode=1, aux=st["h"])
            pl["q" + half] = self._plan(st, "gru.convq" + half, [(st["rh"], 0, 128, 0), (st["mot"], 0, 128, 256)], kh, kw, 128, "tanh",
                                        out_planes=st["hp"], use_bias=False, pre=st["pq" + half], mode=2, aux=st["h"], aux2=st["z"])
        if self.fused_flow_head:
            # FlowHead (update.py:6-13): conv2 has 2 output channels -> its per-pixel part (18 dot products of length 256) runs in
            # fp32 inside conv1's epilogue (conv.cu mode 3) and rpe_tap_gather3x3 adds the nine shifted maps; relu(conv1) is never stored
            pkey = ("fh2_proj",)          # (NOT `key`: that is the shape key this state is cached under)
            if pkey not in self._packed:
                self._packed[pkey] = W[pre + "flow_head.conv2.weight"].float().permute(1, 2, 3, 0).reshape(256, 18).contiguous()
            pl["fh1"] = self._plan(st, "flow_head.conv1", [(st["hp"], 0, 128, 0)], 3, 3, 256, "relu", out_f32=st["fpart"], mode=3,
                                   aux2=self._packed[pkey])
        else:
            pl["fh1"] = self._plan(st, "flow_head.conv1", [(st["hp"], 0, 128, 0)], 3, 3, 256, "relu", out_planes=st["fh"])
            pl["fh2"] = self._plan(st, "flow_head.conv2", [(st["fh"], 0, 256, 0)], 3, 3, 2, "none", out_f32=st["delta"])
        # The mask head only shapes the convex up-sampling of the last iteration; the per-layer study (profiles/
        # r2_precision_study_per_layer.txt) puts the effect of ONE weight plane there at 3.5e-7 relative translation / 1e-8 rad,
        # against 1e-3 ... 3e-3 for every other layer group -- so it runs two products per multiply-add (RPE_MASK_HEAD_X3=1: three)
        ws = os.environ.get("RPE_MASK_HEAD_X3", "0") != "1"
        pl["mask0"] = self._plan(st, "mask.0", [(st["hp"], 0, 128, 0)], 3, 3, 256, "relu", out_planes=st["mk"], weight_single=ws)
        pl["mask2"] = self._plan(st, "mask.2", [(st["mk"], 0, 256, 0)], 1, 1, 576, "none", out_f32=st["mask"], scale=0.25, weight_single=ws)
        st["plans"] = pl
        self._shapes[key] = st
        return st

    # ---- execution ---------------------------------------------------------------------------------------
    @staticmethod
    def _run(st, name):
        st["plans"][name].run()

    def refine(self, corr_pyr, net, inp, iters=12, flow_init=None, want_mask=True):
        """corr_pyr: ops.CorrPyramid; net, inp (B,128,h,w) fp32 NCHW.  -> (flow_up or None, net NCHW, flow_lo (B,2,h,w))."""
        B, _, h, w = net.shape
        dev = net.device
        st = self._state(B, h, w, dev)
        l = _lib.lib()
        s = _stream()
        # initial state: h (fp32 + planes), inp planes
        check(l.rpe_nchw_to_nhwc_split(_p(net.float().contiguous()), _p(st["hp"].hi), _p(st["hp"].lo), _p(st["h"]), B, 128, h, w, 128, 0,
                                       128, 0, s), "rpe_nchw_to_nhwc_split")
        check(l.rpe_nchw_to_nhwc_split(_p(inp.float().contiguous()), _p(st["inp"].hi), _p(st["inp"].lo), None, B, 128, h, w, 128, 0, 0, 0, s),
              "rpe_nchw_to_nhwc_split")
        flow_up, flow_lo = self.refine_state(corr_pyr, B, h, w, dev, iters, flow_init, want_mask)
        net_out = torch.empty((B, 128, h, w), dtype=torch.float32, device=dev)
        check(l.rpe_nhwc_to_nchw(_p(st["h"]), _p(net_out), B, 128, h, w, 128, 0, s), "rpe_nhwc_to_nchw")
        return flow_up, net_out, flow_lo

    def state(self, B, h, w, device):
        """The per-shape buffers: 'h' (fp32 NHWC hidden state), 'hp' / 'inp' (split planes of the hidden state / the context
        input) are what the encoders of the batched tracker write into before ``refine_state``."""
        return self._state(B, h, w, device)

    def refine_state(self, corr_pyr, B, h, w, dev, iters=12, flow_init=None, want_mask=True):
        """The refinement loop on state buffers that already hold tanh(net) ('h', 'hp') and relu(inp) ('inp') of the B samples.
        The final hidden state stays in 'h' / 'hp'.  -> (flow_up (B,2,8h,8w) or None, flow_lo (B,2,h,w))."""
        st = self._state(B, h, w, dev)
        l = _lib.lib()
        s = _stream()
        # gate contributions of the (constant) context input, coords1 = grid (+ flow_init)
        for name in ("pzr1", "pq1", "pzr2", "pq2"):
            self._run(st, name)
        if "grid" not in st:
            ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
            st["grid"] = torch.stack((xs, ys), 0).float()
        grid = st["grid"]
        st["coords1"].copy_(grid[None].expand(B, 2, h, w) if flow_init is None else grid[None] + flow_init)
        coords1 = st["coords1"]
        for it in range(iters):
            # coords1 += delta of the previous iteration; flow = coords1 - coords0 -> im2col planes + GRU-input slots
            with _timed("flow_step", B):
                check(l.rpe_flow_step(_p(coords1), _p(st["delta"]) if it > 0 else None, 4, _p(st["col"].hi), _p(st["col"].lo), 128,
                                      _p(st["mot"].hi), _p(st["mot"].lo), 128, 126, B, h, w, s), "rpe_flow_step")
            ops.corr_lookup_planes(corr_pyr, coords1, st["corr"])
            for name in ("convc1", "convc2", "convf1", "convf2", "conv", "zr1", "q1", "zr2", "q2", "fh1"):
                self._run(st, name)
            if self.fused_flow_head:
                with _timed("tap_gather", B):
                    check(l.rpe_tap_gather3x3(_p(st["fpart"]), 36, _p(self._bias("flow_head.conv2")), _p(st["delta"]), 4, B, h, w, s),
                          "rpe_tap_gather3x3")
            else:
                self._run(st, "fh2")
        if iters > 0:                                   # iters = 0: the reference returns flow_init unchanged
            coords1.add_(st["delta"][..., :2].permute(0, 3, 1, 2))
        flow_lo = coords1 - grid[None]
        flow_up = None
        if want_mask:
            self._run(st, "mask0")
            self._run(st, "mask2")
            flow_up = torch.empty((B, 2, 8 * h, 8 * w), dtype=torch.float32, device=dev)
            with _timed("convex_upsample8", B):
                check(l.rpe_convex_upsample8_nhwc(_p(flow_lo.contiguous()), _p(st["mask"]), 576, _p(flow_up), B, h, w, s),
                      "rpe_convex_upsample8_nhwc")
        return flow_up, flow_lo
