"""The two TinyUNet confidence heads on the tcgen05 convolution kernels (csrc/conv.cu, csrc/heads.cu).

Same arithmetic graph as ``unet.tiny_unet_forward`` wrapped with Sigmoid (reference: /root/reference/core/unet/unet.py:8-82,
core/pose/pose_net.py:24-27, 110-115) with every convolution an error-compensated fp16x3 implicit GEMM:

  * an un-padded 3x3 convolution is the interior of the "same" convolution of its input grid, so every tensor keeps its
    grid and carries a VALID REGION (origin, size); pooling / up-sampling kernels hop between grids (rpe_pool2_planes,
    rpe_upcat_planes) and crop the skip connections on the way;
  * eval-mode BatchNorm is folded: into the convolution before it in a DownBlock (conv, BN, ReLU, conv) and into the
    convolution after it in an UpBlock (conv, ReLU, BN, conv: scale on the input channels, shift into the bias -- exact on
    the valid region, where all nine taps see real data);
  * ConvTranspose2d(k=2, s=2) is a 1x1 convolution to 4 x C channels, un-shuffled by rpe_upcat_planes;
  * the first convolution of both heads is ONE convolution with 32 output channels over a source list (1/8 down-sampled
    geometry 16 ch | GRU state 128 | context 128) -- the 264 / 272-channel inputs of pose_net.py:114-115 are never assembled
    (the 2-D head's weights are zero on the 8 channels only the 3-D head reads);
  * rpe_resize_sigmoid evaluates the final bilinear resize + sigmoid in one pass.
All buffers and plans are created once per (batch, grid) shape."""
import torch

from ... import _lib
from ...ops import _p, _stream, _timed, check
from ...tc import ConvPlan, Planes, pack_weight

_ENC = (16, 32, 64)
_DEC = (64, 32, 16)


class HeadsTC:
    def __init__(self, weights):
        """weights: flat {key: tensor} table holding ``weight_head_2d.0.*`` and ``weight_head_3d.0.*`` (fp32, on the device)."""
        self.W = weights
        self._packed = None
        self._shapes = {}

    # ---- weights ---------------------------------------------------------------------------------------
    def _bn(self, name):
        W = self.W
        s = W[name + ".weight"].float() / torch.sqrt(W[name + ".running_var"].float() + 1e-5)
        t = W[name + ".bias"].float() - W[name + ".running_mean"].float() * s
        return s, t

    def _prepare(self):
        if self._packed is not None:
            return self._packed
        W = self.W
        pk = {}
        heads = ("weight_head_2d.0.", "weight_head_3d.0.")
        # ---- first convolution of both heads, BN folded, merged: (32, 272, 3, 3); input channels = [inp1 8 | inp2 8 | gru 128 | ctx 128]
        ws, bs = [], []
        for k, pre in enumerate(heads):
            p = pre + "encoder.enc_blocks.0."
            w, b = W[p + "conv1.weight"].float(), W[p + "conv1.bias"].float()
            s, t = self._bn(p + "norm")
            w, b = w * s[:, None, None, None], b * s + t
            if k == 0:                                   # 2-D head: cat(inp1, gru, ctx) -> zero weights on the inp2 slot
                w = torch.cat((w[:, :8], torch.zeros_like(w[:, :8]), w[:, 8:]), 1)
            ws.append(w), bs.append(b)
        w1 = torch.cat(ws, 0)
        pk["l1"] = ([pack_weight(w1, 0, 16, 32), pack_weight(w1, 16, 144, 32), pack_weight(w1, 144, 272, 32)], torch.cat(bs).contiguous())
        for k, pre in enumerate(heads):
            pk[(k, "enc0.conv2")] = self._plain(pre + "encoder.enc_blocks.0.conv2")
            for i in (1, 2):
                p = f"{pre}encoder.enc_blocks.{i}."
                w, b = W[p + "conv1.weight"].float(), W[p + "conv1.bias"].float()
                s, t = self._bn(p + "norm")
                pk[(k, f"enc{i}.conv1")] = self._pack(w * s[:, None, None, None], b * s + t)
                pk[(k, f"enc{i}.conv2")] = self._plain(p + "conv2")
            for i in (0, 1):
                wt, bt = W[f"{pre}decoder.upconvs.{i}.weight"].float(), W[f"{pre}decoder.upconvs.{i}.bias"].float()
                cin, cout = wt.shape[:2]
                w = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1, 1)          # row (dy * 2 + dx) * cout + co
                pk[(k, f"up{i}")] = self._pack(w, bt.repeat(4))
                p = f"{pre}decoder.dec_blocks.{i}."
                pk[(k, f"dec{i}.conv1")] = self._plain(p + "conv1")
                w2, b2 = W[p + "conv2.weight"].float(), W[p + "conv2.bias"].float()
                s, t = self._bn(p + "norm")                                      # conv2(s * relu + t)
                pk[(k, f"dec{i}.conv2")] = self._pack(w2 * s[None, :, None, None], b2 + (w2 * t[None, :, None, None]).sum((1, 2, 3)))
            pk[(k, "head")] = self._plain(pre + "head")
        self._packed = pk
        return pk

    def _pack(self, w, b):
        cout = w.shape[0]
        return pack_weight(w.contiguous(), 0, w.shape[1], (cout + 15) // 16 * 16), b.contiguous()

    def _plain(self, name):
        return self._pack(self.W[name + ".weight"].float(), self.W[name + ".bias"].float())

    # ---- per-shape state -------------------------------------------------------------------------------
    def _state(self, n, h, w, device, gru, ctx):
        """gru, ctx: Planes (>= n samples, h, w, 128) the first convolution reads in place."""
        key = (n, h, w, device.index, gru.hi.data_ptr(), ctx.hi.data_ptr())
        st = self._shapes.get(key)
        if st is not None:
            return st
        pk = self._prepare()
        f32 = lambda hh, ww, c: torch.zeros((n, hh, ww, c), dtype=torch.float32, device=device)
        P = lambda hh, ww, c: Planes(n, hh, ww, c, device)
        st = {"ds": P(h, w, 16), "steps": [], "logits": []}
        steps = st["steps"]

        def conv(srcs, dims, k, cout, act, wb, out_f32=None, out_planes=None):
            wts, bias = wb
            wts = wts if isinstance(wts, list) else [wts]
            plan = ConvPlan("heads", [(pl, off, cnt, wt) for (pl, off, cnt), wt in zip(srcs, wts)], dims, k, k, cout, act, bias=bias,
                            out_f32=out_f32, out_planes=out_planes)
            steps.append(("conv", plan))

        a1 = P(h, w, 32)
        conv([(st["ds"], 0, 16), (gru, 0, 128), (ctx, 0, 128)], (n, h, w), 3, 32, "relu", pk["l1"], out_planes=a1)
        for k in (0, 1):
            # ---- encoder: grid g0 = (h, w), valid region of enc0 = origin 2, size (h - 4, w - 4)
            e0 = f32(h, w, 16)
            conv([(a1, 16 * k, 16)], (n, h, w), 3, 16, "none", pk[(k, "enc0.conv2")], out_f32=e0)
            h1, w1 = (h - 4) // 2, (w - 4) // 2
            p0 = P(h1, w1, 16)
            steps.append(("pool", e0, (h, w, 16, 0, 2, 2), p0, (h1, w1, 16)))
            b1 = P(h1, w1, 32)
            conv([(p0, 0, 16)], (n, h1, w1), 3, 32, "relu", pk[(k, "enc1.conv1")], out_planes=b1)
            e1 = f32(h1, w1, 32)
            conv([(b1, 0, 32)], (n, h1, w1), 3, 32, "none", pk[(k, "enc1.conv2")], out_f32=e1)
            h2, w2 = (h1 - 4) // 2, (w1 - 4) // 2
            p1 = P(h2, w2, 32)
            steps.append(("pool", e1, (h1, w1, 32, 0, 2, 2), p1, (h2, w2, 32)))
            c1 = P(h2, w2, 64)
            conv([(p1, 0, 32)], (n, h2, w2), 3, 64, "relu", pk[(k, "enc2.conv1")], out_planes=c1)
            e2p = P(h2, w2, 64)
            conv([(c1, 0, 64)], (n, h2, w2), 3, 64, "none", pk[(k, "enc2.conv2")], out_planes=e2p)      # valid origin 2, size (h2 - 4, w2 - 4)
            # ---- decoder 0: up-convolution of e2's valid region, cat with the centre crop of e1's valid region (h1 - 4, w1 - 4)
            u0 = f32(h2, w2, 128)
            conv([(e2p, 0, 64)], (n, h2, w2), 1, 128, "none", pk[(k, "up0")], out_f32=u0)
            hd0, wd0 = 2 * (h2 - 4), 2 * (w2 - 4)
            dh, dw = ((h1 - 4) - hd0) // 2, ((w1 - 4) - wd0) // 2
            d0in = P(hd0, wd0, 64)
            steps.append(("upcat", u0, (h2, w2, 128, 2, 2, 32), e1, (h1, w1, 32, 0, 2 + dh, 2 + dw, 32), d0in, (hd0, wd0, 64)))
            d0a = P(hd0, wd0, 32)
            conv([(d0in, 0, 64)], (n, hd0, wd0), 3, 32, "relu", pk[(k, "dec0.conv1")], out_planes=d0a)
            d0p = P(hd0, wd0, 32)
            conv([(d0a, 0, 32)], (n, hd0, wd0), 3, 32, "none", pk[(k, "dec0.conv2")], out_planes=d0p)     # valid origin 2, size (hd0 - 4, wd0 - 4)
            # ---- decoder 1
            u1 = f32(hd0, wd0, 64)
            conv([(d0p, 0, 32)], (n, hd0, wd0), 1, 64, "none", pk[(k, "up1")], out_f32=u1)
            hd1, wd1 = 2 * (hd0 - 4), 2 * (wd0 - 4)
            dh, dw = ((h - 4) - hd1) // 2, ((w - 4) - wd1) // 2
            d1in = P(hd1, wd1, 32)
            steps.append(("upcat", u1, (hd0, wd0, 64, 2, 2, 16), e0, (h, w, 16, 0, 2 + dh, 2 + dw, 16), d1in, (hd1, wd1, 32)))
            d1a = P(hd1, wd1, 16)
            conv([(d1in, 0, 32)], (n, hd1, wd1), 3, 16, "relu", pk[(k, "dec1.conv1")], out_planes=d1a)
            d1p = P(hd1, wd1, 16)
            conv([(d1a, 0, 16)], (n, hd1, wd1), 3, 16, "none", pk[(k, "dec1.conv2")], out_planes=d1p)     # valid origin 2, size (hd1 - 4, wd1 - 4)
            lg = f32(hd1, wd1, 4)
            conv([(d1p, 0, 16)], (n, hd1, wd1), 1, 1, "none", pk[(k, "head")], out_f32=lg)
            st["logits"].append((lg, (hd1, wd1, 4, 0, 2, 2, hd1 - 4, wd1 - 4)))
        self._shapes[key] = st
        return st

    # ---- execution ---------------------------------------------------------------------------------------
    def forward(self, inp1_srcs, inp2_srcs, gru, ctx, n, out_size):
        """inp1_srcs / inp2_srcs: lists of NCHW fp32 tensors (n, c, H, W) whose 1/8 down-samplings form the 8 + 8 geometry
        channels (stereo flow, image, point cloud of frame 1 / of the warped frame 2); gru, ctx: Planes of the GRU state and the
        context features at 1/8 resolution (first n samples).  -> conf1, conf2 (n,1,H,W) fp32."""
        H, W = out_size
        h, w = H // 8, W // 8
        dev = gru.hi.device
        st = self._state(n, h, w, dev, gru, ctx)
        l = _lib.lib()
        s = _stream()
        with _timed("heads", n):
            for k, srcs in enumerate((inp1_srcs, inp2_srcs)):
                a = list(srcs) + [None] * (3 - len(srcs))
                ch = [0 if t is None else t.shape[1] for t in a]
                u8 = sum(1 << i for i, t in enumerate(a) if t is not None and t.dtype == torch.uint8)      # uint8 frames are read as they are
                check(l.rpe_downsample8_planes(_p(a[0]), ch[0], _p(a[1]), ch[1], _p(a[2]), ch[2], u8, _p(st["ds"].hi), _p(st["ds"].lo), 16,
                                               8 * k, n, H, W, s), "rpe_downsample8_planes")
            for step in st["steps"]:
                if step[0] == "conv":
                    step[1].run("conv_tc_heads")
                elif step[0] == "pool":
                    _, x, (hh, ww, ld, coff, y0, x0), out, (oh, ow, c) = step
                    check(l.rpe_pool2_planes(_p(x), hh, ww, ld, coff, y0, x0, _p(out.hi), _p(out.lo), oh, ow, out.c, c, n, s), "rpe_pool2_planes")
                else:
                    _, up, (hu, wu, ldu, uy0, ux0, cup), skip, (hk, wk, ldk, koff, ky0, kx0, cskip), out, (oh, ow, c) = step
                    check(l.rpe_upcat_planes(_p(up), hu, wu, ldu, uy0, ux0, cup, _p(skip), hk, wk, ldk, koff, ky0, kx0, cskip, _p(out.hi),
                                             _p(out.lo), oh, ow, out.c, n, s), "rpe_upcat_planes")
            confs = []
            for lg, (hl, wl, ld, ch_, y0, x0, ih, iw) in st["logits"]:
                conf = torch.empty((n, 1, H, W), dtype=torch.float32, device=dev)
                check(l.rpe_resize_sigmoid(_p(lg), hl, wl, ld, ch_, y0, x0, ih, iw, _p(conf), n, H, W, s), "rpe_resize_sigmoid")
                confs.append(conf)
        return confs[0], confs[1]
