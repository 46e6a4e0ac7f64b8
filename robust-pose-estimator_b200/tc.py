"""Plumbing of the tcgen05 convolution plans (csrc/conv.cu, include/rpe_b200.h ``rpe_conv_plan_*``): NHWC fp16 split planes,
weight packing, and a plan object that keeps every buffer it points to alive.  No arithmetic happens here."""
import ctypes as C

import torch

from . import _lib
from .ops import _p, _stream, _timed, check

ACT = {"none": 0, "relu": 1, "sigmoid": 2, "tanh": 3}


PLANE_DTYPE = torch.float16          # both planes fp16: hi = fp16(v), lo = fp16(v - hi)
PLANE_LO_DTYPE = torch.float16
PLANE_MAX = 65504.0


def split_planes(t):
    """fp32 tensor -> (hi, lo) fp16 planes with hi + lo == t to 22 significant bits while |t| >= 2^-3 (an absolute 2^-25 below:
    the lo plane turns subnormal); saturating like the kernels' conversions."""
    hi = t.clamp(-PLANE_MAX, PLANE_MAX).to(PLANE_DTYPE)
    lo = (t - hi.float()).clamp(-PLANE_MAX, PLANE_MAX).to(PLANE_LO_DTYPE)
    return hi.contiguous(), lo.contiguous()


def weight_scale(w):
    """Power of two that lifts a weight tensor (typically O(0.01)) to max |w| in [2^12, 2^13): the fp16 lo plane of all but
    vanishing weights is then a normal number, i.e. the packed weights keep 22 significant bits.  The convolution multiplies its
    accumulators by the inverse (rpe_conv_desc.acc_scale); powers of two make both steps exact."""
    m = float(w.detach().abs().max())
    if not (m > 0.0) or m != m:
        return 1.0
    import math
    return 2.0 ** max(-20, min(24, 12 - math.floor(math.log2(m)) - 1))


def pack_weight(w, c_lo, c_hi, cout_pad, scale=None):
    """(Cout, Cin, kh, kw) fp32 -> the channel window [c_lo, c_hi) as [taps][cout_pad][c_pad64] planes (hi, lo, 1 / scale) of
    scale * w.  ``scale``: a power of two shared by every window of the same convolution (default: weight_scale of this tensor)."""
    cout, _, kh, kw = w.shape
    if scale is None:
        scale = weight_scale(w)
    cs = c_hi - c_lo
    cpad = (cs + 63) // 64 * 64
    out = torch.zeros((kh * kw, cout_pad, cpad), dtype=torch.float32, device=w.device)
    out[:, :cout, :cs] = w[:, c_lo:c_hi].permute(2, 3, 0, 1).reshape(kh * kw, cout, cs) * scale
    hi, lo = split_planes(out)
    return hi, lo, 1.0 / scale


class Planes:
    """NHWC fp16 split planes (hi, lo) of shape (N, H, W, C); channels are zero-initialised so padded channels read as 0."""

    def __init__(self, n, h, w, c, device):
        self.hi = torch.zeros((n, h, w, c), dtype=PLANE_DTYPE, device=device)
        self.lo = torch.zeros((n, h, w, c), dtype=PLANE_LO_DTYPE, device=device)
        self.c = c
        self.shape = (n, h, w, c)

    def float(self):
        return self.hi.float() + self.lo.float()

    def view(self, a, b=None):
        """Planes of the images [a, b) without copying (a contiguous slice along the batch axis)."""
        v = Planes.__new__(Planes)
        v.hi, v.lo, v.c = self.hi[a:b], self.lo[a:b], self.c
        v.shape = tuple(v.hi.shape)
        return v

    def copy_(self, other):
        self.hi.copy_(other.hi)
        self.lo.copy_(other.lo)
        return self

    def clone(self):
        v = Planes.__new__(Planes)
        v.hi, v.lo, v.c, v.shape = self.hi.clone(), self.lo.clone(), self.c, self.shape
        return v


class ConvPlan:
    """One convolution bound to its input / output buffers.  ``inputs``: [(Planes, c_offset, c_count, (w_hi, w_lo))]."""

    def __init__(self, name, inputs, dims, kh, kw, cout, act="none", bias=None, stride=1, out_f32=None, f32_off=0, out_planes=None,
                 bf_off=0, scale=1.0, pre=None, res=None, single_pass=False, mode=0, aux=None, aux2=None, stat_partials=None,
                 act_single=False, res_planes=None, weight_single=False):
        n, h, w = dims
        cout_pad = (cout + 15) // 16 * 16
        d = _lib.ConvDesc()
        self._keep = [bias, pre, res, out_f32, out_planes, aux, aux2, stat_partials, res_planes]
        if stat_partials is not None:               # instance-norm partial sums (conv.cu kind 6), fp32 [N * tiles * 4][cout_pad][2]
            d.stat_partials = stat_partials.data_ptr()
        d.mode = mode
        if aux is not None:
            d.aux, d.aux_ld = aux.data_ptr(), aux.shape[-1]
        if aux2 is not None:
            d.aux2, d.aux2_ld = aux2.data_ptr(), aux2.shape[-1]
        inv_scales = {float(wt[2]) for _, _, _, wt in inputs}
        assert len(inv_scales) == 1, f"{name}: the sources of one convolution must share the weight scale, got {sorted(inv_scales)}"
        d.acc_scale = inv_scales.pop()
        for k, (planes, c_off, c_cnt, (w_hi, w_lo, _)) in enumerate(inputs):
            assert planes.shape[1:3] == (h, w) and planes.shape[0] >= n, f"{name}: source {k} has shape {planes.shape}, expected {(n, h, w)}"
            assert w_hi.shape[1] == cout_pad, f"{name}: weight packed for cout_pad {w_hi.shape[1]}, plan needs {cout_pad}"
            s = d.src[k]
            # act_single: the activations are exact in their hi plane (raw uint8 frames); the weights keep both planes
            s.act_hi, s.act_lo = planes.hi.data_ptr(), (0 if (single_pass or act_single) else planes.lo.data_ptr())
            s.c_total, s.c_offset, s.c_count = planes.c, c_off, (c_cnt + 15) // 16 * 16
            # weight_single: only the hi plane of the weights (two products per multiply-add: a_hi w_hi + a_lo w_hi); for layers whose
            # rounding to 11 weight bits is invisible downstream (the mask head, profiles/r2_precision_study_per_layer.txt)
            s.w_hi, s.w_lo, s.w_cstride = w_hi.data_ptr(), (0 if (single_pass or weight_single) else w_lo.data_ptr()), w_hi.shape[-1]
            self._keep += [planes, w_hi, w_lo]
        d.n_sources = len(inputs)
        d.N, d.H, d.W = n, h, w
        d.kh, d.kw, d.stride, d.cout, d.cout_pad = kh, kw, stride, cout, cout_pad
        d.bias = 0 if bias is None else bias.data_ptr()
        if pre is not None:
            d.pre, d.pre_ld = pre.data_ptr(), pre.shape[-1]
        if res is not None:
            d.res, d.res_ld = res.data_ptr(), res.shape[-1]
        if res_planes is not None:                  # residual read from split planes (may be out_planes: in-place block output)
            assert res is None
            d.res_hi, d.res_lo, d.res_ld = res_planes.hi.data_ptr(), res_planes.lo.data_ptr(), res_planes.c
        d.activation, d.out_scale = ACT[act], scale
        if out_f32 is not None:
            d.out_f32, d.f32_ld, d.f32_offset = out_f32.data_ptr(), out_f32.shape[-1], f32_off
        if out_planes is not None:
            d.out_hi, d.bf_ld, d.bf_offset = out_planes.hi.data_ptr(), out_planes.c, bf_off
            d.out_lo = 0 if single_pass else out_planes.lo.data_ptr()
        self.name = name
        self._h = C.c_void_p()
        check(_lib.lib().rpe_conv_plan_create(C.byref(d), C.byref(self._h)), f"rpe_conv_plan_create({name})")
        self.flops = float(_lib.lib().rpe_conv_plan_flops(self._h))
        self.tiles_per_image = int(_lib.lib().rpe_conv_plan_tiles_per_image(self._h))

    def run(self, stage="conv_tc"):
        with _timed(stage, self.flops):
            check(_lib.lib().rpe_conv_plan_run(self._h, _stream()), f"rpe_conv_plan_run({self.name})")

    def __del__(self):
        try:
            if self._h:
                _lib.lib().rpe_conv_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def nchw_to_planes(x, planes, c_off=0, f32=None, f32_off=0):
    """NCHW fp32 -> split planes at channel offset c_off (+ optional fp32 NHWC copy)."""
    n, c, h, w = x.shape
    check(_lib.lib().rpe_nchw_to_nhwc_split(_p(x), _p(planes.hi), _p(planes.lo), _p(f32), n, c, h, w, planes.c, c_off,
                                            0 if f32 is None else f32.shape[-1], f32_off, _stream()), "rpe_nchw_to_nhwc_split")


def nhwc_to_nchw(x, c, c_off=0):
    """fp32 NHWC (n,h,w,ld) channel window -> NCHW fp32."""
    n, h, w, ld = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.lib().rpe_nhwc_to_nchw(_p(x), _p(out), n, c, h, w, ld, c_off, _stream()), "rpe_nhwc_to_nchw")
    return out
