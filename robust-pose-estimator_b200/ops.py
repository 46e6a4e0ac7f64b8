"""torch-tensor front end of the C ABI (include/rpe_b200.h).  Pure plumbing: argument checks, output
allocation, raw pointers + the current CUDA stream.  Every function raises ``RpeError`` when the
library is missing or a tensor is not a contiguous CUDA tensor -- there is no fallback path."""
import ctypes as C

import torch

from . import _lib
from ._lib import RpeError, check

__all__ = ["depth_proj", "proj", "warp8_mask", "downsample8_cat", "pose_solve", "PoseSolution", "CorrPyramid",
           "mask_specularities", "SOLVER_LBFGS_REF", "SOLVER_GN", "SOLVER_EVAL_ONLY", "CORR_TF32", "CORR_TF32X3", "CORR_F16X3"]

SOLVER_LBFGS_REF, SOLVER_GN, SOLVER_EVAL_ONLY = _lib.SOLVER_LBFGS_REF, _lib.SOLVER_GN, _lib.SOLVER_EVAL_ONLY
CORR_TF32, CORR_TF32X3, CORR_F16X3 = _lib.CORR_TF32, _lib.CORR_TF32X3, _lib.CORR_F16X3


def _stream():
    """The current stream of the current device.  The package is single-stream per device: scratch workspaces (pose solver
    barriers, per-shape trunk buffers) are shared by all calls on a device, so concurrent calls from several streams or threads
    are not supported (the reference is single-threaded, single-stream as well)."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Optional per-stage CUDA-event timing (bench.py): {stage: [(start_event, end_event, units), ...]}.
_timers = None


def enable_timers(on=True):
    global _timers
    _timers = {} if on else None
    return _timers


class _timed:
    """Brackets one C-ABI call with CUDA events on the launching stream when timers are enabled."""

    def __init__(self, stage, units=1):
        self.stage, self.units = stage, units

    def __enter__(self):
        if _timers is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _timers is not None:
            self.e1.record()
            _timers.setdefault(self.stage, []).append((self.e0, self.e1, self.units))
        return False


def _chk(t, dtype, name, shape=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RpeError(f"{name}: expected a CUDA tensor (rpe_b200 has no CPU path), got "
                       f"{type(t).__name__}{'' if not isinstance(t, torch.Tensor) else ' on ' + str(t.device)}")
    if t.device.index != torch.cuda.current_device():
        raise RpeError(f"{name}: tensor lives on {t.device} but the current device is cuda:{torch.cuda.current_device()} "
                       "(kernels launch on the current device's stream; use torch.cuda.set_device / torch.cuda.device)")
    if t.dtype != dtype:
        raise RpeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RpeError(f"{name}: tensor must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RpeError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------------------------------------
# input pipeline
# ---------------------------------------------------------------------------------------------
def mask_specularities(img, mask=None, spec_thr=0.96, radius=5):
    """rpe_mask_specularities (reference dataset/stereo_dataset.py:12-16): img (n,3,H,W) uint8 RGB on the device, mask
    (n,1,H,W) bool or None -> bool (n,1,H,W): valid & (R+G+B < 3*255*spec_thr), eroded with a (2r+1)^2 box."""
    import math
    n, _, H, W = img.shape
    _chk(img, torch.uint8, "img", (n, 3, H, W))
    if mask is not None:
        _chk(mask, torch.bool, "mask", (n, 1, H, W))
    out = torch.empty((n, 1, H, W), device=img.device, dtype=torch.bool)
    max_sum = int(math.ceil(3 * 255 * spec_thr)) - 1            # integer sums: sum < t  <=>  sum <= ceil(t) - 1
    with _timed("mask_specularities", n):
        check(_lib.lib().rpe_mask_specularities(_p(img), _p(mask), _p(out), n, H, W, max_sum, int(radius), _stream()),
              "rpe_mask_specularities")
    return out


# stage 2
# ---------------------------------------------------------------------------------------------
def depth_proj(stereo_flow, bf, K, mask=None, want_pcl=True):
    """rpe_depth_proj: (depth (n,1,H,W), valid bool (n,1,H,W), pcl (n,3,H,W)); ``mask &= valid`` in place."""
    n, _, H, W = stereo_flow.shape
    _chk(stereo_flow, torch.float32, "stereo_flow", (n, 2, H, W))
    _chk(bf, torch.float32, "baseline", (n,))
    _chk(K, torch.float32, "intrinsics", (n, 3, 3))
    if mask is not None:
        _chk(mask, torch.bool, "mask", (n, 1, H, W))
    depth = torch.empty((n, 1, H, W), device=stereo_flow.device, dtype=torch.float32)
    valid = torch.empty((n, 1, H, W), device=stereo_flow.device, dtype=torch.bool)
    pcl = torch.empty((n, 3, H, W), device=stereo_flow.device, dtype=torch.float32) if want_pcl else None
    with _timed("depth_proj", n):
        check(_lib.lib().rpe_depth_proj(_p(stereo_flow), _p(bf), _p(K), _p(mask), _p(depth), _p(valid), _p(pcl),
                                        n, H, W, _stream()), "rpe_depth_proj")
    return depth, valid, pcl


def proj(depth, K, rescale=None):
    """rpe_proj: pcl (n,3,H,W) = depth * K^-1 [u+.5, v+.5, 1]; ``rescale`` = tracker scale for the fp32 round trip."""
    n, _, H, W = depth.shape
    _chk(depth, torch.float32, "depth", (n, 1, H, W))
    _chk(K, torch.float32, "intrinsics", (n, 3, 3))
    pcl = torch.empty((n, 3, H, W), device=depth.device, dtype=torch.float32)
    with _timed("proj", n):
        check(_lib.lib().rpe_proj(_p(depth), _p(K), _p(pcl), 0 if rescale is None else 1,
                                  1.0 if rescale is None else float(rescale), n, H, W, _stream()), "rpe_proj")
    return pcl


# ---------------------------------------------------------------------------------------------
# stage 5
# ---------------------------------------------------------------------------------------------
def warp8_mask(pcl2, img2, sflow2, mask2, flow):
    """rpe_warp8_mask: returns (pcl2w, img2w, sflow2w, mask2w); any source may be None."""
    n, _, H, W = flow.shape
    _chk(flow, torch.float32, "flow", (n, 2, H, W))
    outs = []
    img_u8 = img2 is not None and img2.dtype == torch.uint8       # camera frames as they are; the warped image is fp32 all the same
    for t, c, name in ((pcl2, 3, "pcl2"), (img2, 3, "img2"), (sflow2, 2, "sflow2")):
        if t is None:
            outs.append(None)
        else:
            _chk(t, torch.uint8 if (name == "img2" and img_u8) else torch.float32, name, (n, c, H, W))
            outs.append(torch.empty(t.shape, dtype=torch.float32, device=t.device))
    m_out = None
    if mask2 is not None:
        _chk(mask2, torch.bool, "mask2", (n, 1, H, W))
        m_out = torch.empty_like(mask2)
    fn = _lib.lib().rpe_warp8_mask_u8 if img_u8 else _lib.lib().rpe_warp8_mask
    with _timed("warp8_mask", n):
        check(fn(_p(pcl2), _p(img2), _p(sflow2), _p(mask2), _p(flow), _p(outs[0]), _p(outs[1]),
                                        _p(outs[2]), _p(m_out), n, H, W, _stream()), "rpe_warp8_mask")
    return outs[0], outs[1], outs[2], m_out


def remap_bilinear(x, flow):
    """rpe_remap_bilinear: generic remap_from_flow (any channel count)."""
    n, Cc, H, W = x.shape
    _chk(x, torch.float32, "x")
    _chk(flow, torch.float32, "flow", (n, 2, H, W))
    out = torch.empty_like(x)
    check(_lib.lib().rpe_remap_bilinear(_p(x), Cc, _p(flow), _p(out), n, H, W, _stream()), "rpe_remap_bilinear")
    return out


def remap_nearest(x, flow):
    """rpe_remap_nearest: generic remap_from_flow_nearest (float in / float out)."""
    n, Cc, H, W = x.shape
    _chk(x, torch.float32, "x")
    _chk(flow, torch.float32, "flow", (n, 2, H, W))
    out = torch.empty_like(x)
    check(_lib.lib().rpe_remap_nearest(_p(x), Cc, _p(flow), _p(out), n, H, W, _stream()), "rpe_remap_nearest")
    return out


def downsample8_cat(srcs, out=None, ch_offset=0):
    """rpe_downsample8_cat: 1/8 bilinear down-sampling of up to 3 NCHW tensors, concatenated on channels."""
    srcs = [s for s in srcs if s is not None]
    if not 1 <= len(srcs) <= 3:
        raise RpeError("downsample8_cat takes 1..3 sources")
    n, _, H, W = srcs[0].shape
    for i, s in enumerate(srcs):
        _chk(s, torch.float32, f"src{i}", (n, s.shape[1], H, W))
    ctot = sum(s.shape[1] for s in srcs)
    if out is None:
        out = torch.empty((n, ctot, H // 8, W // 8), device=srcs[0].device, dtype=torch.float32)
        ch_offset = 0
    _chk(out, torch.float32, "out")
    a = srcs + [None] * (3 - len(srcs))
    ch = [0 if s is None else s.shape[1] for s in a]
    check(_lib.lib().rpe_downsample8_cat(_p(a[0]), ch[0], _p(a[1]), ch[1], _p(a[2]), ch[2], _p(out), ch_offset,
                                         out.shape[1], n, H, W, _stream()), "rpe_downsample8_cat")
    return out


# ---------------------------------------------------------------------------------------------
# stages 3 + 4
# ---------------------------------------------------------------------------------------------
class PoseSolution:
    """Device-resident result of rpe_pose_solve; nothing is copied to the host until asked."""

    def __init__(self, out, pose_f32, log_f32, trace):
        self.out, self.pose, self.log, self.trace = out, pose_f32, log_f32, trace

    @property
    def pose64(self):
        return self.out[:, 0:7]

    @property
    def log64(self):
        return self.out[:, 7:13]

    @property
    def loss(self):
        return self.out[:, 13]

    @property
    def n_evals(self):
        return self.out[:, 16]

    @property
    def status(self):
        return self.out[:, 18]

    @property
    def grad(self):
        return self.out[:, 19:25]

    def hessian(self):
        """(n,6,6) symmetric GN Hessian assembled from the stored upper triangle."""
        n = self.out.shape[0]
        H = torch.zeros((n, 6, 6), dtype=torch.float64, device=self.out.device)
        iu = torch.triu_indices(6, 6, device=self.out.device)
        H[:, iu[0], iu[1]] = self.out[:, 25:46]
        return H + H.transpose(1, 2) - torch.diag_embed(torch.diagonal(H, dim1=1, dim2=2))


_pose_ws = {}


def _pose_workspace(device, n):
    key = (device.index, )
    need = _lib.lib().rpe_pose_workspace_bytes(n)
    ws = _pose_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _pose_ws[key] = ws
    return ws


def pose_solve(flow, pcl1, pcl2, w1, w2, m1, m2, K, lw, mode=SOLVER_LBFGS_REF, max_iter=20, init_pose=None,
               with_hessian=False, trace_cap=0):
    """rpe_pose_solve on n independent pairs.  w1 / w2 may be None (unit confidence)."""
    n, _, H, W = flow.shape
    _chk(flow, torch.float32, "flow", (n, 2, H, W))
    _chk(pcl1, torch.float32, "pcl1", (n, 3, H, W))
    _chk(pcl2, torch.float32, "pcl2", (n, 3, H, W))
    if w1 is not None:
        _chk(w1, torch.float32, "weights1", (n, 1, H, W))
    if w2 is not None:
        _chk(w2, torch.float32, "weights2", (n, 1, H, W))
    _chk(m1, torch.bool, "mask1", (n, 1, H, W))
    _chk(m2, torch.bool, "mask2", (n, 1, H, W))
    _chk(K, torch.float32, "intrinsics", (n, 3, 3))
    _chk(lw, torch.float32, "loss_weight", (n, 2))
    if init_pose is not None:
        _chk(init_pose, torch.float64, "init_pose", (n, 7))
    dev = flow.device
    out = torch.empty((n, _lib.POSE_OUT_STRIDE), dtype=torch.float64, device=dev)      # the kernel writes every field it defines
    pose32 = torch.empty((n, 7), dtype=torch.float32, device=dev)
    log32 = torch.empty((n, 6), dtype=torch.float32, device=dev)
    trace = torch.zeros((n, trace_cap, 16), dtype=torch.float64, device=dev) if trace_cap > 0 else None
    ws = _pose_workspace(dev, n)
    pb = _lib.PoseProblem(flow.data_ptr(), pcl1.data_ptr(), pcl2.data_ptr(), 0 if w1 is None else w1.data_ptr(),
                          0 if w2 is None else w2.data_ptr(), m1.data_ptr(), m2.data_ptr(), K.data_ptr(),
                          lw.data_ptr(), 0 if init_pose is None else init_pose.data_ptr(), n, H, W)
    with _timed("pose_solve", n):
        check(_lib.lib().rpe_pose_solve(C.byref(pb), int(mode), int(max_iter), 1 if with_hessian else 0, _p(out),
                                        _p(pose32), _p(log32), _p(trace), int(trace_cap), _p(ws), ws.numel(), _stream()),
              "rpe_pose_solve")
    return PoseSolution(out, pose32, log32, trace)


# ---------------------------------------------------------------------------------------------
# stage 1
# ---------------------------------------------------------------------------------------------
class CorrPyramid:
    """Device-resident 4-level all-pairs correlation pyramid (rpe_corr_build) + window lookup
    (rpe_corr_lookup).  Every instance owns its volume (like the reference's CorrBlock); the blocks come from torch's
    caching allocator, so a tracker that builds one pyramid per chunk reuses the same memory."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, precision=CORR_TF32):
        B, Cc, h, w = fmap1.shape
        _chk(fmap1, torch.float32, "fmap1", (B, Cc, h, w))
        _chk(fmap2, torch.float32, "fmap2", (B, Cc, h, w))
        self.B, self.C, self.h, self.w = B, Cc, h, w
        self.num_levels, self.radius = num_levels, radius
        l = _lib.lib()
        self.pyramid = torch.empty(l.rpe_corr_pyramid_bytes(B, h, w, num_levels) // 4, dtype=torch.float32, device=fmap1.device)
        self._ws = torch.empty(l.rpe_corr_workspace_bytes(B, Cc, h, w, precision) + 1024, dtype=torch.uint8, device=fmap1.device)
        ws_ptr = (self._ws.data_ptr() + 1023) & ~1023
        with _timed("corr_build", B):
            check(l.rpe_corr_build(_p(fmap1), _p(fmap2), _p(self.pyramid), B, Cc, h, w, num_levels, int(precision),
                                   C.c_void_p(ws_ptr), self._ws.numel() - (ws_ptr - self._ws.data_ptr()), _stream()),
                  "rpe_corr_build")

    @classmethod
    def from_planes(cls, f1, f2, B, num_levels=4, radius=4, f1_wrap=0, f1_sub=0):
        """rpe_corr_build_planes: the volume + pooled pyramid of B samples from feature maps that already are NHWC split planes
        (tc.Planes views over an image list): sample s correlates f1 image (s < f1_wrap ? s : s - f1_sub) with f2 image s."""
        self = cls.__new__(cls)
        _, h, w, Cc = f1.shape
        self.B, self.C, self.h, self.w = B, Cc, h, w
        self.num_levels, self.radius = num_levels, radius
        l = _lib.lib()
        self.pyramid = torch.empty(l.rpe_corr_pyramid_bytes(B, h, w, num_levels) // 4, dtype=torch.float32, device=f1.hi.device)
        self._ws = None
        n_f1 = (B if f1_wrap <= 0 else max(f1_wrap, B - f1_sub))
        if f1.shape[0] < n_f1 or f2.shape[0] < B or tuple(f2.shape[1:]) != tuple(f1.shape[1:]):
            raise RpeError(f"corr planes: f1 {tuple(f1.shape)} / f2 {tuple(f2.shape)} too small for {B} samples (wrap {f1_wrap}, sub {f1_sub})")
        with _timed("corr_build", B):
            check(l.rpe_corr_build_planes(_p(f1.hi), _p(f1.lo), _p(f2.hi), _p(f2.lo), _p(self.pyramid), B, Cc, h, w, num_levels,
                                          int(f1_wrap), int(f1_sub), _stream()), "rpe_corr_build_planes")
        return self

    def level(self, l):
        """View of pyramid level l as (B*h*w, 1, h>>l, w>>l) like the reference's corr_pyramid[l]."""
        off = _lib.lib().rpe_corr_level_offset(self.B, self.h, self.w, l) // 4
        hl, wl = self.h >> l, self.w >> l
        n = self.B * self.h * self.w
        return self.pyramid[off:off + n * hl * wl].view(n, 1, hl, wl)

    def __call__(self, coords):
        _chk(coords, torch.float32, "coords", (self.B, 2, self.h, self.w))
        n = 2 * self.radius + 1
        out = torch.empty((self.B, self.num_levels * n * n, self.h, self.w), dtype=torch.float32, device=coords.device)
        with _timed("corr_lookup", self.B):
            check(_lib.lib().rpe_corr_lookup(_p(self.pyramid), _p(coords), _p(out), self.B, self.h, self.w,
                                             self.num_levels, self.radius, _stream()), "rpe_corr_lookup")
        return out


def corr_lookup_planes(pyr, coords, planes):
    """rpe_corr_lookup_nhwc_split: CorrBlock.__call__ written as NHWC bf16 hi/lo planes (tc.Planes with >= 324 channels)."""
    _chk(coords, torch.float32, "coords", (pyr.B, 2, pyr.h, pyr.w))
    with _timed("corr_lookup", pyr.B):
        check(_lib.lib().rpe_corr_lookup_nhwc_split(_p(pyr.pyramid), _p(coords), _p(planes.hi), _p(planes.lo), planes.c, pyr.B, pyr.h, pyr.w,
                                                   pyr.num_levels, pyr.radius, _stream()), "rpe_corr_lookup_nhwc_split")
    return planes


def convex_upsample8(flow, mask):
    """rpe_convex_upsample8: flow (B,2,h,w), mask (B,576,h,w) -> (B,2,8h,8w)."""
    B, _, h, w = flow.shape
    _chk(flow, torch.float32, "flow", (B, 2, h, w))
    _chk(mask, torch.float32, "mask", (B, 576, h, w))
    out = torch.empty((B, 2, 8 * h, 8 * w), dtype=torch.float32, device=flow.device)
    with _timed("convex_upsample8", B):
        check(_lib.lib().rpe_convex_upsample8(_p(flow), _p(mask), _p(out), B, h, w, _stream()), "rpe_convex_upsample8")
    return out


def conv2d_f16x3(x, weight, bias=None, activation="none", scale=1.0, stride=1, pre=None, res=None, single_pass=False):
    """Convolution ('same'-style padding k//2, stride 1 or 2) of an NCHW fp32 tensor on the tcgen05 implicit-GEMM kernel with
    fp16x3 error compensation (rpe_conv_plan_*).  ``pre`` / ``res``: optional NCHW fp32 addend before the activation / residual
    after it (out = relu(act(v) * scale + res)).  Convenience / test entry: it re-packs weights and activations on every call;
    the RAFT trunk keeps persistent plans instead (core/RAFT/core/update_tc.py, encoder_tc.py)."""
    from . import tc
    n, cin, H, W = x.shape
    cout, _, kh, kw = weight.shape
    _chk(x, torch.float32, "x")
    cpad = (cin + 15) // 16 * 16
    cout_pad = (cout + 15) // 16 * 16
    planes = tc.Planes(n, H, W, cpad, x.device)
    tc.nchw_to_planes(x, planes)
    wts = tc.pack_weight(weight.float(), 0, cin, cout_pad)
    OH, OW = (H + 2 * (kh // 2) - kh) // stride + 1, (W + 2 * (kw // 2) - kw) // stride + 1
    ld = (cout + 3) // 4 * 4
    out = torch.zeros((n, OH, OW, ld), dtype=torch.float32, device=x.device)
    nhwc = lambda t: None if t is None else torch.nn.functional.pad(t.permute(0, 2, 3, 1), (0, ld - cout)).contiguous()
    b = None if bias is None else bias.float().contiguous()
    plan = tc.ConvPlan("conv2d_f16x3", [(planes, 0, cin, wts)], (n, H, W), kh, kw, cout, activation, bias=b, stride=stride,
                       out_f32=out, scale=float(scale), pre=nhwc(pre), res=nhwc(res), single_pass=single_pass)
    plan.run()
    res_t = tc.nhwc_to_nchw(out, cout)
    torch.cuda.current_stream().synchronize()
    return res_t
