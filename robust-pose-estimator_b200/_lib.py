"""ctypes binding of librpe_b200.so (the C ABI in include/rpe_b200.h).  Fails loudly: a missing
library or a failing call raises -- there is no CPU / PyTorch fallback behind these entry points."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "librpe_b200.so")
_lib = None

POSE_OUT_STRIDE = 64
SOLVER_LBFGS_REF, SOLVER_GN, SOLVER_EVAL_ONLY = 0, 1, 2
CORR_TF32, CORR_TF32X3, CORR_F16X3 = 0, 1, 2


class RpeError(RuntimeError):
    pass


class PoseProblem(C.Structure):
    _fields_ = [("flow", C.c_void_p), ("pcl1", C.c_void_p), ("pcl2", C.c_void_p), ("w1", C.c_void_p),
                ("w2", C.c_void_p), ("m1", C.c_void_p), ("m2", C.c_void_p), ("K", C.c_void_p),
                ("lw", C.c_void_p), ("init_pose", C.c_void_p), ("n", C.c_int), ("H", C.c_int), ("W", C.c_int)]


class ConvSource(C.Structure):
    _fields_ = [("act_hi", C.c_void_p), ("act_lo", C.c_void_p), ("c_total", C.c_int), ("c_offset", C.c_int), ("c_count", C.c_int),
                ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("w_cstride", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("n_sources", C.c_int), ("src", ConvSource * 4), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("cout", C.c_int), ("cout_pad", C.c_int), ("bias", C.c_void_p),
                ("pre", C.c_void_p), ("pre_ld", C.c_int), ("res", C.c_void_p), ("res_ld", C.c_int),
                ("activation", C.c_int), ("out_scale", C.c_float), ("out_f32", C.c_void_p), ("f32_ld", C.c_int),
                ("f32_offset", C.c_int), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("bf_ld", C.c_int),
                ("bf_offset", C.c_int), ("mode", C.c_int), ("aux", C.c_void_p), ("aux_ld", C.c_int), ("aux2", C.c_void_p),
                ("aux2_ld", C.c_int), ("stat_partials", C.c_void_p), ("acc_scale", C.c_float),
                ("res_hi", C.c_void_p), ("res_lo", C.c_void_p)]


_P, _I, _F, _Z = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SIGNATURES = {
    # name: (restype, argtypes)
    "rpe_version": (_I, []),
    "rpe_status_string": (C.c_char_p, [_I]),
    "rpe_last_cuda_error": (_I, []),
    "rpe_device_sm_count": (_I, []),
    "rpe_l2_fetch_granularity": (_I, [_I]),
    "rpe_launch_count": (C.c_longlong, []),
    "rpe_mask_specularities": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rpe_depth_proj": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "rpe_proj": (_I, [_P, _P, _P, _I, _F, _I, _I, _I, _P]),
    "rpe_warp8_mask": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "rpe_remap_bilinear": (_I, [_P, _I, _P, _P, _I, _I, _I, _P]),
    "rpe_remap_nearest": (_I, [_P, _I, _P, _P, _I, _I, _I, _P]),
    "rpe_downsample8_cat": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "rpe_pose_workspace_bytes": (_Z, [_I]),
    "rpe_pose_set_groups": (_I, [_I]),
    "rpe_pose_set_group_size": (_I, [_I]),
    "rpe_pose_solve": (_I, [C.POINTER(PoseProblem), _I, _I, _I, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "rpe_compose_trajectory_host": (_I, [_P, _P, _I, _P, _F, _P, _P]),
    "rpe_corr_pyramid_bytes": (_Z, [_I, _I, _I, _I]),
    "rpe_corr_level_offset": (_Z, [_I, _I, _I, _I]),
    "rpe_corr_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "rpe_corr_build": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "rpe_corr_build_planes": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rpe_corr_lookup": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rpe_convex_upsample8": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "rpe_convex_upsample8_nhwc": (_I, [_P, _P, _I, _P, _I, _I, _I, _P]),
    "rpe_conv_plan_create": (_I, [C.POINTER(ConvDesc), C.POINTER(C.c_void_p)]),
    "rpe_conv_plan_run": (_I, [_P, _P]),
    "rpe_conv_plan_flops": (C.c_double, [_P]),
    "rpe_conv_plan_tiles_per_image": (_I, [_P]),
    "rpe_instnorm_stats_from_partials": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _P, _Z, _P]),
    "rpe_conv_plan_destroy": (_I, [_P]),
    "rpe_corr_lookup_nhwc_split": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "rpe_nchw_to_nhwc_split": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rpe_nhwc_to_nchw": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "rpe_flow_step": (_I, [_P, _P, _I, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rpe_im2col7s2_split": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "rpe_instnorm_workspace_bytes": (_Z, [_I, _I]),
    "rpe_instnorm_stats": (_I, [_P, _P, _I, _I, _I, _F, _P, _Z, _P]),
    "rpe_norm_act_split": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "rpe_norm_act_split_res": (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P]),
    "rpe_tap_gather3x3": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "rpe_downsample8_planes": (_I, [_P, _I, _P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rpe_im2col7s2_split_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "rpe_warp8_mask_u8": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "rpe_resize_crop": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rpe_pool2_planes": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rpe_upcat_planes": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "rpe_resize_sigmoid": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    "rpe_gru_gate": (_I, [_P, _P, _P, _P, _P, _I, _I, C.c_longlong, _I, _P]),
}


def lib():
    """The loaded shared library (loads on first use)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RpeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                           "(rpe_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        l = lib()
        msg = l.rpe_status_string(status).decode()
        extra = f" (cudaError {l.rpe_last_cuda_error()})" if status == -4 else ""
        raise RpeError(f"{what} failed: {msg}{extra}")
