/*
 * rpe_b200 -- C ABI of the B200-native (sm_100a) per-frame frame-to-frame pose path of
 * aimi-lab/robust-pose-estimator.
 *
 * The reference has NO C ABI / FFI / plugin registry on this path: it is a
 * pure-Python API over torch + lietorch.  Each entry point below therefore cites the reference
 * Python callable it replaces.  Conventions for every call:
 *   - all pointers are DEVICE pointers unless the name ends in _host; tensors are dense, row-major,
 *     laid out exactly like the reference's torch tensors (NCHW planar fp32, bool masks as 1 byte);
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; every call is asynchronous on it;
 *   - no hidden allocation: workspaces are caller-provided, sized by the *_workspace_bytes queries;
 *   - the return value is 0 (RPE_OK) or a negative rpe_status; nothing throws;
 *   - numerical failure (NaN pose, non-convergence) is NOT an error, exactly as in the reference
 *     (core/pose/pose_estimator.py:81-85 handles it on the host).
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 */
#ifndef RPE_B200_H
#define RPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rpe_status {
    RPE_OK = 0,
    RPE_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, unsupported shape        */
    RPE_ERR_ALIGNMENT = -2,     /* pointer / stride not aligned as the entry point requires  */
    RPE_ERR_WORKSPACE = -3,     /* workspace too small                                        */
    RPE_ERR_CUDA = -4,          /* a CUDA runtime / driver call failed (see rpe_last_cuda_error) */
    RPE_ERR_UNSUPPORTED_DEVICE = -5 /* not an sm_100 device                                    */
} rpe_status;

/* Library / device introspection. */
int rpe_version(void);
const char *rpe_status_string(int status);
int rpe_last_cuda_error(void);                 /* cudaError_t of the last failing CUDA call      */
int rpe_device_sm_count(void);
/* cudaLimitMaxL2FetchGranularity of the current device: sets it when bytes > 0 (32, 64 or 128 -- a hint), returns the value in force
 * (or a negative status).  The windowed gathers of the correlation lookup touch 40-byte rows; see DESIGN.md for the measurement. */
int rpe_l2_fetch_granularity(int bytes);
long long rpe_launch_count(void);               /* kernels launched by this library so far        */

/* ------------------------------------------------------------------------------------------------
 * Input pipeline ("next" row 8f-4 of SURVEY.md): the specularity mask of the reference's datasets.
 * Replaces mask_specularities (/root/reference/dataset/stereo_dataset.py:12-16):
 *   spec = img.sum(-1) < 3*255*spec_thr ; mask &= spec ; mask = cv2.erode(mask, ones((2r+1, 2r+1)))   (r = 5)
 * img (n,3,H,W) u8 planar RGB on the device, mask_in (n,1,H,W) u8 or NULL (all valid), mask_out (n,1,H,W) u8 (0/1,
 * must not alias mask_in).  max_sum = ceil(3*255*spec_thr) - 1 (= 734 for the reference's 0.96): the channel sum is an
 * integer, so `sum < t` is `sum <= max_sum`.  Out-of-image pixels do not constrain the erosion (cv2's default border).
 * ---------------------------------------------------------------------------------------------- */
int rpe_mask_specularities(const uint8_t *img, const uint8_t *mask_in, uint8_t *mask_out, int n, int H, int W, int max_sum,
                           int radius, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2 -- stereo depth lifting and pinhole back-projection (coalesced per-pixel kernels).
 * ---------------------------------------------------------------------------------------------- */

/* Replaces PoseNet.infer's depth block + PoseNet.proj
 *   (/root/reference/core/pose/pose_net.py:73-79, 121-125; flow2depth :127-135;
 *    reproject: core/geometry/pinhole_transforms.py:79-87, create_img_coords_t :7-19).
 *   depth = bf / -stereo_flow.x ; valid = (depth > 0) & (depth <= 1) ; depth[~valid] = 1 ;
 *   mask &= valid (in place, only if mask_inout != NULL) ; pcl = depth * K^-1 [u+.5, v+.5, 1].
 * stereo_flow (n,2,H,W) f32, bf (n) f32, K (n,3,3) f32, mask_inout (n,1,H,W) u8 or NULL,
 * depth (n,1,H,W) f32 out, valid (n,1,H,W) u8 out (may be NULL), pcl (n,3,H,W) f32 out (may be NULL). */
int rpe_depth_proj(const float *stereo_flow, const float *bf, const float *K, uint8_t *mask_inout,
                   float *depth, uint8_t *valid, float *pcl, int n, int H, int W, void *stream);

/* Replaces PoseNet.proj (/root/reference/core/pose/pose_net.py:121-125) for a given depth map.
 * The depth fed to the projection is ((depth * pre_div_recip_a) * pre_mul_b) evaluated in fp32 with
 * one rounding per operation when `rescale` != 0 -- this reproduces the tracker's
 * `frame.depth = depth / scale` ... `depth1 = frame.depth * scale` round trip
 * (core/pose/pose_estimator.py:115,121): pass rescale=1, scale=the tracker's fp32 scale.
 * depth (n,1,H,W) f32, K (n,3,3) f32, pcl (n,3,H,W) f32 out. */
int rpe_proj(const float *depth, const float *K, float *pcl, int rescale, float scale,
             int n, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 5 -- flow warping.
 * ---------------------------------------------------------------------------------------------- */

/* Replaces the three remap_from_flow calls + remap_from_flow_nearest + mask combination of
 * PoseNet.get_weight_maps (/root/reference/core/pose/pose_net.py:104-108;
 * core/interpol/flow_utils.py:4-26): bilinear backward warp (align_corners=True, zeros padding) of
 * pcl2 (n,3,H,W), img2 (n,3,H,W), sflow2 (n,2,H,W) by `flow` (n,2,H,W), and nearest warp of
 * mask2 (n,1,H,W) u8 with  mask2w = (warped > 0) & warped.  Sampling coordinates follow the
 * reference's fp32 operation order exactly so that mask2w is bit-exact.
 * Any of the (src, dst) pairs may be NULL to skip that tensor. */
int rpe_warp8_mask_u8(const float *pcl2, const unsigned char *img2, const float *sflow2, const uint8_t *mask2,
                      const float *flow, float *pcl2w, float *img2w, float *sflow2w, uint8_t *mask2w, int n,
                      int H, int W, void *stream);      /* img2 as uint8 (n,3,H,W); img2w stays fp32 */
int rpe_warp8_mask(const float *pcl2, const float *img2, const float *sflow2, const uint8_t *mask2,
                   const float *flow, float *pcl2w, float *img2w, float *sflow2w, uint8_t *mask2w,
                   int n, int H, int W, void *stream);

/* Operator-level seams with the reference's generic signatures
 * (/root/reference/core/interpol/flow_utils.py:4-14 remap_from_flow, :17-26 remap_from_flow_nearest):
 * x (n,C,H,W) f32 warped by flow (n,2,H,W) -> out (n,C,H,W); zeros outside the image. */
int rpe_remap_bilinear(const float *x, int C, const float *flow, float *out, int n, int H, int W, void *stream);
int rpe_remap_nearest(const float *x, int C, const float *flow, float *out, int n, int H, int W, void *stream);

/* Replaces F.interpolate(cat(stereo_flow, image, pcl), scale_factor=0.125, mode='bilinear')
 * (/root/reference/core/pose/pose_net.py:110-113): out[c, i, j] = mean of the 2x2 pixels
 * (8i+3..8i+4, 8j+3..8j+4).  Up to three sources concatenated along channels:
 * src_k (n,c_k,H,W) -> out (n, c_0+c_1+c_2, H/8, W/8) written at channel offset `out_ch_offset` of
 * a tensor with `out_ch_total` channels. */
int rpe_downsample8_cat(const float *src0, int c0, const float *src1, int c1, const float *src2, int c2,
                        float *out, int out_ch_offset, int out_ch_total, int n, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stages 3 + 4 -- fused residual / Jacobian reduction and the on-device SE(3) solver.
 * ---------------------------------------------------------------------------------------------- */

typedef enum rpe_solver_mode {
    RPE_SOLVER_LBFGS_REF = 0,  /* exact replica of torch.optim.LBFGS as driven by DPoseSE3Head.solve */
    RPE_SOLVER_GN = 1,         /* Gauss-Newton: in-kernel 6x6 Cholesky + left exp-map update         */
    RPE_SOLVER_EVAL_ONLY = 2   /* one evaluation of f, grad (and GN Hessian) at the given poses       */
} rpe_solver_mode;

/* Per-pair result record written by rpe_pose_solve (doubles). */
#define RPE_POSE_OUT_STRIDE 64
/*  [0..6]   pose  tx ty tz qx qy qz qw            (LieGroupParameter.group.vec())
 *  [7..12]  log   tau(3) phi(3)                   (LieGroupParameter.log())
 *  [13]     f at the last evaluation   [14] L2d   [15] L3d
 *  [16]     number of objective evaluations       [17] number of solver iterations
 *  [18]     status: 0 ok, 1 Cholesky failed (GN)
 *  [19..24] gradient (unclipped) at the last evaluation
 *  [25..45] GN Hessian upper triangle, row-major (GN / EVAL_ONLY with hessian)
 */

typedef struct rpe_pose_problem {
    const float *flow;      /* (n,2,H,W)  time flow                                                */
    const float *pcl1;      /* (n,3,H,W)                                                           */
    const float *pcl2;      /* (n,3,H,W)  warped target cloud                                      */
    const float *w1;        /* (n,1,H,W)  2D confidence, or NULL for ones                          */
    const float *w2;        /* (n,1,H,W)  3D confidence, or NULL for ones                          */
    const uint8_t *m1;      /* (n,1,H,W)  bool                                                     */
    const uint8_t *m2;      /* (n,1,H,W)  bool                                                     */
    const float *K;         /* (n,3,3)                                                             */
    const float *lw;        /* (n,2) [w3d, w2d]                                                    */
    const double *init_pose;/* (n,7) starting poses or NULL for identity                           */
    int n, H, W;
} rpe_pose_problem;

/* Replaces DPoseSE3Head.solve + objective (/root/reference/core/pose/pose_head.py:12-79),
 * project/transform (core/geometry/pinhole_transforms.py:28-30,90-99), lietorch SE3 exp/mul/act/log,
 * torch.optim.LBFGS.step, clip_grad_norm_(y, 10) and DeclarativeFunctionLie.forward
 * (core/optimization/declerative_node_lie.py:224-247).  Pairs are solved INDEPENDENTLY
 * (the reference is batch-1).  One persistent cooperative kernel: per evaluation every
 * pixel's 2D/3D residuals, Jacobians and weights are fused and reduced in fp64; the 6-dim solver state
 * lives on the device -- no host synchronisation.
 *   out      (n, RPE_POSE_OUT_STRIDE) f64
 *   pose_f32 (n,7) f32 and log_f32 (n,6) f32: the `.float()` outputs of the declarative layer (may be NULL)
 *   trace    (n, trace_cap, 16) f64 or NULL: per evaluation [pose7 | grad6 | f | L2d | L3d]
 *   max_iter = lbgfs_iters (LBFGS) or GN iterations. */
size_t rpe_pose_workspace_bytes(int n_pairs);
/* Tuning (results never depend on it: the reduction tree is fixed by the pixel index).  rpe_pose_set_groups: upper bound on the
 * pairs solved concurrently by disjoint CTA groups (1..64; default: as many as fit).  rpe_pose_set_group_size: CTAs per group when
 * a batch is solved (power of two <= 128, default 16).  Process-wide settings, not thread-safe -- like the rest of the library's
 * state (cached function attributes, last-error code): the reference contract is one host thread per process. */
int rpe_pose_set_groups(int groups);
int rpe_pose_set_group_size(int ctas);
int rpe_pose_solve(const rpe_pose_problem *problem_host, int mode, int max_iter, int with_hessian,
                   double *out, float *pose_f32, float *log_f32, double *trace, int trace_cap,
                   void *workspace, size_t workspace_bytes, void *stream);

/* Host-side (CPU pointers) trajectory composition -- replaces the per-frame tail of PoseEstimator.forward
 * (/root/reference/core/pose/pose_estimator.py:81-91): failure guard (NaN or |log| > 0.1 -> identity),
 * rel.scale(1/scale), last_pose <- last_pose * rel^-1, in fp32.
 *   rel_host (n,7), log_host (n,6): per-pair solver outputs; init_pose_host (7);
 *   abs_out_host (n+1,7): init pose followed by the pose after every pair; failed_out_host (n) or NULL. */
int rpe_compose_trajectory_host(const float *rel_host, const float *log_host, int n, const float *init_pose_host,
                                float inv_scale, float *abs_out_host, unsigned char *failed_out_host);

/* Input pipeline -- replaces ResizeStereo (/root/reference/dataset/transforms.py:20-39: torchvision resize that
 * conserves the aspect ratio, then centre crop) on frames that are already on the device.
 *   src (n,C,Hi,Wi) uint8 (src_u8 = 1) or fp32; dst (n,C,H,W) fp32 = crop(resize(src, (rh, rw)))[top : top + H, left : left + W]
 *   mode 0: bilinear, align_corners = False, anti-aliased when down-scaling (ATen _upsample_bilinear2d_aa: triangle filter of
 *           support max(scale, 1), weights normalised, rows first then columns) -- images
 *   mode 1: nearest (ATen upsample_nearest2d: src = floor(dst * in / out)) -- masks; dst is then uint8 (n,C,H,W) */
int rpe_resize_crop(const void *src, int src_u8, void *dst, int n, int C, int Hi, int Wi, int rh, int rw, int top, int left, int H,
                    int W, int mode, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 1 -- RAFT CorrBlock: all-pairs correlation GEMM (tcgen05 + TMA), pyramid, radius lookup.
 * ---------------------------------------------------------------------------------------------- */

typedef enum rpe_corr_precision {
    RPE_CORR_TF32 = 0,     /* one tcgen05 kind::tf32 pass (inputs rounded to tf32)                  */
    RPE_CORR_TF32X3 = 1,   /* split hi/lo: hi*hi + lo*hi + hi*lo, fp32-class accuracy               */
    RPE_CORR_F16X3 = 2     /* same split on fp16 planes, full-rate kind::f16 MMAs (22 mantissa bits
                              per operand, the arithmetic of the fp16x3 convolution trunk)          */
} rpe_corr_precision;

/* Replaces CorrBlock.__init__ / CorrBlock.corr (/root/reference/core/RAFT/core/corr.py:13-27, 52-60):
 * level0[b, q, t] = <fmap1[b,:,q], fmap2[b,:,t]> / sqrt(C) and 3 further 2x2 average-pooled levels.
 *   fmap1, fmap2 (B,C,h,w) f32 (C % 32 == 0)
 *   pyramid      level l at byte offset given by rpe_corr_level_offset: (B*h*w, h>>l, w>>l) f32
 *   workspace    rpe_corr_workspace_bytes: K-major tf32 operand copies */
size_t rpe_corr_pyramid_bytes(int B, int h, int w, int num_levels);
size_t rpe_corr_level_offset(int B, int h, int w, int level);
size_t rpe_corr_workspace_bytes(int B, int C, int h, int w, int precision);
int rpe_corr_build(const float *fmap1, const float *fmap2, float *pyramid, int B, int C, int h, int w,
                   int num_levels, int precision, void *workspace, size_t workspace_bytes, void *stream);
/* The RPE_CORR_F16X3 build from feature maps that already are NHWC fp16 split planes (B,h,w,C), C % 64 == 0 -- what the
 * feature encoder's last convolution writes: one tcgen05 kernel (CTA pairs, M = 256 queries x N = 16x16 target block) whose
 * epilogue emits level 0 AND the 2x2 / 4x4 / 8x8 mean-pooled levels from the accumulators (corr.py:25-27 avg_pool2d order). */
int rpe_corr_build_planes(const void *f1_hi, const void *f1_lo, const void *f2_hi, const void *f2_lo, float *pyramid, int B, int C,
                          int h, int w, int num_levels, int f1_wrap, int f1_sub, void *stream);
/*   Sample s correlates f1 image (s < f1_wrap ? s : s - f1_sub) with f2 image s (f1_wrap <= 0: identity).  The batched tracker
 *   keeps the features of a chunk as one image list [previous left | left 0..C-1 | right 0..C-1]: with f2 = f1 + one image the
 *   C temporal pairs (k-1 -> k) and the C stereo pairs (left k -> right k) are the 2C samples of ONE launch (wrap C, sub C-1). */

/* Replaces CorrBlock.__call__ + bilinear_sampler (/root/reference/core/RAFT/core/corr.py:29-50,
 * core/RAFT/core/utils/utils.py:57-71): coords (B,2,h,w) f32 (channel 0 = x) ->
 * out (B, num_levels*(2r+1)^2, h, w) f32, channel = level*(2r+1)^2 + i*(2r+1) + j with i the
 * X-offset index, bilinear, zeros outside. */
int rpe_corr_lookup(const float *pyramid, const float *coords, float *out, int B, int h, int w,
                    int num_levels, int radius, void *stream);

/* Replaces RAFT.upsample_flow (/root/reference/core/RAFT/core/raft.py:66-77): convex 8x upsampling
 * with a softmax over the 9 neighbours.  flow (B,2,h,w), mask (B,576,h,w) -> out (B,2,8h,8w). */
int rpe_convex_upsample8(const float *flow, const float *mask, float *out, int B, int h, int w, void *stream);

int rpe_convex_upsample8_nhwc(const float *flow, const float *mask, int mask_ld, float *out, int B, int h, int w, void *stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" row: the RAFT convolutional trunk on tcgen05 tensor cores.
 * Replaces the convolutions of BasicUpdateBlock (/root/reference/core/RAFT/core/update.py:79-136) and of BasicEncoder
 * (/root/reference/core/RAFT/core/extractor.py:118-192) that the reference runs through cuDNN.  Activations are NHWC fp16 "split" planes (hi = fp16(v), lo = fp16(v - hi));
 * a convolution contracts over a list of (activation plane, weight) sources, so concatenated inputs are never
 * materialised and hi*hi + lo*hi + hi*lo reproduces fp32 convolutions to ~1e-5.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rpe_conv_source {
    const void *act_hi;  /* fp16 NHWC (N,H,W,c_total): hi plane                                       */
    const void *act_lo;  /* lo plane, or NULL: activations exact in one fp16 plane (w_lo may still be set)   */
    int c_total;         /* channel stride of the activation tensor (multiple of 8)                   */
    int c_offset;        /* first channel read (multiple of 8)                                        */
    int c_count;         /* channels read (multiple of 16); channels past the window read as zero     */
    const void *w_hi;    /* fp16 [kh*kw][cout_pad][w_cstride], K-major: hi plane                      */
    const void *w_lo;    /* lo plane or NULL                                                          */
    int w_cstride;       /* channel pitch of the weight rows (>= c_count, multiple of 8)              */
} rpe_conv_source;

typedef struct rpe_conv_desc {
    int n_sources;       /* 1..4; all sources have the same number of planes                          */
    rpe_conv_source src[4];
    int N, H, W;         /* INPUT size; output = (H + 2(kh/2) - kh)/stride + 1, same for W            */
    int kh, kw;          /* odd; zero padding (kh/2, kw/2)                                            */
    int stride;          /* 1 or 2 (0 = 1)                                                            */
    int cout, cout_pad;  /* real / padded (multiple of 16) output channels                            */
    const float *bias;   /* (cout) or NULL                                                            */
    const float *pre;    /* optional fp32 NHWC addend (N,OH,OW,pre_ld), added BEFORE the activation   */
    int pre_ld;
    const float *res;    /* optional fp32 NHWC residual: out = relu(act(v) * scale + res)             */
    int res_ld;
    int activation;      /* 0 none, 1 relu, 2 sigmoid, 3 tanh                                         */
    float out_scale;     /* applied after the activation                                              */
    float *out_f32;      /* NHWC fp32 output (N,OH,OW,f32_ld) at channel f32_offset, or NULL          */
    int f32_ld, f32_offset;
    void *out_hi, *out_lo; /* NHWC fp16 planes (N,OH,OW,bf_ld) at channel bf_offset; out_lo may be NULL */
    int bf_ld, bf_offset;
    /* fused SepConvGRU epilogues (update.py:45-60); 0 = plain epilogue.
     *   mode 1 (z|r gates, activation sigmoid, cout = 2*hidden): z = out[:hidden] -> out_f32, planes <- r * h  (h = aux)
     *   mode 2 (candidate, activation tanh, cout = hidden): h <- (1 - z) * h + z * q in place (h = aux, z = aux2), planes <- h
     * fused head of a following 3x3 convolution with 2 output channels (FlowHead, update.py:6-13):
     *   mode 3 (tap projection, cout <= 256): the activated outputs y are not stored; out_f32 (f32_ld >= 36) receives two slots of
     *           18 partial sums p[(ky*3+kx)*2 + o] = <y[half of the channels], w2[:, (ky*3+kx)*2 + o]>, w2 = aux2 as fp32 [cout][18];
     *           rpe_tap_gather3x3 adds the nine shifted maps and the bias. */
    int mode;
    float *aux;          /* hidden state h, fp32 NHWC (N,OH,OW,aux_ld)                                */
    int aux_ld;
    const float *aux2;   /* update gate z, fp32 NHWC (N,OH,OW,aux2_ld)                                */
    int aux2_ld;
    /* instance-norm partial sums fused into an fp32-only epilogue (mode 0, no pre / res / planes, cout % 16 == 0), or NULL:
     * fp32 [N * tiles_per_image * 4][cout_pad][2] = (sum, sum of squares) of the outputs of 32 pixels each; slots of pixels
     * outside the image hold zeros.  rpe_instnorm_stats_from_partials reduces them (extractor.py:23-56 nn.InstanceNorm2d). */
    float *stat_partials;
    /* The accumulators are multiplied by acc_scale before the bias is added (0 = 1).  Weights are packed pre-multiplied by the
     * inverse, a power of two chosen so that their fp16 lo plane is a normal number (|w| >= 2^-3): exact, and it keeps ~22
     * significant bits on small weights. */
    float acc_scale;
    /* The residual as fp16 split planes (hi + lo, channel pitch res_ld) instead of the fp32 tensor `res` (one or the other):
     * out = relu(act(v) * scale + (res_hi + res_lo)).  They may be the output planes of the same plan (BasicBlock /
     * ResidualBlock outputs of extractor.py:45-56 updated in place). */
    const void *res_hi, *res_lo;
} rpe_conv_desc;

int rpe_conv_plan_create(const rpe_conv_desc *desc, void **plan_out);   /* encodes the TMA descriptors once */
int rpe_conv_plan_run(void *plan, void *stream);
int rpe_conv_plan_tiles_per_image(void *plan); /* 128-pixel tiles per image (4 stat_partials slots each)   */
double rpe_conv_plan_flops(void *plan);      /* tensor-core flops of one run (x3 for the split arithmetic) */
int rpe_conv_plan_destroy(void *plan);

/* CorrBlock.__call__ writing the NHWC fp16 split planes the first motion-encoder convolution reads (ld >= 324). */
int rpe_corr_lookup_nhwc_split(const float *pyramid, const float *coords, void *out_hi, void *out_lo, int ld, int B, int h, int w,
                              int num_levels, int radius, void *stream);
/* Layout / split helpers and the element-wise pieces of the update operator (update.py:45-60, raft.py:112-121). */
int rpe_nchw_to_nhwc_split(const float *x, void *hi, void *lo, float *f32, int n, int C, int H, int W, int ld, int off,
                           int f32_ld, int f32_off, void *stream);
int rpe_nhwc_to_nchw(const float *x, float *out, int n, int C, int H, int W, int ld, int off, void *stream);
int rpe_flow_step(float *coords1, const float *delta, int delta_ld, void *col_hi, void *col_lo, int col_ld, void *x_hi,
                  void *x_lo, int x_ld, int x_off, int n, int h, int w, void *stream);
int rpe_tap_gather3x3(const float *part, int part_ld, const float *bias, float *out, int out_ld, int n, int h, int w, void *stream);
int rpe_gru_gate(const float *zr, float *h, const float *q, void *out_hi, void *out_lo, int out_ld, int out_off,
                 long long npix, int mode, void *stream);

/* Encoder companions (reference: /root/reference/core/RAFT/core/extractor.py:118-192, raft.py:82-83).
 * rpe_im2col7s2_split: 7x7 / stride 2 / pad 3 windows of the normalised image 2*(v/255)-1 as the K axis of a 1x1 convolution,
 *   k = ky*24 + kx*3 + c (168 of `ld` = 176 channels, the last 8 written as zeros); img NCHW fp32 (n,3,H,W) -> fp16 split
 *   planes (n, H/2, W/2, 176).
 * rpe_instnorm_stats: InstanceNorm2d statistics (biased variance) of an NHWC fp32 tensor (n,HW,C): stats (n,C,2) = mean, rstd.
 * rpe_norm_act_split: y = [relu]((a - mean_a) * rstd_a), optionally y = relu(y + (b - mean_b) * rstd_b) (stats may be NULL =
 *   identity); writes fp32 NHWC and / or fp16 split planes with channel pitch ld. */
int rpe_im2col7s2_split(const float *img, void *out_hi, void *out_lo, int n, int H, int W, int ld, void *stream);
/* The same with the frames as uint8 RGB (n,3,H,W), the way a camera delivers them (ship uint8, convert on the GPU):
 * the uint8 -> float conversion of dataset/stereo_dataset.py:36-37 happens inside the kernels that read the image. */
int rpe_im2col7s2_split_u8(const unsigned char *img, void *out_hi, void *out_lo, int n, int H, int W, int ld, void *stream);
/*   out_lo = NULL selects the RAW single-plane form: the window holds the pixel values themselves (integers 0..255 are exact in one
 *   fp16 plane; positions outside the image hold 127.5, whose normalisation is the reference's zero padding) and the caller folds
 *   2 v / 255 - 1 into the stem weights (w' = 2 w / 255, b' = b - sum w).  The convolution then reads one activation plane and
 *   two weight planes (rpe_conv_source.act_lo = NULL, w_lo set): two tensor-core products per multiply-add instead of three. */
size_t rpe_instnorm_workspace_bytes(int n, int C);
int rpe_instnorm_stats(const float *x, float *stats, int n, int HW, int C, float eps, void *workspace, size_t workspace_bytes,
                       void *stream);
/* The same statistics from the partial sums a convolution wrote through rpe_conv_desc.stat_partials
 * ([n * slots_per_image][ld][2] fp32, slots_per_image = 4 * rpe_conv_plan_tiles_per_image): no second pass over the tensor. */
int rpe_instnorm_stats_from_partials(const float *partials, float *stats, int n, int slots_per_image, int C, int ld, int HW, float eps,
                                     void *workspace, size_t workspace_bytes, void *stream);   /* rpe_instnorm_workspace_bytes(n, C) */
int rpe_norm_act_split(const float *a, const float *stats_a, int relu_a, const float *b, const float *stats_b, float *out_f32,
                       void *out_hi, void *out_lo, int ld, int n, int HW, int C, void *stream);
/* The same with the residual addend given as fp16 split planes (b_hi + b_lo, channel pitch b_ld; excludes b / stats_b): the
 * skip connection of extractor.py:45-56 read from the planes the next convolution consumes anyway, so that the encoder keeps no
 * fp32 copy of its residual stream.  The addend planes may be the output planes (in-place block output). */
int rpe_norm_act_split_res(const float *a, const float *stats_a, int relu_a, const float *b, const float *stats_b, const void *b_hi,
                           const void *b_lo, int b_ld, float *out_f32, void *out_hi, void *out_lo, int ld, int n, int HW, int C,
                           void *stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" row: the confidence heads (TinyUNet, /root/reference/core/unet/unet.py:8-82, wrapped with Sigmoid at
 * core/pose/pose_net.py:24-27) on the convolution plans above.  Each un-padded 3x3 convolution runs as a "same" convolution on
 * its input grid and the reference's result is the interior (the VALID REGION: origin + size inside the grid); these kernels
 * carry data between the grids.  All fp32 tensors NHWC with channel pitch ld; outputs fp16 split planes unless noted.
 *   rpe_downsample8_planes  pose_net.py:110-113 F.interpolate(scale_factor=0.125, 'bilinear') of up to three NCHW fp32 tensors,
 *                           concatenated at channel ch_offset of planes (n, H/8, W/8, ld)
 *   rpe_pool2_planes        F.max_pool2d(x, 2) of the valid region (y0, x0, 2 oh, 2 ow) of channels [c_in_off, c_in_off + C)
 *   rpe_upcat_planes        torch.cat((ConvTranspose2d(k=2, s=2)(x), centre_crop(skip)), 1) (unet.py:52-61): `up` holds the
 *                           transposed convolution as 4 groups of c_up channels, group dy * 2 + dx, valid origin (uy0, ux0)
 *   rpe_resize_sigmoid      sigmoid(F.interpolate(logits, (H, W), mode='bilinear')) of channel ch of the valid region
 *                           (y0, x0, ih, iw) -> out (n,1,H,W) fp32 */
int rpe_downsample8_planes(const void *src0, int c0, const void *src1, int c1, const void *src2, int c2, int src_u8_mask, void *out_hi,
                           void *out_lo, int ld, int ch_offset, int n, int H, int W, void *stream);   /* bit s of src_u8_mask: source s is uint8 */
int rpe_pool2_planes(const float *x, int H, int W, int ld_in, int c_in_off, int y0, int x0, void *out_hi, void *out_lo, int oh, int ow,
                     int ld_out, int C, int n, void *stream);
int rpe_upcat_planes(const float *up, int Hu, int Wu, int ld_u, int uy0, int ux0, int c_up, const float *skip, int Hk, int Wk, int ld_k,
                     int k_off, int ky0, int kx0, int c_skip, void *out_hi, void *out_lo, int oh, int ow, int ld_out, int n, void *stream);
int rpe_resize_sigmoid(const float *logits, int Hl, int Wl, int ld, int ch, int y0, int x0, int ih, int iw, float *out, int n, int H, int W,
                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RPE_B200_H */
