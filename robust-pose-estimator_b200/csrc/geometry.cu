// Stage 2 (stereo depth lifting + pinhole back-projection) and stage 5 (flow warping) kernels.
//
// HBM-bound per-pixel work: coalesced 128-bit accesses along W, one pass over every tensor.
// Reference semantics: /root/reference/core/pose/pose_net.py:73-79,104-113,121-125 and
// /root/reference/core/interpol/flow_utils.py:4-26.  Every fp32 operation
// that feeds a mask or an index uses an explicit round-to-nearest intrinsic (no FMA contraction)
// in the reference's operation order, so validity masks and nearest-neighbour indices are bit-exact
// with the reference's CPU fp32 run.
#include "common.cuh"

namespace rpe {

struct Mat3 {
    float m[9];
};

// K^-1 via the adjugate in fp64, rounded once to fp32.  For the upper-triangular pinhole K the
// reference's LU solve yields the correctly rounded 1/f, -c/f, which this reproduces.
__device__ __forceinline__ Mat3 inverse3(const float *__restrict__ K) {
    double a = K[0], b = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
    double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    double det = a * A + b * B + c * C;
    double r = 1.0 / det;
    Mat3 o;
    o.m[0] = (float)(A * r);
    o.m[1] = (float)(-(b * i - c * h) * r);
    o.m[2] = (float)((b * f - c * e) * r);
    o.m[3] = (float)(B * r);
    o.m[4] = (float)((a * i - c * g) * r);
    o.m[5] = (float)(-(a * f - c * d) * r);
    o.m[6] = (float)(C * r);
    o.m[7] = (float)(-(a * h - b * g) * r);
    o.m[8] = (float)((a * e - b * d) * r);
    return o;
}

__device__ __forceinline__ void backproject(const Mat3 &Ki, float u, float v, float d, float &X, float &Y, float &Z) {
    // rays = K^-1 [u v 1]^T accumulated like a GEMM micro-kernel (k ascending), then scaled by depth
    float rx = fmaf(Ki.m[2], 1.0f, fmaf(Ki.m[1], v, Ki.m[0] * u));
    float ry = fmaf(Ki.m[5], 1.0f, fmaf(Ki.m[4], v, Ki.m[3] * u));
    float rz = fmaf(Ki.m[8], 1.0f, fmaf(Ki.m[7], v, Ki.m[6] * u));
    X = __fmul_rn(d, rx);
    Y = __fmul_rn(d, ry);
    Z = __fmul_rn(d, rz);
}

// One thread handles 4 consecutive pixels of a row-major image (HW % 4 == 0 enforced by the host).
template <bool kFromFlow>
__global__ void __launch_bounds__(256) depth_proj_kernel(const float *__restrict__ src,   // stereo flow (n,2,HW) or depth (n,HW)
                                                         const float *__restrict__ bf, const float *__restrict__ K,
                                                         uint8_t *__restrict__ mask_inout, float *__restrict__ depth_out,
                                                         uint8_t *__restrict__ valid_out, float *__restrict__ pcl,
                                                         int rescale, float scale, int H, int W) {
    const int HW = H * W;
    const int b = blockIdx.y;
    __shared__ Mat3 sKi;
    if (threadIdx.x == 0) sKi = inverse3(K + 9 * b);
    __syncthreads();
    const Mat3 Ki = sKi;
    const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= HW) return;
    float d[4];
    if (kFromFlow) {
        const float4 fx = __ldg(reinterpret_cast<const float4 *>(src + (size_t)b * 2 * HW + i4));
        const float base = __ldg(bf + b);
        const float f[4] = {fx.x, fx.y, fx.z, fx.w};
        uint32_t vbits = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float dd = __fdiv_rn(base, -f[k]);
            bool ok = (dd > 0.0f) && (dd <= 1.0f);
            d[k] = ok ? dd : 1.0f;
            vbits |= (ok ? 1u : 0u) << (8 * k);
        }
        if (valid_out) *reinterpret_cast<uint32_t *>(valid_out + (size_t)b * HW + i4) = vbits;
        if (mask_inout) {
            uint32_t *mp = reinterpret_cast<uint32_t *>(mask_inout + (size_t)b * HW + i4);
            uint32_t m = *mp;
            // bool & bool per byte: normalise the input bytes to 0/1 first
            uint32_t mn = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) mn |= (((m >> (8 * k)) & 0xffu) ? 1u : 0u) << (8 * k);
            *mp = mn & vbits;
        }
        if (depth_out) *reinterpret_cast<float4 *>(depth_out + (size_t)b * HW + i4) = make_float4(d[0], d[1], d[2], d[3]);
    } else {
        const float4 dv = __ldg(reinterpret_cast<const float4 *>(src + (size_t)b * HW + i4));
        d[0] = dv.x, d[1] = dv.y, d[2] = dv.z, d[3] = dv.w;
        if (rescale) {
#pragma unroll
            for (int k = 0; k < 4; ++k) d[k] = __fmul_rn(__fdiv_rn(d[k], scale), scale);
        }
    }
    if (!pcl) return;
    float X[4], Y[4], Z[4];
    const int row = i4 / W;            // W % 4 == 0 -> the 4 pixels share a row
    const int col = i4 - row * W;
    const float v = (float)row + 0.5f;
#pragma unroll
    for (int k = 0; k < 4; ++k) backproject(Ki, (float)(col + k) + 0.5f, v, d[k], X[k], Y[k], Z[k]);
    float *p = pcl + (size_t)b * 3 * HW + i4;
    *reinterpret_cast<float4 *>(p) = make_float4(X[0], X[1], X[2], X[3]);
    *reinterpret_cast<float4 *>(p + HW) = make_float4(Y[0], Y[1], Y[2], Y[3]);
    *reinterpret_cast<float4 *>(p + 2 * HW) = make_float4(Z[0], Z[1], Z[2], Z[3]);
}

// ATen grid_sampler (CPU, align_corners=True): ix = (g + 1) * ((size - 1) / 2), with
// g = 2 * (flow + idx) / (size - 1) - 1 from remap_from_flow.  One rounding per operation.
__device__ __forceinline__ float sample_coord(float flow, int idx, int size) {
    const float sm1 = (float)(size - 1);
    float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn(flow, (float)idx)), sm1), 1.0f);
    return __fmul_rn(__fadd_rn(g, 1.0f), sm1 * 0.5f);
}

struct WarpSrc {
    const float *src[3];
    float *dst[3];
    int ch[3];
    int u8_mask;          // bit s: source s holds uint8 values (a camera frame as it is); its warped output is fp32 all the same
};
__device__ __forceinline__ float warp_ld(const float *p, int o, bool u8) {
    return u8 ? (float)__ldg(reinterpret_cast<const uint8_t *>(p) + o) : __ldg(p + o);
}

__global__ void __launch_bounds__(256) warp8_mask_kernel(WarpSrc t, const uint8_t *__restrict__ mask2,
                                                         const float *__restrict__ flow, uint8_t *__restrict__ mask2w,
                                                         int H, int W) {
    const int HW = H * W;
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int row = i / W, col = i - row * W;
    const float fx = __ldg(flow + (size_t)b * 2 * HW + i);
    const float fy = __ldg(flow + (size_t)b * 2 * HW + HW + i);
    const float x = sample_coord(fx, col, W);
    const float y = sample_coord(fy, row, H);

    // ---- nearest-neighbour mask warp: round-half-even, zeros outside
    if (mask2w) {
        const float xn = rintf(x), yn = rintf(y);
        uint8_t m = 0;
        if (xn > -1.0f && xn < (float)W && yn > -1.0f && yn < (float)H)
            m = __ldg(mask2 + (size_t)b * HW + (int)yn * W + (int)xn) ? 1 : 0;
        mask2w[(size_t)b * HW + i] = m;   // (warped > 0) & bool(warped)
    }

    // ---- bilinear warp, zeros padding
    const float x0 = floorf(x), y0 = floorf(y);
    const float wx = __fsub_rn(x, x0), ex = __fsub_rn(1.0f, wx);
    const float wy = __fsub_rn(y, y0), ey = __fsub_rn(1.0f, wy);
    const float w_nw = __fmul_rn(ey, ex), w_ne = __fmul_rn(ey, wx), w_sw = __fmul_rn(wy, ex), w_se = __fmul_rn(wy, wx);
    const bool in_w = x0 > -1.0f && x0 < (float)W, in_e = (x0 + 1.0f) > -1.0f && (x0 + 1.0f) < (float)W;
    const bool in_n = y0 > -1.0f && y0 < (float)H, in_s = (y0 + 1.0f) > -1.0f && (y0 + 1.0f) < (float)H;
    const bool any = (in_w || in_e) && (in_n || in_s);
    int o_nw = 0;
    if (any) o_nw = (int)y0 * W + (int)x0;   // corners addressed relative to (y0, x0); only dereferenced when in range
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        if (t.src[s] == nullptr) continue;
        const bool u8 = (t.u8_mask >> s) & 1;
        for (int c = 0; c < t.ch[s]; ++c) {
            const float *p = u8 ? reinterpret_cast<const float *>(reinterpret_cast<const uint8_t *>(t.src[s]) + ((size_t)b * t.ch[s] + c) * HW)
                                : t.src[s] + ((size_t)b * t.ch[s] + c) * HW;
            float acc = 0.0f;
            if (in_n && in_w) acc = warp_ld(p, o_nw, u8) * w_nw;
            if (in_n && in_e) acc += warp_ld(p, o_nw + 1, u8) * w_ne;
            if (in_s && in_w) acc += warp_ld(p, o_nw + W, u8) * w_sw;
            if (in_s && in_e) acc += warp_ld(p, o_nw + W + 1, u8) * w_se;
            t.dst[s][((size_t)b * t.ch[s] + c) * HW + i] = acc;
        }
    }
}

// Nearest-neighbour warp of a float tensor (remap_from_flow_nearest on arbitrary inputs).
__global__ void __launch_bounds__(256) remap_nearest_kernel(const float *__restrict__ x, const float *__restrict__ flow,
                                                            float *__restrict__ out, int C, int H, int W) {
    const int HW = H * W;
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int row = i / W, col = i - row * W;
    const float xs = rintf(sample_coord(__ldg(flow + (size_t)b * 2 * HW + i), col, W));
    const float ys = rintf(sample_coord(__ldg(flow + (size_t)b * 2 * HW + HW + i), row, H));
    const bool ok = xs > -1.0f && xs < (float)W && ys > -1.0f && ys < (float)H;
    const int o = ok ? (int)ys * W + (int)xs : 0;
    for (int c = 0; c < C; ++c) out[((size_t)b * C + c) * HW + i] = ok ? __ldg(x + ((size_t)b * C + c) * HW + o) : 0.0f;
}

struct CatSrc {
    const float *src[3];
    int ch[3];
};

// out[c, i, j] = 0.5 * (0.5 v[8i+3, 8j+3] + 0.5 v[8i+3, 8j+4]) + 0.5 * (0.5 v[8i+4, 8j+3] + 0.5 v[8i+4, 8j+4])
// (ATen upsample_bilinear2d, align_corners=False, scale 8: source index 8i + 3.5).
__global__ void __launch_bounds__(256) downsample8_cat_kernel(CatSrc t, float *__restrict__ out, int out_ch_offset,
                                                              int out_ch_total, int H, int W) {
    const int h8 = H / 8, w8 = W / 8;
    const int ctot = t.ch[0] + t.ch[1] + t.ch[2];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (idx >= ctot * h8 * w8) return;
    const int j = idx % w8, i = (idx / w8) % h8;
    int c = idx / (w8 * h8);
    const int cc = c;
    const float *p;
    if (c < t.ch[0]) {
        p = t.src[0] + ((size_t)b * t.ch[0] + c) * H * W;
    } else if (c < t.ch[0] + t.ch[1]) {
        c -= t.ch[0];
        p = t.src[1] + ((size_t)b * t.ch[1] + c) * H * W;
    } else {
        c -= t.ch[0] + t.ch[1];
        p = t.src[2] + ((size_t)b * t.ch[2] + c) * H * W;
    }
    p += (size_t)(8 * i + 3) * W + 8 * j + 3;
    const float2 a = make_float2(__ldg(p), __ldg(p + 1));
    const float2 d = make_float2(__ldg(p + W), __ldg(p + W + 1));
    const float top = __fadd_rn(__fmul_rn(0.5f, a.x), __fmul_rn(0.5f, a.y));
    const float bot = __fadd_rn(__fmul_rn(0.5f, d.x), __fmul_rn(0.5f, d.y));
    out[(((size_t)b * out_ch_total + out_ch_offset + cc) * h8 + i) * w8 + j] =
        __fadd_rn(__fmul_rn(0.5f, top), __fmul_rn(0.5f, bot));
}

}  // namespace rpe

extern "C" {

int rpe_depth_proj(const float *stereo_flow, const float *bf, const float *K, uint8_t *mask_inout, float *depth,
                   uint8_t *valid, float *pcl, int n, int H, int W, void *stream) {
    if (!stereo_flow || !bf || !K || n <= 0 || H <= 0 || W <= 0) return RPE_ERR_INVALID_ARG;
    if (W % 4 != 0) return RPE_ERR_INVALID_ARG;
    if (!rpe::aligned16(stereo_flow) || (depth && !rpe::aligned16(depth)) || (pcl && !rpe::aligned16(pcl)) ||
        (valid && (reinterpret_cast<uintptr_t>(valid) & 3u)) || (mask_inout && (reinterpret_cast<uintptr_t>(mask_inout) & 3u)))
        return RPE_ERR_ALIGNMENT;
    const int HW = H * W;
    dim3 grid((HW / 4 + 255) / 256, n);
    rpe::depth_proj_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(stereo_flow, bf, K, mask_inout, depth, valid, pcl,
                                                                        0, 1.0f, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_proj(const float *depth, const float *K, float *pcl, int rescale, float scale, int n, int H, int W, void *stream) {
    if (!depth || !K || !pcl || n <= 0 || H <= 0 || W <= 0) return RPE_ERR_INVALID_ARG;
    if (W % 4 != 0) return RPE_ERR_INVALID_ARG;
    if (!rpe::aligned16(depth) || !rpe::aligned16(pcl)) return RPE_ERR_ALIGNMENT;
    const int HW = H * W;
    dim3 grid((HW / 4 + 255) / 256, n);
    rpe::depth_proj_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(depth, nullptr, K, nullptr, nullptr, nullptr, pcl,
                                                                         rescale, scale, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

static int warp8_mask_impl(const float *pcl2, const void *img2, int img_u8, const float *sflow2, const uint8_t *mask2, const float *flow,
                           float *pcl2w, float *img2w, float *sflow2w, uint8_t *mask2w, int n, int H, int W, void *stream);

int rpe_warp8_mask(const float *pcl2, const float *img2, const float *sflow2, const uint8_t *mask2, const float *flow,
                   float *pcl2w, float *img2w, float *sflow2w, uint8_t *mask2w, int n, int H, int W, void *stream) {
    return warp8_mask_impl(pcl2, img2, 0, sflow2, mask2, flow, pcl2w, img2w, sflow2w, mask2w, n, H, W, stream);
}

int rpe_warp8_mask_u8(const float *pcl2, const unsigned char *img2, const float *sflow2, const uint8_t *mask2, const float *flow,
                      float *pcl2w, float *img2w, float *sflow2w, uint8_t *mask2w, int n, int H, int W, void *stream) {
    return warp8_mask_impl(pcl2, img2, 1, sflow2, mask2, flow, pcl2w, img2w, sflow2w, mask2w, n, H, W, stream);
}

static int warp8_mask_impl(const float *pcl2, const void *img2, int img_u8, const float *sflow2, const uint8_t *mask2, const float *flow,
                           float *pcl2w, float *img2w, float *sflow2w, uint8_t *mask2w, int n, int H, int W, void *stream) {
    if (!flow || n <= 0 || H <= 1 || W <= 1) return RPE_ERR_INVALID_ARG;
    if ((pcl2 == nullptr) != (pcl2w == nullptr) || (img2 == nullptr) != (img2w == nullptr) ||
        (sflow2 == nullptr) != (sflow2w == nullptr) || (mask2 == nullptr) != (mask2w == nullptr))
        return RPE_ERR_INVALID_ARG;
    rpe::WarpSrc t;
    t.src[0] = pcl2, t.dst[0] = pcl2w, t.ch[0] = 3;
    t.src[1] = reinterpret_cast<const float *>(img2), t.dst[1] = img2w, t.ch[1] = 3;
    t.src[2] = sflow2, t.dst[2] = sflow2w, t.ch[2] = 2;
    t.u8_mask = img_u8 ? 2 : 0;
    dim3 grid((H * W + 255) / 256, n);
    rpe::warp8_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, mask2, flow, mask2w, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_remap_bilinear(const float *x, int C, const float *flow, float *out, int n, int H, int W, void *stream) {
    if (!x || !flow || !out || C <= 0 || n <= 0 || H <= 1 || W <= 1) return RPE_ERR_INVALID_ARG;
    rpe::WarpSrc t;
    t.src[0] = x, t.dst[0] = out, t.ch[0] = C;
    t.src[1] = nullptr, t.dst[1] = nullptr, t.ch[1] = 0;
    t.src[2] = nullptr, t.dst[2] = nullptr, t.ch[2] = 0;
    t.u8_mask = 0;
    dim3 grid((H * W + 255) / 256, n);
    rpe::warp8_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, nullptr, flow, nullptr, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_remap_nearest(const float *x, int C, const float *flow, float *out, int n, int H, int W, void *stream) {
    if (!x || !flow || !out || C <= 0 || n <= 0 || H <= 1 || W <= 1) return RPE_ERR_INVALID_ARG;
    dim3 grid((H * W + 255) / 256, n);
    rpe::remap_nearest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, flow, out, C, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_downsample8_cat(const float *src0, int c0, const float *src1, int c1, const float *src2, int c2, float *out,
                        int out_ch_offset, int out_ch_total, int n, int H, int W, void *stream) {
    if (!out || n <= 0 || H < 8 || W < 8 || (H % 8) || (W % 8)) return RPE_ERR_INVALID_ARG;
    rpe::CatSrc t;
    t.src[0] = src0, t.ch[0] = src0 ? c0 : 0;
    t.src[1] = src1, t.ch[1] = src1 ? c1 : 0;
    t.src[2] = src2, t.ch[2] = src2 ? c2 : 0;
    const int ctot = t.ch[0] + t.ch[1] + t.ch[2];
    if (ctot <= 0 || out_ch_offset < 0 || out_ch_offset + ctot > out_ch_total) return RPE_ERR_INVALID_ARG;
    const int total = ctot * (H / 8) * (W / 8);
    dim3 grid((total + 255) / 256, n);
    rpe::downsample8_cat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, out, out_ch_offset, out_ch_total, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

}  // extern "C"
