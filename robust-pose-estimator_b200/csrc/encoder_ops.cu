// Element-wise / reduction companions of the tcgen05 convolution path of the RAFT encoders
// (reference: /root/reference/core/RAFT/core/extractor.py:118-192 BasicEncoder, :6-56 ResidualBlock; input scaling
// core/RAFT/core/raft.py:82-83).  All activations NHWC; "split" = fp16 hi/lo planes (conv.cu).
#include <cuda_fp16.h>
#include "common.cuh"

namespace rpe {

__device__ __forceinline__ void split4(const float *v, uint2 &hi, uint2 &lo) {
    plane_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = to_plane(v[k]);
        l[k] = to_plane_lo(v[k] - plane_to_float(h[k]));
    }
    hi = *reinterpret_cast<uint2 *>(h);
    lo = *reinterpret_cast<uint2 *>(l);
}

// ------------------------------------------------------------------------------------------------
// Stem im2col: 7x7 / stride 2 / pad 3 windows of the normalised image 2*(v/255)-1 (raft.py:82-83), written as the K axis of
// a 1x1 convolution: k = ky*24 + kx*3 + c (21 real values per filter row, padded to 24 so that every row is three aligned
// 16-byte stores per plane).  img NCHW fp32 (n,3,H,W) -> planes (n, H/2, W/2, ld).  One thread = one (pixel, ky).
// ------------------------------------------------------------------------------------------------
// CTA = 32 consecutive output pixels of one output row: the 7 input rows x 3 channels x 69 columns they need are staged in
// shared memory with coalesced loads (normalised on the way in, zeros outside the image), then thread (pixel, ky) assembles
// its 24 values in a second shared tile and the CTA copies its contiguous 32 x 352-byte output range in 16-byte chunks.
constexpr int kStemPx = 32, kStemCols = 2 * kStemPx + 5, kStemPitch = 72, kStemLd = 176;
// kRaw (uint8 frames only): the window holds the RAW pixel values -- integers 0..255 are exact in ONE fp16 plane, so no lo plane
// is written -- and positions outside the image hold 127.5, the raw value whose normalisation 2 v / 255 - 1 is the zero the
// reference pads with; the affine map itself is folded into the stem weights (w' = 2 w / 255, b' = b - sum w; encoder_tc.py).
// Half the bytes of the split form and two tensor-core products per multiply-add instead of three.
template <typename T, bool kRaw>      // T: float (the reference's 0..255 float tensors) or uint8_t (camera frames as they are)
__global__ void __launch_bounds__(256) im2col7s2_kernel(const T *__restrict__ img, plane_t *__restrict__ hi,
                                                        plane_t *__restrict__ lo, int H, int W, int OH, int OW, int ld) {
    __shared__ float tile[7][3][kStemPitch];
    const int ox0 = blockIdx.x * kStemPx, oy = blockIdx.y, n = blockIdx.z;
    const int x0 = 2 * ox0 - 3, y0 = 2 * oy - 3;
    const T *base = img + (size_t)n * 3 * H * W;
    for (int i = threadIdx.x; i < 7 * 3 * kStemPitch; i += blockDim.x) {
        const int col = i % kStemPitch, rc = i / kStemPitch;
        const int c = rc % 3, r = rc / 3;
        const int y = y0 + r, x = x0 + col;
        float v = kRaw ? 127.5f : 0.0f;
        if (col < kStemCols && y >= 0 && y < H && x >= 0 && x < W) {
            v = (float)__ldg(base + ((size_t)c * H + y) * W + x);
            if (!kRaw) v = 2.0f * (v / 255.0f) - 1.0f;
        }
        tile[r][c][col] = v;
    }
    __syncthreads();
    // assemble the (pixel, ky) rows in shared memory, then copy the CTA's contiguous output range with 16-byte chunks
    __shared__ __align__(16) plane_t s_hi[kStemPx][kStemLd], s_lo[kStemPx][kStemLd];
    const int npx = min(kStemPx, OW - ox0);
    if (threadIdx.x < kStemPx * 7) {
        const int ky = threadIdx.x % 7, p = threadIdx.x / 7;
        float v[24];
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[kx * 3 + c] = tile[ky][c][2 * p + kx];
        }
        v[21] = v[22] = v[23] = 0.0f;
#pragma unroll
        for (int g = 0; g < 6; ++g) {
            uint2 h, l;
            split4(v + g * 4, h, l);
            *reinterpret_cast<uint2 *>(&s_hi[p][ky * 24 + g * 4]) = h;
            if (!kRaw) *reinterpret_cast<uint2 *>(&s_lo[p][ky * 24 + g * 4]) = l;
        }
        if (ky == 0) {                                       // channels 168..175: padding of the K axis, kept at zero
            *reinterpret_cast<uint4 *>(&s_hi[p][168]) = make_uint4(0, 0, 0, 0);
            if (!kRaw) *reinterpret_cast<uint4 *>(&s_lo[p][168]) = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    const size_t pix0 = ((size_t)n * OH + oy) * OW + ox0;
    uint4 *ghi = reinterpret_cast<uint4 *>(hi + pix0 * kStemLd);
    const uint4 *shi = reinterpret_cast<const uint4 *>(&s_hi[0][0]), *slo = reinterpret_cast<const uint4 *>(&s_lo[0][0]);
    const int chunks = npx * (kStemLd * 2 / 16);
    if (kRaw) {
        for (int i = threadIdx.x; i < chunks; i += blockDim.x) ghi[i] = shi[i];
    } else {
        uint4 *glo = reinterpret_cast<uint4 *>(lo + pix0 * kStemLd);
        for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
            ghi[i] = shi[i];
            glo[i] = slo[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Instance-norm statistics (InstanceNorm2d without affine parameters, eps 1e-5, biased variance): per (sample, channel)
// mean and 1/sqrt(var + eps) of an NHWC fp32 tensor.  Two deterministic passes: per-block partial sums in fp64, then a
// fixed-order reduction of the partials.
// ------------------------------------------------------------------------------------------------
constexpr int kStatBlocks = 64;     // partial blocks per sample

__global__ void __launch_bounds__(256) instnorm_partial_kernel(const float *__restrict__ x, double *__restrict__ partial, int HW, int C,
                                                               int ld) {
    extern __shared__ double sred[];                 // [rows][C][2]
    const int n = blockIdx.y, blk = blockIdx.x;
    const int c4n = C / 4;
    const int rows = 256 / c4n;
    const int c4 = threadIdx.x % c4n, r = threadIdx.x / c4n;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (r < rows) {
        const int per = (HW + kStatBlocks - 1) / kStatBlocks;
        const int p0 = blk * per, p1 = min(HW, p0 + per);
        const float *base = x + (size_t)n * HW * ld + c4 * 4;
        for (int p = p0 + r; p < p1; p += rows) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + (size_t)p * ld));
            s[0] += v.x, s[1] += v.y, s[2] += v.z, s[3] += v.w;
            q[0] += (double)v.x * v.x, q[1] += (double)v.y * v.y, q[2] += (double)v.z * v.z, q[3] += (double)v.w * v.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sred[((size_t)r * C + c4 * 4 + k) * 2] = s[k];
            sred[((size_t)r * C + c4 * 4 + k) * 2 + 1] = q[k];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a = 0, b = 0;
        for (int rr = 0; rr < rows; ++rr) {
            a += sred[((size_t)rr * C + c) * 2];
            b += sred[((size_t)rr * C + c) * 2 + 1];
        }
        partial[(((size_t)n * kStatBlocks + blk) * C + c) * 2] = a;
        partial[(((size_t)n * kStatBlocks + blk) * C + c) * 2 + 1] = b;
    }
}

__global__ void instnorm_final_kernel(const double *__restrict__ partial, float *__restrict__ stats, int HW, int C, float eps) {
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a = 0, b = 0;
        for (int blk = 0; blk < kStatBlocks; ++blk) {
            a += partial[(((size_t)n * kStatBlocks + blk) * C + c) * 2];
            b += partial[(((size_t)n * kStatBlocks + blk) * C + c) * 2 + 1];
        }
        const double mean = a / HW;
        double var = b / HW - mean * mean;
        if (var < 0) var = 0;
        stats[((size_t)n * C + c) * 2] = (float)mean;
        stats[((size_t)n * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// Statistics from the per-tile partial sums that the convolution epilogue writes (conv.cu kind 6): partial
// [n * slots][ld][2] fp32 (sum, sum of squares over 32 pixels each).  First level of a fixed-order fp64 reduction: block
// (blk, n) sums the slots blk, blk + kStatBlocks, ... into the same [n][kStatBlocks][C][2] layout instnorm_final_kernel reads.
__global__ void __launch_bounds__(256) instnorm_from_partials_kernel(const float *__restrict__ partial, double *__restrict__ blocks, int slots,
                                                                     int C, int ld) {
    __shared__ double sred[256][2];
    const int n = blockIdx.y, blk = blockIdx.x;
    const int rows = 256 / C;                            // C <= 256 (checked by the host)
    const int c = threadIdx.x % C, r = threadIdx.x / C;
    double a = 0, b = 0;
    if (r < rows) {
        const float2 *base = reinterpret_cast<const float2 *>(partial) + (size_t)n * slots * ld + c;
        for (int sl = blk + kStatBlocks * r; sl < slots; sl += kStatBlocks * rows) {
            const float2 v = __ldg(base + (size_t)sl * ld);
            a += (double)v.x;
            b += (double)v.y;
        }
    }
    sred[threadIdx.x][0] = a, sred[threadIdx.x][1] = b;
    __syncthreads();
    if (threadIdx.x < C) {
        a = 0, b = 0;
        for (int rr = 0; rr < rows; ++rr) a += sred[rr * C + threadIdx.x][0], b += sred[rr * C + threadIdx.x][1];
        blocks[(((size_t)n * kStatBlocks + blk) * C + threadIdx.x) * 2] = a;
        blocks[(((size_t)n * kStatBlocks + blk) * C + threadIdx.x) * 2 + 1] = b;
    }
}

// ------------------------------------------------------------------------------------------------
//   ya = stats_a ? (a - mean_a) * rstd_a : a;   if relu_a: ya = max(ya, 0)
//   y  = b ? max(ya + (stats_b ? (b - mean_b) * rstd_b : b), 0) : ya
// a, b fp32 NHWC with C channels (dense); y -> optional fp32 NHWC (dense) and / or split planes with channel pitch ld.
// The addend may instead be given as split planes (b_hi + b_lo, channel pitch b_ld, no statistics): the residual stream of the
// encoder then lives in its planes only (hi + lo carries 22 mantissa bits) and no fp32 copy of it is written or read.  The
// addend planes may alias the output planes (every thread reads its 4 channels before it writes them).
// ------------------------------------------------------------------------------------------------
template <typename idx_t>
__global__ void __launch_bounds__(256) norm_act_kernel(const float *__restrict__ a, const float *__restrict__ sa, int relu_a,
                                                       const float *__restrict__ b, const float *__restrict__ sb,
                                                       const plane_t *b_hi, const plane_t *b_lo, int b_ld, float *__restrict__ out,
                                                       plane_t *hi, plane_t *lo, int ld, int HW, int C, long long total4) {
    const idx_t i = (idx_t)blockIdx.x * (idx_t)blockDim.x + threadIdx.x;
    if ((long long)i >= total4) return;
    const int c4n = C / 4;
    const idx_t pix = i / (idx_t)c4n;
    const int c = (int)(i - pix * (idx_t)c4n) * 4;
    const int n = (int)(pix / (idx_t)HW);
    const float4 av = __ldg(reinterpret_cast<const float4 *>(a + pix * C + c));
    float y[4] = {av.x, av.y, av.z, av.w};
    if (sa) {
        const float *s = sa + ((size_t)n * C + c) * 2;
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = (y[k] - __ldg(s + 2 * k)) * __ldg(s + 2 * k + 1);
    }
    if (relu_a) {
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = fmaxf(y[k], 0.0f);
    }
    if (b) {
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(b + pix * C + c));
        float z[4] = {bv.x, bv.y, bv.z, bv.w};
        if (sb) {
            const float *s = sb + ((size_t)n * C + c) * 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) z[k] = (z[k] - __ldg(s + 2 * k)) * __ldg(s + 2 * k + 1);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = fmaxf(y[k] + z[k], 0.0f);
    } else if (b_hi) {
        const uint2 h = *reinterpret_cast<const uint2 *>(b_hi + pix * b_ld + c), l = *reinterpret_cast<const uint2 *>(b_lo + pix * b_ld + c);
        const float2 h01 = __half22float2(*reinterpret_cast<const plane2_t *>(&h.x)), h23 = __half22float2(*reinterpret_cast<const plane2_t *>(&h.y));
        const float2 l01 = __half22float2(*reinterpret_cast<const plane2_t *>(&l.x)), l23 = __half22float2(*reinterpret_cast<const plane2_t *>(&l.y));
        y[0] = fmaxf(y[0] + (h01.x + l01.x), 0.0f), y[1] = fmaxf(y[1] + (h01.y + l01.y), 0.0f);
        y[2] = fmaxf(y[2] + (h23.x + l23.x), 0.0f), y[3] = fmaxf(y[3] + (h23.y + l23.y), 0.0f);
    }
    if (out) *reinterpret_cast<float4 *>(out + pix * C + c) = make_float4(y[0], y[1], y[2], y[3]);
    if (hi) {
        uint2 h, l;
        split4(y, h, l);
        *reinterpret_cast<uint2 *>(hi + pix * ld + c) = h;
        *reinterpret_cast<uint2 *>(lo + pix * ld + c) = l;
    }
}

}  // namespace rpe

extern "C" {

static int im2col7s2_impl(const void *img, int is_u8, void *out_hi, void *out_lo, int n, int H, int W, int ld, void *stream) {
    if (!img || !out_hi || (!out_lo && !is_u8) || n <= 0 || H <= 0 || W <= 0 || ld < 168 || (ld % 8)) return RPE_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(out_hi) & 15u) || (reinterpret_cast<uintptr_t>(out_lo) & 15u)) return RPE_ERR_ALIGNMENT;
    const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
    if (OH > 65535 || n > 65535 || ld != rpe::kStemLd) return RPE_ERR_INVALID_ARG;      // the staged copy assumes the 176-channel pitch
    dim3 grid((OW + rpe::kStemPx - 1) / rpe::kStemPx, OH, n);
    if (is_u8 && !out_lo)        // raw single-plane form (out_lo = NULL)
        rpe::im2col7s2_kernel<uint8_t, true><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)img, (rpe::plane_t *)out_hi, nullptr, H, W,
                                                                                     OH, OW, ld);
    else if (is_u8)
        rpe::im2col7s2_kernel<uint8_t, false><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)img, (rpe::plane_t *)out_hi,
                                                                                      (rpe::plane_t *)out_lo, H, W, OH, OW, ld);
    else
        rpe::im2col7s2_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)img, (rpe::plane_t *)out_hi,
                                                                                    (rpe::plane_t *)out_lo, H, W, OH, OW, ld);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_im2col7s2_split(const float *img, void *out_hi, void *out_lo, int n, int H, int W, int ld, void *stream) {
    return im2col7s2_impl(img, 0, out_hi, out_lo, n, H, W, ld, stream);
}

int rpe_im2col7s2_split_u8(const unsigned char *img, void *out_hi, void *out_lo, int n, int H, int W, int ld, void *stream) {
    return im2col7s2_impl(img, 1, out_hi, out_lo, n, H, W, ld, stream);
}

size_t rpe_instnorm_workspace_bytes(int n, int C) { return (size_t)n * rpe::kStatBlocks * C * 2 * sizeof(double); }

int rpe_instnorm_stats(const float *x, float *stats, int n, int HW, int C, float eps, void *workspace, size_t workspace_bytes,
                       void *stream) {
    if (!x || !stats || !workspace || n <= 0 || HW <= 0 || C <= 0 || (C % 4) || C > 1024) return RPE_ERR_INVALID_ARG;
    if (workspace_bytes < rpe_instnorm_workspace_bytes(n, C)) return RPE_ERR_WORKSPACE;
    if (!rpe::aligned16(x) || (reinterpret_cast<uintptr_t>(workspace) & 7u)) return RPE_ERR_ALIGNMENT;
    const int rows = 256 / (C / 4);
    if (rows < 1) return RPE_ERR_INVALID_ARG;
    const size_t smem = (size_t)rows * C * 2 * sizeof(double);
    dim3 grid(rpe::kStatBlocks, n);
    rpe::instnorm_partial_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, (double *)workspace, HW, C, C);
    RPE_LAUNCH_CHECK();
    rpe::instnorm_final_kernel<<<n, 128, 0, (cudaStream_t)stream>>>((const double *)workspace, stats, HW, C, eps);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_instnorm_stats_from_partials(const float *partials, float *stats, int n, int slots_per_image, int C, int ld, int HW, float eps,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    if (!partials || !stats || !workspace || n <= 0 || slots_per_image <= 0 || C <= 0 || C > 256 || ld < C || HW <= 0)
        return RPE_ERR_INVALID_ARG;
    if (workspace_bytes < rpe_instnorm_workspace_bytes(n, C)) return RPE_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(partials) & 7u) || (reinterpret_cast<uintptr_t>(workspace) & 7u)) return RPE_ERR_ALIGNMENT;
    dim3 grid(rpe::kStatBlocks, n);
    rpe::instnorm_from_partials_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(partials, (double *)workspace, slots_per_image, C, ld);
    RPE_LAUNCH_CHECK();
    rpe::instnorm_final_kernel<<<n, 128, 0, (cudaStream_t)stream>>>((const double *)workspace, stats, HW, C, eps);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_norm_act_split_res(const float *a, const float *stats_a, int relu_a, const float *b, const float *stats_b, const void *b_hi,
                           const void *b_lo, int b_ld, float *out_f32, void *out_hi, void *out_lo, int ld, int n, int HW, int C,
                           void *stream) {
    if (!a || n <= 0 || HW <= 0 || C <= 0 || (C % 4) || (!out_f32 && !out_hi) || ((out_hi == nullptr) != (out_lo == nullptr)))
        return RPE_ERR_INVALID_ARG;
    if (out_hi && (ld < C || (ld % 4))) return RPE_ERR_INVALID_ARG;
    if ((b_hi == nullptr) != (b_lo == nullptr) || (b_hi && (b || stats_b || b_ld < C || (b_ld % 4)))) return RPE_ERR_INVALID_ARG;
    if (!rpe::aligned16(a) || (b && !rpe::aligned16(b)) || (out_f32 && !rpe::aligned16(out_f32))) return RPE_ERR_ALIGNMENT;
    if ((reinterpret_cast<uintptr_t>(b_hi) & 7u) || (reinterpret_cast<uintptr_t>(b_lo) & 7u) || (reinterpret_cast<uintptr_t>(out_hi) & 7u) ||
        (reinterpret_cast<uintptr_t>(out_lo) & 7u))
        return RPE_ERR_ALIGNMENT;
    const long long total4 = (long long)n * HW * (C / 4);
    // 32-bit element offsets when every tensor has fewer than 2^32 elements (pix * max(C, ld, b_ld) below)
    const int max_ld = b_hi && b_ld > ld ? b_ld : ld;
    const long long elems = (long long)n * HW * (long long)(max_ld > C ? max_ld : C);
    if (elems + 1024 < (1ll << 32))
        rpe::norm_act_kernel<unsigned><<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            a, stats_a, relu_a, b, stats_b, (const rpe::plane_t *)b_hi, (const rpe::plane_t *)b_lo, b_ld, out_f32, (rpe::plane_t *)out_hi,
            (rpe::plane_t *)out_lo, ld, HW, C, total4);
    else
        rpe::norm_act_kernel<unsigned long long><<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            a, stats_a, relu_a, b, stats_b, (const rpe::plane_t *)b_hi, (const rpe::plane_t *)b_lo, b_ld, out_f32, (rpe::plane_t *)out_hi,
            (rpe::plane_t *)out_lo, ld, HW, C, total4);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_norm_act_split(const float *a, const float *stats_a, int relu_a, const float *b, const float *stats_b, float *out_f32,
                       void *out_hi, void *out_lo, int ld, int n, int HW, int C, void *stream) {
    return rpe_norm_act_split_res(a, stats_a, relu_a, b, stats_b, nullptr, nullptr, 0, out_f32, out_hi, out_lo, ld, n, HW, C, stream);
}

}  // extern "C"
