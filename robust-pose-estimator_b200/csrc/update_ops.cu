// Element-wise companions of the tcgen05 convolution path of the RAFT update operator
// (reference: /root/reference/core/RAFT/core/update.py:33-60 SepConvGRU gating, :79-97 motion encoder inputs,
// core/RAFT/core/raft.py:112-121 coordinate update).  All tensors NHWC; "split" = fp16 hi/lo planes (conv.cu).
#include <cuda_fp16.h>
#include "common.cuh"

namespace rpe {

__device__ __forceinline__ void split_store(float v, plane_t *hi, plane_t *lo, size_t o) {
    const plane_t h = to_plane(v);
    hi[o] = h;
    lo[o] = to_plane_lo(v - plane_to_float(h));
}

// NCHW fp32 (n,C,H,W) -> NHWC split planes at channel offset `off` of a tensor with `ld` channels (+ optional fp32 NHWC copy).
__global__ void __launch_bounds__(256) nchw_to_nhwc_split_kernel(const float *__restrict__ x, plane_t *__restrict__ hi,
                                                                 plane_t *__restrict__ lo, float *__restrict__ f32, int C,
                                                                 int HW, int ld, int off, int f32_ld, int f32_off) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + ty + 8 * j, p = p0 + tx;
        tile[ty + 8 * j][tx] = (c < C && p < HW) ? __ldg(x + ((size_t)n * C + c) * HW + p) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = p0 + ty + 8 * j, c = c0 + tx;
        if (p < HW && c < C) {
            const float v = tile[tx][ty + 8 * j];
            const size_t pix = (size_t)n * HW + p;
            if (hi) split_store(v, hi, lo, pix * ld + off + c);
            if (f32) f32[pix * f32_ld + f32_off + c] = v;
        }
    }
}

// NHWC fp32 (window of C channels at offset) -> NCHW fp32.
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float *__restrict__ x, float *__restrict__ out, int C, int HW, int ld,
                                                           int off) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = p0 + ty + 8 * j, c = c0 + tx;
        tile[ty + 8 * j][tx] = (c < C && p < HW) ? __ldg(x + ((size_t)n * HW + p) * ld + off + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + ty + 8 * j, p = p0 + tx;
        if (p < HW && c < C) out[((size_t)n * C + c) * HW + p] = tile[tx][ty + 8 * j];
    }
}

// coords1 += delta (NHWC fp32, 2 of `d_ld` channels): raft.py:121.  Separate launch: the im2col below reads neighbours.
__global__ void __launch_bounds__(256) coords_add_kernel(float *__restrict__ coords1, const float *__restrict__ delta, int d_ld, int hw) {
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hw) return;
    float *c1 = coords1 + (size_t)n * 2 * hw;
    const float *d = delta + ((size_t)n * hw + p) * d_ld;
    c1[p] = c1[p] + d[0];
    c1[hw + p] = c1[hw + p] + d[1];
}

// flow = coords1 - coords0 with coords0 = pixel grid; writes (i) the 7x7 im2col of the flow (98 of `col_ld` channels,
// tap-major: ch = (ky*7+kx)*2 + c) as split planes and (ii) the flow itself into 2 channels at `x_off` of the GRU input
// tensor.  One thread = one (pixel, tap): consecutive threads write consecutive 4-byte (fx, fy) pairs.
// idx_t = unsigned for every realistic size (32-bit divisions; the 64-bit ones made this kernel instruction-bound).
template <typename idx_t>
__global__ void __launch_bounds__(256) flow_im2col_kernel(const float *__restrict__ coords1, plane_t *__restrict__ col_hi,
                                                          plane_t *__restrict__ col_lo, int col_ld, plane_t *__restrict__ x_hi,
                                                          plane_t *__restrict__ x_lo, int x_ld, int x_off, int h, int w,
                                                          long long total) {
    const idx_t i = (idx_t)blockIdx.x * (idx_t)blockDim.x + threadIdx.x;
    if ((long long)i >= total) return;
    const int tap = (int)(i % 49);
    const idx_t pix = i / 49;
    const int hw = h * w;
    const int n = (int)(pix / (idx_t)hw);
    const int p = (int)(pix - (idx_t)n * (idx_t)hw);
    const int y = p / w, x = p - y * w;
    const int ky = tap / 7, kx = tap - ky * 7;
    const int yy = y + ky - 3, xx = x + kx - 3;
    float fx = 0.0f, fy = 0.0f;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const float *c1 = coords1 + (size_t)n * 2 * hw;
        fx = __ldg(c1 + yy * w + xx) - (float)xx;
        fy = __ldg(c1 + hw + yy * w + xx) - (float)yy;
    }
    const plane_t hx = to_plane(fx), hy = to_plane(fy);
    const plane2_t hi2 = __halves2half2(hx, hy);
    const plane2_t lo2 = __halves2half2(to_plane_lo(fx - plane_to_float(hx)), to_plane_lo(fy - plane_to_float(hy)));
    const size_t o = (size_t)pix * col_ld + tap * 2;
    *reinterpret_cast<plane2_t *>(col_hi + o) = hi2;
    *reinterpret_cast<plane2_t *>(col_lo + o) = lo2;
    if (tap == 24) {
        *reinterpret_cast<plane2_t *>(x_hi + (size_t)pix * x_ld + x_off) = hi2;
        *reinterpret_cast<plane2_t *>(x_lo + (size_t)pix * x_ld + x_off) = lo2;
    }
}

// GRU gating.  zr: fp32 NHWC (.., 256) = [z | r] after sigmoid; h: fp32 NHWC (.., 128).
//   mode 0: rh = r * h                  -> split planes (conv input of the candidate state)
//   mode 1: h  = (1 - z) * h + z * q    -> fp32 h (in place) + split planes
__global__ void __launch_bounds__(256) gru_gate_kernel(const float *__restrict__ zr, float *__restrict__ h, const float *__restrict__ q,
                                                       plane_t *__restrict__ o_hi, plane_t *__restrict__ o_lo, int o_ld,
                                                       int o_off, size_t npix, int mode) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 4 channels of one pixel
    if (i >= npix * 32) return;
    const size_t pix = i >> 5;
    const int c = (int)(i & 31) * 4;
    const float4 hv = *reinterpret_cast<const float4 *>(h + pix * 128 + c);
    float o[4];
    if (mode == 0) {
        const float4 r = *reinterpret_cast<const float4 *>(zr + pix * 256 + 128 + c);
        o[0] = r.x * hv.x, o[1] = r.y * hv.y, o[2] = r.z * hv.z, o[3] = r.w * hv.w;
    } else {
        const float4 z = *reinterpret_cast<const float4 *>(zr + pix * 256 + c);
        const float4 qv = *reinterpret_cast<const float4 *>(q + pix * 128 + c);
        o[0] = (1.0f - z.x) * hv.x + z.x * qv.x;
        o[1] = (1.0f - z.y) * hv.y + z.y * qv.y;
        o[2] = (1.0f - z.z) * hv.z + z.z * qv.z;
        o[3] = (1.0f - z.w) * hv.w + z.w * qv.w;
        *reinterpret_cast<float4 *>(h + pix * 128 + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
    plane_t hh[4], ll[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        hh[k] = to_plane(o[k]);
        ll[k] = to_plane_lo(o[k] - plane_to_float(hh[k]));
    }
    *reinterpret_cast<uint2 *>(o_hi + pix * o_ld + o_off + c) = *reinterpret_cast<uint2 *>(hh);
    *reinterpret_cast<uint2 *>(o_lo + pix * o_ld + o_off + c) = *reinterpret_cast<uint2 *>(ll);
}

// Second half of a 3x3 convolution whose per-pixel part ran in the epilogue of the preceding convolution (conv.cu mode 3):
// part (n,h,w,ld) holds two slots (one per half of the contracted channels) of p[t*2 + o] = <y(pixel), W[o][:, t]>, t = ky*3 + kx;
// out[y][x][o] = bias[o] + sum_t p_t[y + ky - 1][x + kx - 1][o] with zero padding (update.py:6-13, FlowHead.conv2).
__global__ void __launch_bounds__(256) tap_gather3x3_kernel(const float *__restrict__ part, int ld, const float *__restrict__ bias,
                                                            float *__restrict__ out, int out_ld, int h, int w, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const long long img = i / ((long long)w * h);
    float a0 = bias ? __ldg(bias) : 0.0f, a1 = bias ? __ldg(bias + 1) : 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1;
            if (xx < 0 || xx >= w) continue;
            const float *p = part + ((img * h + yy) * w + xx) * ld + (ky * 3 + kx) * 2;
            const float2 u = __ldg(reinterpret_cast<const float2 *>(p)), v = __ldg(reinterpret_cast<const float2 *>(p + 18));
            a0 += u.x + v.x;
            a1 += u.y + v.y;
        }
    }
    *reinterpret_cast<float2 *>(out + i * out_ld) = make_float2(a0, a1);
}

}  // namespace rpe

extern "C" {

int rpe_nchw_to_nhwc_split(const float *x, void *hi, void *lo, float *f32, int n, int C, int H, int W, int ld, int off, int f32_ld,
                           int f32_off, void *stream) {
    if (!x || (!hi && !f32) || ((hi == nullptr) != (lo == nullptr)) || n <= 0 || C <= 0 || H <= 0 || W <= 0) return RPE_ERR_INVALID_ARG;
    const int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, n);
    rpe::nchw_to_nhwc_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (rpe::plane_t *)hi, (rpe::plane_t *)lo, f32, C, HW, ld, off,
                                                                          f32_ld, f32_off);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_nhwc_to_nchw(const float *x, float *out, int n, int C, int H, int W, int ld, int off, void *stream) {
    if (!x || !out || n <= 0 || C <= 0 || H <= 0 || W <= 0) return RPE_ERR_INVALID_ARG;
    const int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, n);
    rpe::nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, C, HW, ld, off);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_flow_step(float *coords1, const float *delta, int delta_ld, void *col_hi, void *col_lo, int col_ld, void *x_hi, void *x_lo,
                  int x_ld, int x_off, int n, int h, int w, void *stream) {
    if (!coords1 || !col_hi || !col_lo || !x_hi || !x_lo || n <= 0 || h <= 0 || w <= 0 || col_ld < 98 || (col_ld % 2) || (x_ld % 2) ||
        (x_off % 2))
        return RPE_ERR_INVALID_ARG;
    if (delta) {
        dim3 grid((h * w + 255) / 256, n);
        rpe::coords_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(coords1, delta, delta_ld, h * w);
        RPE_LAUNCH_CHECK();
    }
    const long long total = (long long)n * h * w * 49;
    if (total + 256 < (1ll << 32))
        rpe::flow_im2col_kernel<unsigned><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            coords1, (rpe::plane_t *)col_hi, (rpe::plane_t *)col_lo, col_ld, (rpe::plane_t *)x_hi, (rpe::plane_t *)x_lo, x_ld, x_off, h, w,
            total);
    else
        rpe::flow_im2col_kernel<unsigned long long><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            coords1, (rpe::plane_t *)col_hi, (rpe::plane_t *)col_lo, col_ld, (rpe::plane_t *)x_hi, (rpe::plane_t *)x_lo, x_ld, x_off, h, w,
            total);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_tap_gather3x3(const float *part, int part_ld, const float *bias, float *out, int out_ld, int n, int h, int w, void *stream) {
    if (!part || !out || n <= 0 || h <= 0 || w <= 0 || part_ld < 36 || (part_ld % 2) || (out_ld % 2) || out_ld < 2 ||
        (reinterpret_cast<uintptr_t>(part) & 7u) || (reinterpret_cast<uintptr_t>(out) & 7u))
        return RPE_ERR_INVALID_ARG;
    const long long total = (long long)n * h * w;
    rpe::tap_gather3x3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part, part_ld, bias, out, out_ld, h, w, total);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_gru_gate(const float *zr, float *h, const float *q, void *out_hi, void *out_lo, int out_ld, int out_off, long long npix,
                 int mode, void *stream) {
    if (!zr || !h || !out_hi || !out_lo || npix <= 0 || (mode == 1 && !q) || (out_ld % 4) || (out_off % 4)) return RPE_ERR_INVALID_ARG;
    const size_t threads = (size_t)npix * 32;
    rpe::gru_gate_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(zr, h, q, (rpe::plane_t *)out_hi,
                                                                                             (rpe::plane_t *)out_lo, out_ld, out_off,
                                                                                             (size_t)npix, mode);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

}  // extern "C"
