// RAFT convex 8x flow up-sampling (reference: /root/reference/core/RAFT/core/raft.py:66-77).
//   out[b, c, 8y+i, 8x+j] = sum_k softmax_k(mask[b, k*64 + i*8 + j, y, x]) * 8 * flow[b, c, y+ky-1, x+kx-1]
// with k = 3*ky + kx over the zero-padded 3x3 neighbourhood (F.unfold(8*flow, 3, padding=1)).
// CTA = one coarse row y, 8 coarse columns: the 576 x 8 mask tile (18 KB, so ~10 CTAs per SM keep the loads of one tile
// behind the arithmetic of the others; the first version staged 32 columns = 76 KB and ran at 2 CTAs per SM, 0.18 of HBM) is
// staged through shared memory with coalesced reads; thread = (cell, pair of fine rows, fine column).
#include "common.cuh"

namespace rpe {

constexpr int kUpCells = 8;

__global__ void __launch_bounds__(256) convex_upsample8_kernel(const float *__restrict__ flow, const float *__restrict__ mask,
                                                               float *__restrict__ out, int h, int w, int mask_nhwc_ld) {
    extern __shared__ float sm[];
    float *s_mask = sm;                                  // [576][33]
    float *s_flow = sm + 576 * (kUpCells + 1);           // [2][3][kUpCells + 2]
    const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * kUpCells;
    const int hw = h * w;
    const int ncell = min(kUpCells, w - x0);
    if (mask_nhwc_ld > 0) {               // mask (B,h,w,ld) channels-last (tcgen05 convolution path)
        for (int e = threadIdx.x; e < 576 * kUpCells; e += blockDim.x) {
            const int xl = e / 576, ch = e - xl * 576;
            s_mask[ch * (kUpCells + 1) + xl] = xl < ncell ? __ldg(mask + ((size_t)b * hw + y * w + x0 + xl) * mask_nhwc_ld + ch) : 0.0f;
        }
    } else {
        for (int e = threadIdx.x; e < 576 * kUpCells; e += blockDim.x) {
            const int ch = e / kUpCells, xl = e - ch * kUpCells;
            s_mask[ch * (kUpCells + 1) + xl] = xl < ncell ? __ldg(mask + ((size_t)b * 576 + ch) * hw + y * w + x0 + xl) : 0.0f;
        }
    }
    for (int e = threadIdx.x; e < 2 * 3 * (kUpCells + 2); e += blockDim.x) {
        const int c = e / (3 * (kUpCells + 2));
        const int r = (e / (kUpCells + 2)) % 3;
        const int xl = e % (kUpCells + 2);
        const int yy = y + r - 1, xx = x0 + xl - 1;
        float v = 0.0f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = 8.0f * __ldg(flow + ((size_t)b * 2 + c) * hw + yy * w + xx);
        s_flow[e] = v;
    }
    __syncthreads();
    const int xl = threadIdx.x >> 5, i0 = ((threadIdx.x >> 3) & 3) * 2, j = threadIdx.x & 7;
    if (xl >= ncell) return;
    const int W8 = 8 * w;
#pragma unroll
    for (int i = i0; i < i0 + 2; ++i) {
        float m[9];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            m[k] = s_mask[(k * 64 + i * 8 + j) * (kUpCells + 1) + xl];
            mx = fmaxf(mx, m[k]);
        }
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            m[k] = expf(m[k] - mx);
            sum += m[k];
        }
        float ox = 0.0f, oy = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int ky = k / 3, kx = k - 3 * ky;
            const float wgt = m[k] / sum;
            ox += wgt * s_flow[(0 * 3 + ky) * (kUpCells + 2) + xl + kx];
            oy += wgt * s_flow[(1 * 3 + ky) * (kUpCells + 2) + xl + kx];
        }
        const size_t o = (size_t)(8 * y + i) * W8 + 8 * (x0 + xl) + j;
        out[((size_t)b * 2 + 0) * 64 * hw + o] = ox;
        out[((size_t)b * 2 + 1) * 64 * hw + o] = oy;
    }
}

}  // namespace rpe

static int convex_upsample_impl(const float *flow, const float *mask, float *out, int B, int h, int w, int nhwc_ld, void *stream);

extern "C" int rpe_convex_upsample8(const float *flow, const float *mask, float *out, int B, int h, int w, void *stream) {
    return convex_upsample_impl(flow, mask, out, B, h, w, 0, stream);
}

extern "C" int rpe_convex_upsample8_nhwc(const float *flow, const float *mask, int mask_ld, float *out, int B, int h, int w, void *stream) {
    if (mask_ld < 576) return RPE_ERR_INVALID_ARG;
    return convex_upsample_impl(flow, mask, out, B, h, w, mask_ld, stream);
}

static int convex_upsample_impl(const float *flow, const float *mask, float *out, int B, int h, int w, int nhwc_ld, void *stream) {
    using namespace rpe;
    if (!flow || !mask || !out || B <= 0 || h <= 0 || w <= 0) return RPE_ERR_INVALID_ARG;
    const size_t smem = (576 * (kUpCells + 1) + 2 * 3 * (kUpCells + 2)) * sizeof(float);
    RPE_CUDA_TRY(cudaFuncSetAttribute(convex_upsample8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
    dim3 grid((w + kUpCells - 1) / kUpCells, h, B);
    convex_upsample8_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(flow, mask, out, h, w, nhwc_ld);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}
