// Input-pipeline row of the scope table: the specularity mask of the reference's datasets,
// /root/reference/dataset/stereo_dataset.py:12-16 (also dataset/video_dataset.py, tum_dataset.py call sites):
//     spec_mask = img.sum(axis=-1) < 3 * 255 * spec_thr ;  mask &= spec_mask ;  mask = cv2.erode(mask, ones((11, 11)))
// evaluated on the uint8 frame that is already on the device (the tracker receives uint8 frames, bench.py e2e), so the mask
// never has to be computed on the host.  Byte work, bit-exact: integer channel sum against an integer limit, then a
// (2r+1) x (2r+1) minimum with out-of-image pixels ignored (cv2.erode's default border value).
//
// CTA = 64 x 16 output pixels.  The (16 + 2r) x (64 + 2r) neighbourhood of validity bytes is built in shared memory with
// coalesced byte loads of the three colour planes, reduced along x, then along y.
#include "common.cuh"

namespace rpe {

constexpr int kMsTW = 64, kMsTH = 16, kMsMaxR = 8;

__global__ void __launch_bounds__(256) mask_specularities_kernel(const uint8_t *__restrict__ img, const uint8_t *__restrict__ mask_in,
                                                                 uint8_t *__restrict__ mask_out, int H, int W, int max_sum, int r) {
    __shared__ uint8_t s_v[kMsTH + 2 * kMsMaxR][kMsTW + 2 * kMsMaxR];
    __shared__ uint8_t s_h[kMsTH + 2 * kMsMaxR][kMsTW];
    const int n = blockIdx.z, x0 = blockIdx.x * kMsTW, y0 = blockIdx.y * kMsTH;
    const size_t plane = (size_t)H * W;
    const uint8_t *im = img + (size_t)n * 3 * plane;
    const uint8_t *mi = mask_in ? mask_in + (size_t)n * plane : nullptr;
    const int tw = kMsTW + 2 * r, th = kMsTH + 2 * r;
    for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
        const int ty = i / tw, tx = i - ty * tw;
        const int y = y0 + ty - r, x = x0 + tx - r;
        uint8_t v = 1;                                           // outside the image: does not constrain the minimum
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const size_t o = (size_t)y * W + x;
            const int sum = (int)im[o] + (int)im[plane + o] + (int)im[2 * plane + o];
            v = (sum <= max_sum && (!mi || mi[o] != 0)) ? 1 : 0;
        }
        s_v[ty][tx] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < th * kMsTW; i += blockDim.x) {
        const int ty = i / kMsTW, tx = i - ty * kMsTW;
        uint8_t m = 1;
        for (int k = 0; k <= 2 * r; ++k) m &= s_v[ty][tx + k];
        s_h[ty][tx] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMsTH * kMsTW; i += blockDim.x) {
        const int ty = i / kMsTW, tx = i - ty * kMsTW;
        const int y = y0 + ty, x = x0 + tx;
        if (y >= H || x >= W) continue;
        uint8_t m = 1;
        for (int k = 0; k <= 2 * r; ++k) m &= s_h[ty + k][tx];
        mask_out[(size_t)n * plane + (size_t)y * W + x] = m;
    }
}

}  // namespace rpe

extern "C" {

int rpe_mask_specularities(const uint8_t *img, const uint8_t *mask_in, uint8_t *mask_out, int n, int H, int W, int max_sum, int radius,
                           void *stream) {
    if (!img || !mask_out || n <= 0 || H <= 0 || W <= 0 || radius < 0 || radius > rpe::kMsMaxR || n > 65535) return RPE_ERR_INVALID_ARG;
    if (mask_out == mask_in) return RPE_ERR_INVALID_ARG;          // neighbourhood reads: not an in-place operation
    dim3 grid((W + rpe::kMsTW - 1) / rpe::kMsTW, (H + rpe::kMsTH - 1) / rpe::kMsTH, n);
    if (grid.y > 65535) return RPE_ERR_INVALID_ARG;
    rpe::mask_specularities_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, mask_in, mask_out, H, W, max_sum, radius);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

}  // extern "C"
