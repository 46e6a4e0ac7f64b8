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

// ------------------------------------------------------------------------------------------------
// ResizeStereo on the device (reference dataset/transforms.py:20-39): resize (aspect ratio kept) + centre crop in one pass.
// Bilinear with anti-aliasing restates ATen's _upsample_bilinear2d_aa (what torchvision.transforms.functional.resize runs on
// float tensors): per output index i, scale = in / out, support = max(scale, 1), centre = scale * (i + 0.5),
// taps [xmin, xmin + xsize) = [max(0, int(centre - support + 0.5)), min(in, int(centre + support + 0.5))), weight of tap j =
// max(0, 1 - |(j + xmin - centre + 0.5) / max(scale, 1)|) / sum.  Horizontal pass first (per source row), then vertical.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void aa_taps(int i, float scale, int in_size, int &xmin, int &xsize, float &invs, float &center) {
    const float support = scale >= 1.0f ? scale : 1.0f;
    invs = scale >= 1.0f ? 1.0f / scale : 1.0f;
    center = scale * ((float)i + 0.5f);
    xmin = max(0, (int)(center - support + 0.5f));
    xsize = min(in_size, (int)(center + support + 0.5f)) - xmin;
}
__device__ __forceinline__ float aa_weight(int j, int xmin, float center, float invs) {
    const float x = fabsf(((float)(j + xmin) - center + 0.5f) * invs);
    return x < 1.0f ? 1.0f - x : 0.0f;
}

template <typename T>
__global__ void __launch_bounds__(256) resize_crop_aa_kernel(const T *__restrict__ src, float *__restrict__ dst, int C, int Hi, int Wi, int rh,
                                                             int rw, int top, int left, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nc = blockIdx.z;
    if (x >= W) return;
    const float sy = (float)Hi / (float)rh, sx = (float)Wi / (float)rw;
    int ymin, ysize, xmin, xsize;
    float yinv, ycen, xinv, xcen;
    aa_taps(y + top, sy, Hi, ymin, ysize, yinv, ycen);
    aa_taps(x + left, sx, Wi, xmin, xsize, xinv, xcen);
    float wxsum = 0.0f, wysum = 0.0f;
    for (int j = 0; j < xsize; ++j) wxsum += aa_weight(j, xmin, xcen, xinv);
    for (int j = 0; j < ysize; ++j) wysum += aa_weight(j, ymin, ycen, yinv);
    const T *base = src + (size_t)nc * Hi * Wi;
    float acc = 0.0f;
    for (int r = 0; r < ysize; ++r) {
        const T *row = base + (size_t)(ymin + r) * Wi + xmin;
        float t = 0.0f;
        for (int j = 0; j < xsize; ++j) t += (float)__ldg(row + j) * (aa_weight(j, xmin, xcen, xinv) / wxsum);
        acc += t * (aa_weight(r, ymin, ycen, yinv) / wysum);
    }
    dst[((size_t)nc * H + y) * W + x] = acc;
}

template <typename T>
__global__ void __launch_bounds__(256) resize_crop_nearest_kernel(const T *__restrict__ src, uint8_t *__restrict__ dst, int Hi, int Wi, int rh,
                                                                  int rw, int top, int left, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nc = blockIdx.z;
    if (x >= W) return;
    const float sy = (float)Hi / (float)rh, sx = (float)Wi / (float)rw;
    const int ys = min((int)floorf((float)(y + top) * sy), Hi - 1), xs = min((int)floorf((float)(x + left) * sx), Wi - 1);
    dst[((size_t)nc * H + y) * W + x] = __ldg(src + ((size_t)nc * Hi + ys) * Wi + xs) != (T)0 ? 1 : 0;
}

}  // namespace rpe

extern "C" {

int rpe_resize_crop(const void *src, int src_u8, void *dst, int n, int C, int Hi, int Wi, int rh, int rw, int top, int left, int H, int W,
                    int mode, void *stream) {
    if (!src || !dst || n <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || rh <= 0 || rw <= 0 || H <= 0 || W <= 0) return RPE_ERR_INVALID_ARG;
    if (top < 0 || left < 0 || top + H > rh || left + W > rw || (mode != 0 && mode != 1)) return RPE_ERR_INVALID_ARG;
    if ((long long)n * C > 65535 || H > 65535) return RPE_ERR_INVALID_ARG;
    dim3 grid((W + 255) / 256, H, n * C);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        if (src_u8) rpe::resize_crop_aa_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)src, (float *)dst, C, Hi, Wi, rh, rw, top, left, H, W);
        else rpe::resize_crop_aa_kernel<float><<<grid, 256, 0, st>>>((const float *)src, (float *)dst, C, Hi, Wi, rh, rw, top, left, H, W);
    } else {
        if (src_u8) rpe::resize_crop_nearest_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)src, (uint8_t *)dst, Hi, Wi, rh, rw, top, left, H, W);
        else rpe::resize_crop_nearest_kernel<float><<<grid, 256, 0, st>>>((const float *)src, (uint8_t *)dst, Hi, Wi, rh, rw, top, left, H, W);
    }
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}


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
