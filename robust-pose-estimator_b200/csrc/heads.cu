// Companions of the confidence heads: TinyUNet (/root/reference/core/unet/unet.py:8-82) on the tcgen05
// convolution kernel.  Every un-padded 3x3 convolution of the reference runs as a "same" convolution on the grid of its input
// (csrc/conv.cu); the reference's output is the interior of that result, so each tensor carries a VALID REGION (offset + size
// inside its grid) that shrinks by one pixel per convolution.  The kernels below move data between those grids:
//
//   downsample8_planes      the 1/8 bilinear down-sampling of pose_net.py:110-113 (= mean of the centre 2x2 of every 8x8 block,
//                           SURVEY.md A.3) of up to three NCHW tensors, written as NHWC fp16 split planes (head input channels)
//   pool2_planes            F.max_pool2d(x, 2) of the valid region of an fp32 NHWC tensor -> compact split planes
//   upcat_planes            decoder input: ConvTranspose2d(k = 2, s = 2) evaluated as a 1x1 convolution with 4 x C output
//                           channels (one group per (dy, dx)), un-shuffled here, concatenated with the centre crop of the skip
//                           tensor (unet.py:52-61)
//   resize_sigmoid          F.interpolate(logits, (H, W), mode='bilinear') (align_corners=False) + Sigmoid (unet.py:71-77,
//                           pose_net.py:24-27) of the valid region of the 1-channel head output -> (n,1,H,W) fp32
#include "common.cuh"

namespace rpe {

__device__ __forceinline__ void hd_split_store(float v, plane_t *hi, plane_t *lo, size_t o) {
    const plane_t h = to_plane(v);
    hi[o] = h;
    lo[o] = to_plane_lo(v - plane_to_float(h));
}

struct HdCatSrc {
    const float *src[3];
    int ch[3];
    int u8_mask;          // bit s: source s holds uint8 values
};

__global__ void __launch_bounds__(256) downsample8_planes_kernel(HdCatSrc t, plane_t *__restrict__ hi, plane_t *__restrict__ lo, int ld,
                                                                 int ch_offset, int H, int W) {
    const int h8 = H / 8, w8 = W / 8;
    const int ctot = t.ch[0] + t.ch[1] + t.ch[2];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (idx >= ctot * h8 * w8) return;
    const int cc = idx % ctot;                                  // channel fastest: contiguous NHWC stores
    const int j = (idx / ctot) % w8, i = idx / (ctot * w8);
    int c = cc, s = 0;
    if (c >= t.ch[0]) c -= t.ch[0], s = 1;
    if (s == 1 && c >= t.ch[1]) c -= t.ch[1], s = 2;
    const size_t o = ((size_t)b * t.ch[s] + c) * H * W + (size_t)(8 * i + 3) * W + 8 * j + 3;
    float2 a, d;
    if ((t.u8_mask >> s) & 1) {
        const uint8_t *p = reinterpret_cast<const uint8_t *>(t.src[s]) + o;
        a = make_float2((float)__ldg(p), (float)__ldg(p + 1));
        d = make_float2((float)__ldg(p + W), (float)__ldg(p + W + 1));
    } else {
        const float *p = t.src[s] + o;
        a = make_float2(__ldg(p), __ldg(p + 1));
        d = make_float2(__ldg(p + W), __ldg(p + W + 1));
    }
    // ATen upsample_bilinear2d, align_corners=False, scale 8: source index 8i + 3.5 -> weights 0.5 / 0.5 on both axes
    const float top = __fadd_rn(__fmul_rn(0.5f, a.x), __fmul_rn(0.5f, a.y));
    const float bot = __fadd_rn(__fmul_rn(0.5f, d.x), __fmul_rn(0.5f, d.y));
    const float v = __fadd_rn(__fmul_rn(0.5f, top), __fmul_rn(0.5f, bot));
    hd_split_store(v, hi, lo, (((size_t)b * h8 + i) * w8 + j) * ld + ch_offset + cc);
}

// x: fp32 NHWC (n, H, W, ld_in), valid region rows [y0, y0 + 2 oh), columns [x0, x0 + 2 ow) -> planes (n, oh, ow, ld_out)
__global__ void __launch_bounds__(256) pool2_planes_kernel(const float *__restrict__ x, int H, int W, int ld_in, int c_in_off, int y0, int x0,
                                                           plane_t *__restrict__ hi, plane_t *__restrict__ lo, int oh, int ow, int ld_out, int C,
                                                           long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int j = (int)(r % ow);
    r /= ow;
    const int i = (int)(r % oh), b = (int)(r / oh);
    const float *p = x + (((size_t)b * H + y0 + 2 * i) * W + x0 + 2 * j) * ld_in + c_in_off + c;
    const float v = fmaxf(fmaxf(__ldg(p), __ldg(p + ld_in)), fmaxf(__ldg(p + (size_t)W * ld_in), __ldg(p + (size_t)(W + 1) * ld_in)));
    hd_split_store(v, hi, lo, (((size_t)b * oh + i) * ow + j) * ld_out + c);
}

// up:   fp32 NHWC (n, Hu, Wu, ld_u): 4 groups of c_up channels, group (dy * 2 + dx); valid region offset (uy0, ux0)
// skip: fp32 NHWC (n, Hk, Wk, ld_k) channels [k_off, k_off + c_skip), crop origin (ky0, kx0)
// out:  planes (n, oh, ow, ld_out): [0, c_up) = transposed-conv output, [c_up, c_up + c_skip) = cropped skip
__global__ void __launch_bounds__(256) upcat_planes_kernel(const float *__restrict__ up, int Hu, int Wu, int ld_u, int uy0, int ux0, int c_up,
                                                           const float *__restrict__ skip, int Hk, int Wk, int ld_k, int k_off, int ky0, int kx0,
                                                           int c_skip, plane_t *__restrict__ hi, plane_t *__restrict__ lo, int oh, int ow,
                                                           int ld_out, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int C = c_up + c_skip;
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int x = (int)(r % ow);
    r /= ow;
    const int y = (int)(r % oh), b = (int)(r / oh);
    float v;
    if (c < c_up)
        v = __ldg(up + (((size_t)b * Hu + uy0 + (y >> 1)) * Wu + ux0 + (x >> 1)) * ld_u + (((y & 1) << 1) | (x & 1)) * c_up + c);
    else
        v = __ldg(skip + (((size_t)b * Hk + ky0 + y) * Wk + kx0 + x) * ld_k + k_off + (c - c_up));
    hd_split_store(v, hi, lo, (((size_t)b * oh + y) * ow + x) * ld_out + c);
}

// ATen upsample_bilinear2d (align_corners = False): src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out (fp32),
// i1 = min(i0 + 1, in - 1), value = (1 - ly) * ((1 - lx) v00 + lx v01) + ly * ((1 - lx) v10 + lx v11); then 1 / (1 + exp(-v)).
__global__ void __launch_bounds__(256) resize_sigmoid_kernel(const float *__restrict__ logits, int Hl, int Wl, int ld, int ch, int y0, int x0,
                                                             int ih, int iw, float *__restrict__ out, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const float sy = (float)ih / (float)H, sx = (float)iw / (float)W;
    float fy = sy * ((float)y + 0.5f) - 0.5f, fx = sx * ((float)x + 0.5f) - 0.5f;
    fy = fy < 0.0f ? 0.0f : fy;
    fx = fx < 0.0f ? 0.0f : fx;
    const int iy0 = (int)fy, ix0 = (int)fx;
    const int iy1 = iy0 + (iy0 < ih - 1 ? 1 : 0), ix1 = ix0 + (ix0 < iw - 1 ? 1 : 0);
    const float ly = fy - (float)iy0, lx = fx - (float)ix0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const float *base = logits + ((size_t)b * Hl * Wl) * ld + ch;
    const float v00 = __ldg(base + ((size_t)(y0 + iy0) * Wl + x0 + ix0) * ld), v01 = __ldg(base + ((size_t)(y0 + iy0) * Wl + x0 + ix1) * ld);
    const float v10 = __ldg(base + ((size_t)(y0 + iy1) * Wl + x0 + ix0) * ld), v11 = __ldg(base + ((size_t)(y0 + iy1) * Wl + x0 + ix1) * ld);
    const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
    out[((size_t)b * H + y) * W + x] = 1.0f / (1.0f + expf(-v));
}

}  // namespace rpe

extern "C" {

int rpe_downsample8_planes(const void *src0, int c0, const void *src1, int c1, const void *src2, int c2, int src_u8_mask, void *out_hi,
                           void *out_lo, int ld, int ch_offset, int n, int H, int W, void *stream) {
    if (!out_hi || !out_lo || n <= 0 || H < 8 || W < 8 || (H % 8) || (W % 8)) return RPE_ERR_INVALID_ARG;
    rpe::HdCatSrc t;
    t.src[0] = (const float *)src0, t.ch[0] = src0 ? c0 : 0;
    t.src[1] = (const float *)src1, t.ch[1] = src1 ? c1 : 0;
    t.src[2] = (const float *)src2, t.ch[2] = src2 ? c2 : 0;
    t.u8_mask = src_u8_mask;
    const int ctot = t.ch[0] + t.ch[1] + t.ch[2];
    if (ctot <= 0 || ch_offset < 0 || ch_offset + ctot > ld) return RPE_ERR_INVALID_ARG;
    const int total = ctot * (H / 8) * (W / 8);
    dim3 grid((total + 255) / 256, n);
    rpe::downsample8_planes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, (rpe::plane_t *)out_hi, (rpe::plane_t *)out_lo, ld, ch_offset, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_pool2_planes(const float *x, int H, int W, int ld_in, int c_in_off, int y0, int x0, void *out_hi, void *out_lo, int oh, int ow,
                     int ld_out, int C, int n, void *stream) {
    if (!x || !out_hi || !out_lo || n <= 0 || C <= 0 || oh <= 0 || ow <= 0 || y0 < 0 || x0 < 0 || y0 + 2 * oh > H || x0 + 2 * ow > W ||
        c_in_off < 0 || c_in_off + C > ld_in || C > ld_out)
        return RPE_ERR_INVALID_ARG;
    const long long total = (long long)n * oh * ow * C;
    rpe::pool2_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, H, W, ld_in, c_in_off, y0, x0, (rpe::plane_t *)out_hi, (rpe::plane_t *)out_lo, oh, ow, ld_out, C, total);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_upcat_planes(const float *up, int Hu, int Wu, int ld_u, int uy0, int ux0, int c_up, const float *skip, int Hk, int Wk, int ld_k,
                     int k_off, int ky0, int kx0, int c_skip, void *out_hi, void *out_lo, int oh, int ow, int ld_out, int n, void *stream) {
    if (!up || !skip || !out_hi || !out_lo || n <= 0 || c_up <= 0 || c_skip <= 0 || oh <= 0 || ow <= 0) return RPE_ERR_INVALID_ARG;
    if (uy0 < 0 || ux0 < 0 || uy0 + (oh + 1) / 2 > Hu || ux0 + (ow + 1) / 2 > Wu || 4 * c_up > ld_u) return RPE_ERR_INVALID_ARG;
    if (ky0 < 0 || kx0 < 0 || ky0 + oh > Hk || kx0 + ow > Wk || k_off < 0 || k_off + c_skip > ld_k || c_up + c_skip > ld_out)
        return RPE_ERR_INVALID_ARG;
    const long long total = (long long)n * oh * ow * (c_up + c_skip);
    rpe::upcat_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        up, Hu, Wu, ld_u, uy0, ux0, c_up, skip, Hk, Wk, ld_k, k_off, ky0, kx0, c_skip, (rpe::plane_t *)out_hi, (rpe::plane_t *)out_lo, oh, ow,
        ld_out, total);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_resize_sigmoid(const float *logits, int Hl, int Wl, int ld, int ch, int y0, int x0, int ih, int iw, float *out, int n, int H, int W,
                       void *stream) {
    if (!logits || !out || n <= 0 || H <= 0 || W <= 0 || ih <= 0 || iw <= 0 || y0 < 0 || x0 < 0 || y0 + ih > Hl || x0 + iw > Wl || ch < 0 ||
        ch >= ld)
        return RPE_ERR_INVALID_ARG;
    dim3 grid((W + 255) / 256, H, n);
    rpe::resize_sigmoid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, Hl, Wl, ld, ch, y0, x0, ih, iw, out, H, W);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

}  // extern "C"
