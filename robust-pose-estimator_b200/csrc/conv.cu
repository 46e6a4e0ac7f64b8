// Implicit-GEMM 2-D convolution on tcgen05 tensor cores for the RAFT update operator.
//
//   out[n, y, x, co] = act( bias[co] + sum_s sum_(dy,dx) sum_ci  A_s[n, y+dy, x+dx, ci] * W_s[tap][co][ci] ) * scale
//
// Activations are NHWC bf16 planes; the contraction runs over a LIST of (activation plane, weight) sources, which
// gives both concatenation-free multi-input convolutions (cat(h, x) never materialised) and the error-compensated
// "bf16x3" arithmetic: a value v is stored as two planes hi = bf16(v), lo = bf16(v - hi) and a product is evaluated as
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (16 mantissa bits per operand; measured pose deviation vs fp32
// convolutions < 1e-5).  Stride 1, "same" zero padding (all update-block convolutions).
//
// Kernel: persistent, warp-specialised like corr_gemm_kernel.  M tile = 8x16 output pixels of one image (128 TMEM lanes),
// N tile = up to 256 output channels, K block = 64 bf16 channels of one tap of one source.  The A tile of a tap is a
// TMA box of the 4-D activation tensor map at (c, x0+dx, y0+dy, n): out-of-image coordinates are zero-filled by the TMA
// unit, which IS the zero padding.  Epilogue: TMEM -> registers -> bias / activation -> fp32 and/or bf16 hi/lo NHWC.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <math.h>
#include "common.cuh"
#include "ptx.cuh"

namespace rpe {

constexpr int kCvTH = 8, kCvTW = 16;                  // output tile: 8 rows x 16 cols = 128 pixels
constexpr int kCvBK = 64;                             // bf16 channels per K block (128-byte swizzle row)
constexpr int kCvStages = 4, kCvAcc = 2;
constexpr int kCvThreads = 256;
constexpr int kCvABytes = 128 * kCvBK * 2;            // 16 KB
constexpr int kCvMaxSrc = 8;
constexpr int kCvMaxBN = 256;
constexpr int kCvBBytesMax = kCvMaxBN * kCvBK * 2;    // 32 KB
constexpr int kCvSmem = kCvStages * (kCvABytes + kCvBBytesMax) + 1024 + 256;

struct alignas(64) ConvParams {
    CUtensorMap amap[kCvMaxSrc];
    CUtensorMap wmap[kCvMaxSrc];
    int cblocks[kCvMaxSrc];
    int n_src;
    int N, H, W, tiles_x, tiles_y;
    int kh, kw;
    int cout, bn, n_blocks;
    const float *bias;
    int act;
    float scale;
    float *out_f32;
    int f32_ld, f32_off;
    __nv_bfloat16 *out_hi, *out_lo;
    int bf_ld, bf_off;
};

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint64_t cv_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__device__ __forceinline__ float cv_activate(float v, int act) {
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) return 1.0f / (1.0f + expf(-v));
    if (act == 3) return tanhf(v);
    return v;
}

__global__ void __launch_bounds__(kCvThreads, 1) conv_bf16_kernel(const __grid_constant__ ConvParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + kCvStages * kCvABytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(sB + kCvStages * kCvBBytesMax);
    uint64_t *empty_bar = full_bar + kCvStages;
    uint64_t *tmem_full = empty_bar + kCvStages;
    uint64_t *tmem_empty = tmem_full + kCvAcc;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + kCvAcc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int num_tiles = P.N * tiles_per_img * P.n_blocks;
    const int taps = P.kh * P.kw;
    const int ph = P.kh / 2, pw = P.kw / 2;
    const uint32_t b_bytes = (uint32_t)P.bn * kCvBK * 2;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P.n_src; ++s) {
            prefetch_tmap(&P.amap[s]);
            prefetch_tmap(&P.wmap[s]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kCvStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kCvAcc; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nb = tile % P.n_blocks;
                const int t2 = tile / P.n_blocks;
                const int img = t2 / tiles_per_img;
                const int tr = t2 - img * tiles_per_img;
                const int ty = tr / P.tiles_x, tx = tr - ty * P.tiles_x;
                const int y0 = ty * kCvTH, x0 = tx * kCvTW;
                for (int s = 0; s < P.n_src; ++s) {
                    for (int tap = 0; tap < taps; ++tap) {
                        const int dy = tap / P.kw - ph, dx = tap % P.kw - pw;
                        for (int cb = 0; cb < P.cblocks[s]; ++cb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            mbar_expect_tx(&full_bar[stage], kCvABytes + b_bytes);
                            tma_load_4d(sA + stage * kCvABytes, &P.amap[s], &full_bar[stage], cb * kCvBK, x0 + dx, y0 + dy, img);
                            tma_load_3d(sB + stage * kCvBBytesMax, &P.wmap[s], &full_bar[stage], cb * kCvBK, nb * P.bn, tap);
                            if (++stage == kCvStages) stage = 0, phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // instruction descriptor: D fp32, A/B bf16, K-major, N = bn, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int total_kb = 0;
            for (int s = 0; s < P.n_src; ++s) total_kb += taps * P.cblocks[s];
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kCvMaxBN);
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint64_t da = cv_sw128_desc(smem_u32(sA + stage * kCvABytes));
                    const uint64_t db = cv_sw128_desc(smem_u32(sB + stage * kCvBBytesMax));
#pragma unroll
                    for (int k = 0; k < kCvBK / 16; ++k)
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == total_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == kCvStages) stage = 0, phase ^= 1;
                }
                if (++acc == kCvAcc) acc = 0, acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================== Epilogue =====================
        const int wq = warp - 4;
        const int p = wq * 32 + lane;                  // pixel inside the tile
        const int py = p / kCvTW, px = p - py * kCvTW;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int nb = tile % P.n_blocks;
            const int t2 = tile / P.n_blocks;
            const int img = t2 / tiles_per_img;
            const int tr = t2 - img * tiles_per_img;
            const int ty = tr / P.tiles_x, tx = tr - ty * P.tiles_x;
            const int y = ty * kCvTH + py, x = tx * kCvTW + px;
            const bool inside = (y < P.H) && (x < P.W);
            const size_t pix = ((size_t)img * P.H + y) * P.W + x;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int n_chunks = P.bn / 16;
            for (int c = 0; c < n_chunks; c += 2) {
                // 32 columns per TMEM load when available, 16 for the tail of a bn that is not a multiple of 32
                uint32_t v[32];
                const int cols = (c + 1 < n_chunks) ? 32 : 16;
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * kCvMaxBN + c * 16);
                if (cols == 32) {
                    tmem_ld_32x32b_x32(taddr, v);
                } else {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr)
                        : "memory");
                }
                tmem_ld_wait();
                if (c + 2 >= n_chunks) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                if (!inside) continue;
                const int co0 = nb * P.bn + c * 16;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (j >= cols) break;
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int co = co0 + j + k;
                        const float b = (P.bias != nullptr && co < P.cout) ? __ldg(P.bias + co) : 0.0f;
                        o[k] = cv_activate(__uint_as_float(v[j + k]) + b, P.act) * P.scale;
                    }
                    const int co = co0 + j;
                    if (co + 3 < P.cout) {
                        if (P.out_f32)
                            *reinterpret_cast<float4 *>(P.out_f32 + pix * P.f32_ld + P.f32_off + co) = make_float4(o[0], o[1], o[2], o[3]);
                        if (P.out_hi) {
                            __nv_bfloat16 h[4], l[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                h[k] = __float2bfloat16_rn(o[k]);
                                l[k] = __float2bfloat16_rn(o[k] - __bfloat162float(h[k]));
                            }
                            *reinterpret_cast<uint2 *>(P.out_hi + pix * P.bf_ld + P.bf_off + co) = *reinterpret_cast<uint2 *>(h);
                            *reinterpret_cast<uint2 *>(P.out_lo + pix * P.bf_ld + P.bf_off + co) = *reinterpret_cast<uint2 *>(l);
                        }
                    } else {
                        for (int k = 0; k < 4 && co + k < P.cout; ++k) {
                            if (P.out_f32) P.out_f32[pix * P.f32_ld + P.f32_off + co + k] = o[k];
                            if (P.out_hi) {
                                const __nv_bfloat16 h = __float2bfloat16_rn(o[k]);
                                P.out_hi[pix * P.bf_ld + P.bf_off + co + k] = h;
                                P.out_lo[pix * P.bf_ld + P.bf_off + co + k] = __float2bfloat16_rn(o[k] - __bfloat162float(h));
                            }
                        }
                    }
                }
            }
            if (++acc == kCvAcc) acc = 0, acc_phase ^= 1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side: plans
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_cv_encode = nullptr;

static int cv_load_encode() {
    if (g_cv_encode) return RPE_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return cuda_fail(e == cudaSuccess ? cudaErrorUnknown : e);
    g_cv_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return RPE_OK;
}

struct ConvPlan {
    ConvParams p;
    int grid;
};

}  // namespace rpe

extern "C" {

int rpe_conv_plan_create(const rpe_conv_desc *d, void **plan_out) {
    using namespace rpe;
    if (!d || !plan_out) return RPE_ERR_INVALID_ARG;
    if (d->n_sources < 1 || d->n_sources > kCvMaxSrc) return RPE_ERR_INVALID_ARG;
    if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->kh < 1 || d->kw < 1 || !(d->kh & 1) || !(d->kw & 1)) return RPE_ERR_INVALID_ARG;
    if (d->cout <= 0 || d->cout_pad < d->cout || d->cout_pad % 16 != 0) return RPE_ERR_INVALID_ARG;
    if (!d->out_f32 && !d->out_hi) return RPE_ERR_INVALID_ARG;
    if ((d->out_hi == nullptr) != (d->out_lo == nullptr)) return RPE_ERR_INVALID_ARG;
    if (d->out_f32 && ((d->f32_ld % 4) || (d->f32_offset % 4) || !aligned16(d->out_f32))) return RPE_ERR_ALIGNMENT;
    if (d->out_hi && ((d->bf_ld % 4) || (d->bf_offset % 4) || (reinterpret_cast<uintptr_t>(d->out_hi) & 7u) ||
                      (reinterpret_cast<uintptr_t>(d->out_lo) & 7u)))
        return RPE_ERR_ALIGNMENT;
    int rc = cv_load_encode();
    if (rc != RPE_OK) return rc;
    ConvPlan *pl = new ConvPlan();
    ConvParams &p = pl->p;
    // N tiling: one block if cout_pad <= 256, else equal blocks of <= 256 that are multiples of 16
    int n_blocks = (d->cout_pad + kCvMaxBN - 1) / kCvMaxBN;
    int bn = d->cout_pad / n_blocks;
    if (bn * n_blocks != d->cout_pad || bn % 16 != 0) {
        delete pl;
        return RPE_ERR_INVALID_ARG;
    }
    p.n_src = d->n_sources;
    const int taps = d->kh * d->kw;
    for (int s = 0; s < d->n_sources; ++s) {
        const rpe_conv_source &sc = d->src[s];
        if (!sc.act || !sc.weight || sc.c_count <= 0 || (sc.c_count % kCvBK) || (sc.c_offset % 8) || (sc.c_total % 8) ||
            sc.c_offset + sc.c_count > sc.c_total || (reinterpret_cast<uintptr_t>(sc.act) & 15u) ||
            (reinterpret_cast<uintptr_t>(sc.weight) & 15u)) {
            delete pl;
            return RPE_ERR_INVALID_ARG;
        }
        p.cblocks[s] = sc.c_count / kCvBK;
        {   // activations: (C, W, H, N) bf16, box (64, 16, 8, 1); the channel window starts at c_offset
            const char *base = reinterpret_cast<const char *>(sc.act) + (size_t)sc.c_offset * 2;
            cuuint64_t dims[4] = {(cuuint64_t)sc.c_count, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
            cuuint64_t strides[3] = {(cuuint64_t)sc.c_total * 2, (cuuint64_t)sc.c_total * 2 * d->W,
                                     (cuuint64_t)sc.c_total * 2 * d->W * d->H};
            cuuint32_t box[4] = {kCvBK, kCvTW, kCvTH, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUresult r = g_cv_encode(&p.amap[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char *>(base), dims, strides, box, es,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                g_last_cuda_error = 100000 + (int)r;
                delete pl;
                return RPE_ERR_CUDA;
            }
        }
        {   // weights: (Cin_s, Cout_pad, taps) bf16, box (64, bn, 1)
            cuuint64_t dims[3] = {(cuuint64_t)sc.c_count, (cuuint64_t)d->cout_pad, (cuuint64_t)taps};
            cuuint64_t strides[2] = {(cuuint64_t)sc.c_count * 2, (cuuint64_t)sc.c_count * 2 * d->cout_pad};
            cuuint32_t box[3] = {kCvBK, (cuuint32_t)bn, 1};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = g_cv_encode(&p.wmap[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(sc.weight), dims, strides, box,
                                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                g_last_cuda_error = 100000 + (int)r;
                delete pl;
                return RPE_ERR_CUDA;
            }
        }
    }
    p.N = d->N, p.H = d->H, p.W = d->W;
    p.tiles_x = (d->W + kCvTW - 1) / kCvTW;
    p.tiles_y = (d->H + kCvTH - 1) / kCvTH;
    p.kh = d->kh, p.kw = d->kw;
    p.cout = d->cout, p.bn = bn, p.n_blocks = n_blocks;
    p.bias = d->bias, p.act = d->activation, p.scale = d->out_scale;
    p.out_f32 = d->out_f32, p.f32_ld = d->f32_ld, p.f32_off = d->f32_offset;
    p.out_hi = reinterpret_cast<__nv_bfloat16 *>(d->out_hi), p.out_lo = reinterpret_cast<__nv_bfloat16 *>(d->out_lo);
    p.bf_ld = d->bf_ld, p.bf_off = d->bf_offset;
    const int tiles = p.N * p.tiles_x * p.tiles_y * p.n_blocks;
    pl->grid = sm_count() < tiles ? sm_count() : tiles;
    if (pl->grid < 1) pl->grid = 1;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmem);
        if (e != cudaSuccess) {
            delete pl;
            return cuda_fail(e);
        }
        attr = true;
    }
    *plan_out = pl;
    return RPE_OK;
}

int rpe_conv_plan_run(void *plan, void *stream) {
    using namespace rpe;
    if (!plan) return RPE_ERR_INVALID_ARG;
    ConvPlan *pl = reinterpret_cast<ConvPlan *>(plan);
    conv_bf16_kernel<<<pl->grid, kCvThreads, kCvSmem, (cudaStream_t)stream>>>(pl->p);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_conv_plan_destroy(void *plan) {
    if (!plan) return RPE_ERR_INVALID_ARG;
    delete reinterpret_cast<rpe::ConvPlan *>(plan);
    return RPE_OK;
}

}  // extern "C"
