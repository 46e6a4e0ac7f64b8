// Implicit-GEMM 2-D convolution on tcgen05 tensor cores: the convolutional trunk of RAFT (SURVEY.md section 8f-1;
// reference: /root/reference/core/RAFT/core/update.py:79-136 update operator, core/RAFT/core/extractor.py:118-192 encoders,
// which the reference runs through cuDNN).
//
//   v[n, y, x, co]   = bias[co] + pre[n, y, x, co] + sum_s sum_(ky,kx) sum_ci  A_s[n, y*st+ky-ph, x*st+kx-pw, ci] * W_s[ky*kw+kx][co][ci]
//   out[n, y, x, co] = act(v) * scale                       (then  out = relu(out + res[n, y, x, co])  when a residual is given)
//
// Activations are NHWC bf16 planes.  The contraction runs over a LIST of sources (activation tensor + its weight slice), so
// concatenated inputs are never materialised, and every source may carry two planes hi = bf16(v), lo = bf16(v - hi): the
// "bf16x3" arithmetic evaluates hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (16 mantissa bits per operand; measured
// deviation from fp32 convolutions ~1e-5 relative).  Stride 1 or 2, "same"-style zero padding (k/2).
//
// Kernel: persistent, warp-specialised.  The output tile is 128 pixels = 16 (major) x 8 (minor) positions of one image; `orient`
// picks which image axis is the minor one.  Shared memory holds an activation SLAB of (16 + kmaj - 1) x 8 positions x 64
// channels per plane, laid out [major][minor][64 ch] with the 128-byte swizzle, so the UMMA descriptor of a filter tap that is
// shifted along the major axis is the same slab advanced by 1024 bytes: one TMA load serves all kmaj taps of a filter column
// (3x3: 3 loads instead of 9; 1x5 / 5x1: 1 load instead of 5).  Out-of-image coordinates are zero-filled by the TMA unit, which
// IS the zero padding.  The hi and lo planes of a stage are loaded once and feed all three products.  Weights stream through
// their own ring, one [bn][64] tile per tap.  Roles: warp 0 = activation TMA producer, warp 3 = weight TMA producer, warp 1 =
// tcgen05.mma issuer (one elected thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> bias / addend /
// activation / residual -> fp32 and/or bf16 hi/lo NHWC stores) overlapped with the next tile through two TMEM accumulator stages.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <math.h>
#include "common.cuh"
#include "ptx.cuh"

namespace rpe {

constexpr int kCvBK = 64;                             // bf16 channels per K block (one 128-byte swizzle row)
constexpr int kCvAcc = 2;                             // TMEM accumulator stages
constexpr int kCvThreads = 256;
constexpr int kCvMaxSrc = 4;
constexpr int kCvMaxBN = 256;
constexpr int kCvMaxAStages = 4, kCvMaxBStages = 8;
constexpr int kCvSmemData = 214 * 1024;               // budget for the two rings
constexpr int kCvMaxCout = 1024;                      // bias staged in shared memory
constexpr int kCvSmem = kCvSmemData + 1024 + 512 + kCvMaxCout * 4;     // + alignment slack + barriers + bias

struct alignas(64) ConvParams {
    CUtensorMap amap[kCvMaxSrc][2];
    CUtensorMap wmap[kCvMaxSrc][2];
    int cblocks[kCvMaxSrc];
    int ksteps_last[kCvMaxSrc];                       // 16-channel MMA steps of the last K block (1..4)
    int n_src, n_planes;
    int N, OH, OW, tiles_min, tiles_maj;
    int orient;                                       // 0: minor axis = x, 1: minor axis = y
    int kmin, kmaj, pad_min, pad_maj, stride, kw;
    int reuse;                                        // 1: one slab per filter column serves all kmaj taps (stride 1)
    uint32_t a_plane_bytes, a_stage_bytes, b_plane_bytes, b_stage_bytes;
    int n_a_stages, n_b_stages;
    int cout, bn, n_blocks;
    const float *bias;
    const float *pre;
    int pre_ld;
    const float *res;
    int res_ld;
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

// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t cv_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

template <int ACT>
__device__ __forceinline__ float cv_activate(float v) {
    if (ACT == 1) return fmaxf(v, 0.0f);
    if (ACT == 2) return 1.0f / (1.0f + expf(-v));
    if (ACT == 3) return tanhf(v);
    return v;
}

__device__ __forceinline__ void cv_store_bf16x4(const ConvParams &P, size_t o, const float *v) {
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = __float2bfloat16_rn(v[k]);
        l[k] = __float2bfloat16_rn(v[k] - __bfloat162float(h[k]));
    }
    *reinterpret_cast<uint2 *>(P.out_hi + o) = *reinterpret_cast<uint2 *>(h);
    if (P.out_lo) *reinterpret_cast<uint2 *>(P.out_lo + o) = *reinterpret_cast<uint2 *>(l);
}

__device__ __forceinline__ float cv_activate_rt(float v, int act) {
    return act == 1 ? cv_activate<1>(v) : act == 2 ? cv_activate<2>(v) : act == 3 ? cv_activate<3>(v) : v;
}

// Ragged tail (1..3 channels) of an output-channel count that is not a multiple of 4.
__device__ __noinline__ void cv_epilogue_tail(const ConvParams &P, int act, float a0, float a1, float a2, const float *sbias, int co,
                                              size_t pix) {
    const float acc[3] = {a0, a1, a2};
    for (int k = 0; k < 3 && co + k < P.cout; ++k) {
        float r = acc[k] + sbias[co + k];
        if (P.pre) r += __ldg(P.pre + pix * P.pre_ld + co + k);
        r = cv_activate_rt(r, act) * P.scale;
        if (P.res) r = fmaxf(r + __ldg(P.res + pix * P.res_ld + co + k), 0.0f);
        if (P.out_f32) P.out_f32[pix * P.f32_ld + P.f32_off + co + k] = r;
        if (P.out_hi) {
            const __nv_bfloat16 h = __float2bfloat16_rn(r);
            P.out_hi[pix * P.bf_ld + P.bf_off + co + k] = h;
            if (P.out_lo) P.out_lo[pix * P.bf_ld + P.bf_off + co + k] = __float2bfloat16_rn(r - __bfloat162float(h));
        }
    }
}

// One group of 4 consecutive output channels of one pixel: bias / addend / activation / scale / residual / stores.
template <int ACT>
__device__ __forceinline__ void cv_epilogue_group(const ConvParams &P, const uint32_t *acc4, const float *sbias, int co, size_t pix) {
    if (co >= P.cout) return;
    float o[4];
    if (co + 3 < P.cout) {
        const float4 b = *reinterpret_cast<const float4 *>(sbias + co);
        o[0] = __uint_as_float(acc4[0]) + b.x, o[1] = __uint_as_float(acc4[1]) + b.y;
        o[2] = __uint_as_float(acc4[2]) + b.z, o[3] = __uint_as_float(acc4[3]) + b.w;
        if (P.pre) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(P.pre + pix * P.pre_ld + co));
            o[0] += a.x, o[1] += a.y, o[2] += a.z, o[3] += a.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = cv_activate<ACT>(o[k]) * P.scale;
        if (P.res) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(P.res + pix * P.res_ld + co));
            o[0] = fmaxf(o[0] + a.x, 0.0f), o[1] = fmaxf(o[1] + a.y, 0.0f);
            o[2] = fmaxf(o[2] + a.z, 0.0f), o[3] = fmaxf(o[3] + a.w, 0.0f);
        }
        if (P.out_f32) *reinterpret_cast<float4 *>(P.out_f32 + pix * P.f32_ld + P.f32_off + co) = make_float4(o[0], o[1], o[2], o[3]);
        if (P.out_hi) cv_store_bf16x4(P, pix * P.bf_ld + P.bf_off + co, o);
    } else {                                           // ragged tail of a channel count that is not a multiple of 4
        cv_epilogue_tail(P, ACT, __uint_as_float(acc4[0]), __uint_as_float(acc4[1]), __uint_as_float(acc4[2]), sbias, co, pix);
    }
}

template <int ACT>
__device__ __forceinline__ void cv_epilogue_chunk(const ConvParams &P, const uint32_t *v, int cols, const float *sbias, int co0, size_t pix) {
#pragma unroll
    for (int j = 0; j < 32; j += 4)
        if (j < cols) cv_epilogue_group<ACT>(P, v + j, sbias, co0 + j, pix);
}

struct CvTile {
    int nb, img, omin0, omaj0;
};
__device__ __forceinline__ CvTile cv_decode(const ConvParams &P, int tile) {
    CvTile t;
    t.nb = tile % P.n_blocks;
    const int t2 = tile / P.n_blocks;
    const int per_img = P.tiles_min * P.tiles_maj;
    t.img = t2 / per_img;
    const int tr = t2 - t.img * per_img;
    const int tj = tr / P.tiles_min;
    t.omin0 = (tr - tj * P.tiles_min) * 8;
    t.omaj0 = tj * 16;
    return t;
}

__global__ void __launch_bounds__(kCvThreads, 1) conv_bf16_kernel(const __grid_constant__ ConvParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + (size_t)P.n_a_stages * P.a_stage_bytes;      // weight ring: n_b_stages entries of one plane tile
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kCvSmemData);
    uint64_t *a_full = bars;
    uint64_t *a_empty = a_full + kCvMaxAStages;
    uint64_t *b_full = a_empty + kCvMaxAStages;
    uint64_t *b_empty = b_full + kCvMaxBStages;
    uint64_t *tmem_full = b_empty + kCvMaxBStages;
    uint64_t *tmem_empty = tmem_full + kCvAcc;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + kCvAcc);
    float *sbias = reinterpret_cast<float *>(smem + kCvSmemData + 512);
    for (int i = threadIdx.x; i < P.bn * P.n_blocks; i += kCvThreads) sbias[i] = (P.bias != nullptr && i < P.cout) ? P.bias[i] : 0.0f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = P.N * P.tiles_min * P.tiles_maj * P.n_blocks;
    const int a_units = P.reuse ? 1 : P.kmaj;        // slabs per filter column

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P.n_src; ++s)
            for (int p = 0; p < P.n_planes; ++p) {
                prefetch_tmap(&P.amap[s][p]);
                prefetch_tmap(&P.wmap[s][p]);
            }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kCvMaxAStages; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < kCvMaxBStages; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
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
        // ===================== activation producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const CvTile t = cv_decode(P, tile);
                for (int s = 0; s < P.n_src; ++s)
                    for (int cb = 0; cb < P.cblocks[s]; ++cb)
                        for (int tm = 0; tm < P.kmin; ++tm)
                            for (int u = 0; u < a_units; ++u) {
                                mbar_wait(&a_empty[stage], phase ^ 1);
                                mbar_expect_tx(&a_full[stage], P.a_plane_bytes * P.n_planes);
                                const int cmin = t.omin0 * P.stride + tm - P.pad_min;
                                const int cmaj = t.omaj0 * P.stride + (P.reuse ? 0 : u) - P.pad_maj;
                                uint8_t *dst = sA + (size_t)stage * P.a_stage_bytes;
                                for (int p = 0; p < P.n_planes; ++p)
                                    tma_load_4d(dst + p * P.a_plane_bytes, &P.amap[s][p], &a_full[stage], cb * kCvBK, cmin, cmaj, t.img);
                                if (++stage == P.n_a_stages) stage = 0, phase ^= 1;
                            }
            }
        }
    } else if (warp == 3) {
        // ===================== weight producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const CvTile t = cv_decode(P, tile);
                for (int s = 0; s < P.n_src; ++s)
                    for (int cb = 0; cb < P.cblocks[s]; ++cb)
                        for (int tm = 0; tm < P.kmin; ++tm)
                            for (int tj = 0; tj < P.kmaj; ++tj) {
                                const int tap = P.orient == 0 ? tj * P.kw + tm : tm * P.kw + tj;
                                for (int p = 0; p < P.n_planes; ++p) {          // one ring entry per plane tile
                                    mbar_wait(&b_empty[stage], phase ^ 1);
                                    mbar_expect_tx(&b_full[stage], P.b_plane_bytes);
                                    tma_load_3d(sB + (size_t)stage * P.b_plane_bytes, &P.wmap[s][p], &b_full[stage], cb * kCvBK, t.nb * P.bn, tap);
                                    if (++stage == P.n_b_stages) stage = 0, phase ^= 1;
                                }
                            }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kCvMaxBN);
                uint32_t accumulate = 0;
                for (int s = 0; s < P.n_src; ++s)
                    for (int cb = 0; cb < P.cblocks[s]; ++cb) {
                        const int ksteps = (cb == P.cblocks[s] - 1) ? P.ksteps_last[s] : kCvBK / 16;
                        for (int tm = 0; tm < P.kmin; ++tm)
                            for (int tj = 0; tj < P.kmaj; ++tj) {
                                if (!P.reuse || tj == 0) {
                                    mbar_wait(&a_full[sa], pa);
                                }
                                mbar_wait(&b_full[sb], pb);
                                tcgen05_fence_after();
                                const uint32_t a0 = smem_u32(sA + (size_t)sa * P.a_stage_bytes) + (P.reuse ? (uint32_t)tj * 1024u : 0u);
                                const uint64_t da_hi = cv_sw128_desc(a0);
                                const uint64_t db_hi = cv_sw128_desc(smem_u32(sB + (size_t)sb * P.b_plane_bytes));
                                for (int k = 0; k < ksteps; ++k) {
                                    umma_bf16(d_tmem, da_hi + (uint64_t)(2 * k), db_hi + (uint64_t)(2 * k), idesc, accumulate);
                                    accumulate = 1;
                                }
                                if (P.n_planes == 2) {
                                    const uint64_t da_lo = cv_sw128_desc(a0 + P.a_plane_bytes);
                                    for (int k = 0; k < ksteps; ++k) umma_bf16(d_tmem, da_lo + (uint64_t)(2 * k), db_hi + (uint64_t)(2 * k), idesc, 1u);
                                    umma_commit(&b_empty[sb]);
                                    if (++sb == P.n_b_stages) sb = 0, pb ^= 1;
                                    mbar_wait(&b_full[sb], pb);
                                    tcgen05_fence_after();
                                    const uint64_t db_lo = cv_sw128_desc(smem_u32(sB + (size_t)sb * P.b_plane_bytes));
                                    for (int k = 0; k < ksteps; ++k) umma_bf16(d_tmem, da_hi + (uint64_t)(2 * k), db_lo + (uint64_t)(2 * k), idesc, 1u);
                                }
                                umma_commit(&b_empty[sb]);
                                if (++sb == P.n_b_stages) sb = 0, pb ^= 1;
                                if (!P.reuse || tj == P.kmaj - 1) {
                                    umma_commit(&a_empty[sa]);
                                    if (++sa == P.n_a_stages) sa = 0, pa ^= 1;
                                }
                            }
                    }
                umma_commit(&tmem_full[acc]);
                if (++acc == kCvAcc) acc = 0, acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int wq = warp - 4;
        const int m = wq * 32 + lane;                  // accumulator row = pixel inside the tile
        const int gi = m >> 3, mi = m & 7;             // (major, minor) position
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const CvTile t = cv_decode(P, tile);
            const int y = P.orient == 0 ? t.omaj0 + gi : t.omin0 + mi;
            const int x = P.orient == 0 ? t.omin0 + mi : t.omaj0 + gi;
            const bool inside = (y < P.OH) && (x < P.OW);
            const size_t pix = ((size_t)t.img * P.OH + y) * P.OW + x;
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
                const int co0 = t.nb * P.bn + c * 16;
                switch (P.act) {
                    case 1: cv_epilogue_chunk<1>(P, v, cols, sbias, co0, pix); break;
                    case 2: cv_epilogue_chunk<2>(P, v, cols, sbias, co0, pix); break;
                    case 3: cv_epilogue_chunk<3>(P, v, cols, sbias, co0, pix); break;
                    default: cv_epilogue_chunk<0>(P, v, cols, sbias, co0, pix); break;
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
    double flops;          // real multiply-adds x 2 of one run (all products of the split arithmetic)
};

}  // namespace rpe

extern "C" {

int rpe_conv_plan_create(const rpe_conv_desc *d, void **plan_out) {
    using namespace rpe;
    if (!d || !plan_out) return RPE_ERR_INVALID_ARG;
    if (d->n_sources < 1 || d->n_sources > kCvMaxSrc) return RPE_ERR_INVALID_ARG;
    if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->kh < 1 || d->kw < 1 || !(d->kh & 1) || !(d->kw & 1)) return RPE_ERR_INVALID_ARG;
    const int stride = d->stride <= 0 ? 1 : d->stride;
    if (stride != 1 && stride != 2) return RPE_ERR_INVALID_ARG;
    if (d->cout <= 0 || d->cout_pad < d->cout || d->cout_pad % 16 != 0 || d->cout_pad > kCvMaxCout) return RPE_ERR_INVALID_ARG;
    if (!d->out_f32 && !d->out_hi) return RPE_ERR_INVALID_ARG;
    if (d->out_lo && !d->out_hi) return RPE_ERR_INVALID_ARG;
    if (d->out_f32 && ((d->f32_ld % 4) || (d->f32_offset % 4) || !aligned16(d->out_f32))) return RPE_ERR_ALIGNMENT;
    if (d->out_hi && ((d->bf_ld % 4) || (d->bf_offset % 4) || (reinterpret_cast<uintptr_t>(d->out_hi) & 7u) ||
                      (reinterpret_cast<uintptr_t>(d->out_lo) & 7u)))
        return RPE_ERR_ALIGNMENT;
    if (d->pre && ((d->pre_ld % 4) || !aligned16(d->pre))) return RPE_ERR_ALIGNMENT;
    if (d->res && ((d->res_ld % 4) || !aligned16(d->res))) return RPE_ERR_ALIGNMENT;
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
    const int ph = d->kh / 2, pw = d->kw / 2;
    const int OH = (d->H + 2 * ph - d->kh) / stride + 1, OW = (d->W + 2 * pw - d->kw) / stride + 1;
    // orientation: the filter axis with more taps becomes the major axis (its taps share one slab); 3x3 keeps minor = x
    const int orient = (d->kw > d->kh) ? 1 : 0;
    p.orient = orient;
    p.kmin = orient == 0 ? d->kw : d->kh;
    p.kmaj = orient == 0 ? d->kh : d->kw;
    p.pad_min = orient == 0 ? pw : ph;
    p.pad_maj = orient == 0 ? ph : pw;
    p.stride = stride, p.kw = d->kw;
    p.reuse = (stride == 1) ? 1 : 0;
    const int slab_rows = p.reuse ? 16 + p.kmaj - 1 : 16;
    p.n_src = d->n_sources;
    p.n_planes = d->src[0].act_lo ? 2 : 1;
    p.a_plane_bytes = (uint32_t)slab_rows * 8 * 128;
    p.a_stage_bytes = p.a_plane_bytes * p.n_planes;
    p.b_plane_bytes = (uint32_t)bn * 128;
    p.b_stage_bytes = p.b_plane_bytes * p.n_planes;
    // ring depths: two activation stages (three when the weight tiles are small), every remaining byte goes to the weight
    // ring, whose entries are single plane tiles so that many small loads are in flight
    p.n_a_stages = 2;
    if (3 * (size_t)p.a_stage_bytes + 6 * (size_t)p.b_plane_bytes <= (size_t)kCvSmemData) p.n_a_stages = 3;
    p.n_b_stages = (int)((kCvSmemData - (size_t)p.n_a_stages * p.a_stage_bytes) / p.b_plane_bytes);
    if (p.n_b_stages > kCvMaxBStages) p.n_b_stages = kCvMaxBStages;
    if (p.n_b_stages < 2) {
        delete pl;
        return RPE_ERR_INVALID_ARG;
    }
    const int taps = d->kh * d->kw;
    double macs = 0.0;
    for (int s = 0; s < d->n_sources; ++s) {
        const rpe_conv_source &sc = d->src[s];
        if (!sc.act_hi || !sc.w_hi || ((sc.act_lo != nullptr) != (p.n_planes == 2)) || ((sc.w_lo != nullptr) != (p.n_planes == 2)) ||
            sc.c_count <= 0 || (sc.c_count % 16) || (sc.c_offset % 8) || (sc.c_total % 8) || sc.c_offset + sc.c_count > sc.c_total ||
            sc.w_cstride < sc.c_count || (sc.w_cstride % 8)) {
            delete pl;
            return RPE_ERR_INVALID_ARG;
        }
        p.cblocks[s] = (sc.c_count + kCvBK - 1) / kCvBK;
        p.ksteps_last[s] = (sc.c_count - (p.cblocks[s] - 1) * kCvBK) / 16;
        macs += (double)sc.c_count * taps;
        for (int pln = 0; pln < p.n_planes; ++pln) {
            const void *act = pln == 0 ? sc.act_hi : sc.act_lo;
            const void *wgt = pln == 0 ? sc.w_hi : sc.w_lo;
            if ((reinterpret_cast<uintptr_t>(act) & 15u) || (reinterpret_cast<uintptr_t>(wgt) & 15u)) {
                delete pl;
                return RPE_ERR_ALIGNMENT;
            }
            {   // activations: (C, minor, major, N) bf16; the channel window starts at c_offset, channels beyond it read as zero
                const char *base = reinterpret_cast<const char *>(act) + (size_t)sc.c_offset * 2;
                const cuuint64_t pix = (cuuint64_t)sc.c_total * 2, row = pix * d->W;
                cuuint64_t dims[4] = {(cuuint64_t)sc.c_count, (cuuint64_t)(orient == 0 ? d->W : d->H),
                                      (cuuint64_t)(orient == 0 ? d->H : d->W), (cuuint64_t)d->N};
                cuuint64_t strides[3] = {orient == 0 ? pix : row, orient == 0 ? row : pix, row * d->H};
                cuuint32_t box[4] = {kCvBK, (cuuint32_t)(8 * stride), (cuuint32_t)(slab_rows * stride), 1};
                cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
                CUresult r = g_cv_encode(&p.amap[s][pln], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char *>(base), dims, strides, box,
                                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    g_last_cuda_error = 100000 + (int)r;
                    delete pl;
                    return RPE_ERR_CUDA;
                }
            }
            {   // weights: (Cin_s, Cout_pad, taps) bf16 with row pitch w_cstride, box (64, bn, 1)
                cuuint64_t dims[3] = {(cuuint64_t)sc.c_count, (cuuint64_t)d->cout_pad, (cuuint64_t)taps};
                cuuint64_t strides[2] = {(cuuint64_t)sc.w_cstride * 2, (cuuint64_t)sc.w_cstride * 2 * d->cout_pad};
                cuuint32_t box[3] = {kCvBK, (cuuint32_t)bn, 1};
                cuuint32_t es[3] = {1, 1, 1};
                CUresult r = g_cv_encode(&p.wmap[s][pln], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(wgt), dims, strides, box,
                                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    g_last_cuda_error = 100000 + (int)r;
                    delete pl;
                    return RPE_ERR_CUDA;
                }
            }
        }
    }
    p.N = d->N, p.OH = OH, p.OW = OW;
    p.tiles_min = ((orient == 0 ? OW : OH) + 7) / 8;
    p.tiles_maj = ((orient == 0 ? OH : OW) + 15) / 16;
    p.cout = d->cout, p.bn = bn, p.n_blocks = n_blocks;
    p.bias = d->bias, p.act = d->activation, p.scale = d->out_scale;
    p.pre = d->pre, p.pre_ld = d->pre_ld, p.res = d->res, p.res_ld = d->res_ld;
    p.out_f32 = d->out_f32, p.f32_ld = d->f32_ld, p.f32_off = d->f32_offset;
    p.out_hi = reinterpret_cast<__nv_bfloat16 *>(d->out_hi), p.out_lo = reinterpret_cast<__nv_bfloat16 *>(d->out_lo);
    p.bf_ld = d->bf_ld, p.bf_off = d->bf_offset;
    pl->flops = 2.0 * macs * (double)d->cout * (double)d->N * OH * OW * (p.n_planes == 2 ? 3.0 : 1.0);
    const int tiles = p.N * p.tiles_min * p.tiles_maj * p.n_blocks;
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

double rpe_conv_plan_flops(void *plan) {
    if (!plan) return 0.0;
    return reinterpret_cast<rpe::ConvPlan *>(plan)->flops;
}

int rpe_conv_plan_destroy(void *plan) {
    if (!plan) return RPE_ERR_INVALID_ARG;
    delete reinterpret_cast<rpe::ConvPlan *>(plan);
    return RPE_OK;
}

}  // extern "C"
