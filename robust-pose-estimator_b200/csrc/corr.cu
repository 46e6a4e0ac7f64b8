// Stage 1: RAFT CorrBlock on sm_100a.
//
//   corr_prep_kernel     fmap (B,C,Q) fp32  ->  K-major tf32 operand (B,Q,Kp)   [transpose + cvt.rna.tf32,
//                        optional hi/lo split for the 3xTF32 mode]
//   corr_gemm_kernel     level0[b,q,t] = <f1[b,:,q], f2[b,:,t]> / sqrt(C): persistent, warp-specialised
//                        tcgen05.mma (kind::tf32, M=128, N=256, accumulators in TMEM), operands staged by
//                        TMA (SWIZZLE_128B, 4-stage mbarrier ring), double-buffered TMEM accumulators,
//                        epilogue TMEM -> registers -> swizzled smem -> TMA store.
//   corr_pool_kernel     levels 1..3 (2x2 means) from level 0 in one pass per query slab.
//   corr_lookup_kernel   radius-r bilinear window lookup: warp-cooperative patch loads, smem transposition
//                        so that both the gather and the (B, L*(2r+1)^2, h, w) output are coalesced.
//
// Reference semantics: /root/reference/core/RAFT/core/corr.py:12-60, utils/utils.py:57-71.
#include <cstdio>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "ptx.cuh"

namespace rpe {

// ================================================================================================
// Operand preparation: (B, C, Q) fp32 -> (B, Q, Kp) tf32-rounded, K contiguous.
//   which = 0 (A side, fmap1): [hi | lo | hi]      which = 1 (B side, fmap2): [hi | hi | lo]
// so that the K = 3C contraction yields hi*hi + lo*hi + hi*lo.  Single-pass mode writes [hi] only.
// ================================================================================================
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(256) corr_prep_kernel(const float *__restrict__ fmap, float *__restrict__ out, int C, int Q,
                                                        int split, int which) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int q0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int Kp = split ? 3 * C : C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + ty + 8 * j, q = q0 + tx;
        tile[ty + 8 * j][tx] = (c < C && q < Q) ? __ldg(fmap + ((size_t)b * C + c) * Q + q) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int q = q0 + ty + 8 * j, c = c0 + tx;
        if (q < Q && c < C) {
            const float x = tile[tx][ty + 8 * j];
            const float hi = to_tf32(x);
            float *o = out + ((size_t)b * Q + q) * Kp;
            o[c] = hi;
            if (split) {
                const float lo = to_tf32(x - hi);
                o[C + c] = which == 0 ? lo : hi;
                o[2 * C + c] = which == 0 ? hi : lo;
            }
        }
    }
}

// ================================================================================================
// tcgen05 / TMA GEMM
// ================================================================================================
constexpr int kBM = 128, kBN = 256, kBK = 32;          // tile (elements); kBK * 4 B = one 128-byte swizzle row
constexpr int kStages = 4, kAccStages = 2;
constexpr int kEpiCols = 32;                            // columns per epilogue chunk / TMA store box
constexpr int kGemmThreads = 256;                       // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int kABytes = kBM * kBK * 4, kBBytes = kBN * kBK * 4;
constexpr int kStgBytes = kBM * kEpiCols * 4;
constexpr int kGemmSmem = kStages * (kABytes + kBBytes) + 2 * kStgBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t kTmemCols = kAccStages * kBN;        // 512

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6), a/b_format TF32=2 [7,10)/[10,13),
// a/b K-major (0), n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

struct GemmShape {
    int B, Q, Kp;      // batch, queries (= targets), padded contraction length
    int m_tiles, n_tiles;
    float scale;       // 1 / sqrt(C)
};

__global__ void __launch_bounds__(kGemmThreads, 1)
    corr_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_c, GemmShape s) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + kStages * kABytes;
    uint8_t *sStg = sB + kStages * kBBytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(sStg + 2 * kStgBytes);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *tmem_full = empty_bar + kStages;
    uint64_t *tmem_empty = tmem_full + kAccStages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = s.B * s.m_tiles * s.n_tiles;
    constexpr int kBKe = kBK;                            // elements per 128-byte K block
    const int num_kb = s.Kp / kBKe;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);      // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, kTmemCols);
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
                const int b = tile / (s.m_tiles * s.n_tiles);
                const int r = tile - b * (s.m_tiles * s.n_tiles);
                const int n_blk = r / s.m_tiles, m_blk = r - n_blk * s.m_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], kABytes + kBBytes);
                    tma_load_3d(sA + stage * kABytes, &tmap_a, &full_bar[stage], kb * kBKe, m_blk * kBM, b);
                    tma_load_3d(sB + stage * kBBytes, &tmap_b, &full_bar[stage], kb * kBKe, n_blk * kBN, b);
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kBN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint64_t da = make_sw128_desc(smem_u32(sA + stage * kABytes));
                    const uint64_t db = make_sw128_desc(smem_u32(sB + stage * kBBytes));
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {
                        // advance 8 tf32 = 32 bytes along K inside the swizzle atom: +2 in the (>>4) start address
                        umma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kInstrDesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);                 // frees the smem stage when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
                if (++acc == kAccStages) acc = 0, acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================== Epilogue: TMEM -> regs -> swizzled smem -> TMA store =====================
        const int wq = warp - 4;                       // TMEM lane quarter == warp id % 4
        const int row = wq * 32 + lane;                // row inside the 128-row tile
        const bool leader = (threadIdx.x == 128);
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t chunk_no = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / (s.m_tiles * s.n_tiles);
            const int r = tile - b * (s.m_tiles * s.n_tiles);
            const int n_blk = r / s.m_tiles, m_blk = r - n_blk * s.m_tiles;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            for (int c = 0; c < kBN / kEpiCols; ++c, ++chunk_no) {
                uint8_t *stg = sStg + (chunk_no & 1u) * kStgBytes;
                if (leader) tma_store_wait_read<1>();          // the store that last read this buffer has drained
                named_bar_sync(1, 128);
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * kBN + c * kEpiCols), v);
                tmem_ld_wait();
                if (c == kBN / kEpiCols - 1) {
                    // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                float4 *rowp = reinterpret_cast<float4 *>(stg + row * 128);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o;
                    o.x = __uint_as_float(v[4 * j + 0]) * s.scale;
                    o.y = __uint_as_float(v[4 * j + 1]) * s.scale;
                    o.z = __uint_as_float(v[4 * j + 2]) * s.scale;
                    o.w = __uint_as_float(v[4 * j + 3]) * s.scale;
                    rowp[j ^ (row & 7)] = o;                       // SWIZZLE_128B: 16-byte chunk index XOR (row % 8)
                }
                fence_proxy_async_smem();
                named_bar_sync(2, 128);
                if (leader) {
                    tma_store_3d(&tmap_c, stg, n_blk * kBN + c * kEpiCols, m_blk * kBM, b);
                    tma_store_commit();
                }
            }
            if (++acc == kAccStages) acc = 0, acc_phase ^= 1;
        }
        if (leader) tma_store_wait_read<0>();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ================================================================================================
// Pyramid pooling: one CTA per query slab; levels 1..3 from level 0 through shared memory.
// avg_pool2d(2, 2): ((a00 + a01) + a10 + a11) * 0.25, floor on odd sizes.
// ================================================================================================
struct PyrDims {
    int h[4], w[4];
    size_t off[4];      // element offsets of the levels inside the pyramid buffer
    int levels;
};

__global__ void __launch_bounds__(256) corr_pool_kernel(float *__restrict__ pyr, PyrDims d) {
    extern __shared__ float sm[];
    const size_t slab = blockIdx.x;                      // b * Q + q
    const int n0 = d.h[0] * d.w[0];
    const float *src = pyr + d.off[0] + slab * n0;
    float *cur = sm;
    for (int i = threadIdx.x * 4; i < n0; i += blockDim.x * 4) {
        if (i + 3 < n0 && ((n0 & 3) == 0)) {
            *reinterpret_cast<float4 *>(cur + i) = __ldg(reinterpret_cast<const float4 *>(src + i));
        } else {
            for (int k = i; k < n0 && k < i + 4; ++k) cur[k] = __ldg(src + k);
        }
    }
    __syncthreads();
    float *nxt = sm + n0;
    for (int l = 1; l < d.levels; ++l) {
        const int hp = d.h[l - 1], wp = d.w[l - 1], hl = d.h[l], wl = d.w[l];
        (void)hp;
        float *dst = pyr + d.off[l] + slab * (size_t)(hl * wl);
        for (int i = threadIdx.x; i < hl * wl; i += blockDim.x) {
            const int y = i / wl, x = i - y * wl;
            const float *p = cur + (2 * y) * wp + 2 * x;
            const float v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[wp]), p[wp + 1]), 0.25f);
            nxt[i] = v;
            dst[i] = v;
        }
        __syncthreads();
        cur = nxt;
        nxt = cur + hl * wl;
    }
}

// ================================================================================================
// Window lookup.  CTA = 8 warps = 32 consecutive queries of one batch element.
//   phase 1 (per warp, 4 queries): for each level the (2r+2)^2 integer neighbourhood of the sample
//            centre is fetched with lanes spread over (row, column) -> a handful of 128-byte lines per
//            request; bilinear weights are shared by the whole window (integer offsets).
//   phase 2: the (L*(2r+1)^2) x 32 result tile is written channel-major, 128 contiguous bytes per
//            channel row.
// ================================================================================================
constexpr int kLookupQ = 32;
constexpr int kMaxWin = 10;       // 2r+2 for r = 4

__global__ void __launch_bounds__(256) corr_lookup_kernel(const float *__restrict__ pyr, PyrDims d, const float *__restrict__ coords,
                                                          float *__restrict__ out, int Q, int radius,
                                                          plane_t *__restrict__ out_hi, plane_t *__restrict__ out_lo,
                                                          int nhwc_ld) {
    extern __shared__ float sm[];
    const int n = 2 * radius + 1, win = n + 1;
    const int nch = d.levels * n * n;
    float *s_out = sm;                                   // [nch][kLookupQ + 1]
    float *s_patch = sm + nch * (kLookupQ + 1);          // [8 warps][win*win]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int q_base = blockIdx.x * kLookupQ;
    float *patch = s_patch + warp * (kMaxWin * kMaxWin);

    for (int qi = warp; qi < kLookupQ; qi += 8) {
        const int q = q_base + qi;
        if (q >= Q) break;
        const float cx = __ldg(coords + ((size_t)b * 2 + 0) * Q + q);
        const float cy = __ldg(coords + ((size_t)b * 2 + 1) * Q + q);
        for (int l = 0; l < d.levels; ++l) {
            const int hl = d.h[l], wl = d.w[l];
            const float inv = 1.0f / (float)(1 << l);
            // reference: x = cx / 2^l + dx ; g = 2 x / (wl - 1) - 1 ; ix = (g + 1) * (wl - 1) / 2  (fp32 round trip)
            const float xc = __fmul_rn(cx, inv), yc = __fmul_rn(cy, inv);
            const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, xc), (float)(wl - 1)), 1.0f);
            const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, yc), (float)(hl - 1)), 1.0f);
            const float ix = __fmul_rn(__fadd_rn(gx, 1.0f), (float)(wl - 1) * 0.5f);
            const float iy = __fmul_rn(__fadd_rn(gy, 1.0f), (float)(hl - 1) * 0.5f);
            const float x0f = floorf(ix), y0f = floorf(iy);
            const float fx = ix - x0f, fy = iy - y0f;
            // clamp the window origin so that int conversion is safe; fully outside windows read zeros
            const float lim = 1.0e6f;
            const int x0 = (int)fminf(fmaxf(x0f, -lim), lim) - radius;
            const int y0 = (int)fminf(fmaxf(y0f, -lim), lim) - radius;
            const float *slab = pyr + d.off[l] + ((size_t)b * Q + q) * (size_t)(hl * wl);
            for (int e = lane; e < win * win; e += 32) {
                const int ry = e / win, rx = e - ry * win;
                const int yy = y0 + ry, xx = x0 + rx;
                float v = 0.0f;
                if (yy >= 0 && yy < hl && xx >= 0 && xx < wl) v = __ldg(slab + yy * wl + xx);
                patch[e] = v;
            }
            __syncwarp();
            const float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx, w10 = fy * (1.0f - fx), w11 = fy * fx;
            for (int e = lane; e < n * n; e += 32) {
                const int i = e / n, j = e - i * n;          // i: x-offset index (slow), j: y-offset index (fast)
                const float *p = patch + j * win + i;
                const float v = p[0] * w00 + p[1] * w01 + p[win] * w10 + p[win + 1] * w11;
                s_out[(l * n * n + e) * (kLookupQ + 1) + qi] = v;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    const int nq = min(kLookupQ, Q - q_base);
    if (out_hi != nullptr) {
        // NHWC fp16 hi/lo planes for the tcgen05 convolution path: 324 contiguous channels per query
        for (int e = threadIdx.x; e < nch * kLookupQ; e += blockDim.x) {
            const int qi = e / nch, c = e - qi * nch;
            if (qi < nq) {
                const float v = s_out[c * (kLookupQ + 1) + qi];
                const plane_t h = to_plane(v);
                const size_t o = ((size_t)b * Q + q_base + qi) * nhwc_ld + c;
                out_hi[o] = h;
                out_lo[o] = to_plane_lo(v - plane_to_float(h));
            }
        }
        return;
    }
    for (int e = threadIdx.x; e < nch * kLookupQ; e += blockDim.x) {
        const int c = e / kLookupQ, qi = e - c * kLookupQ;
        if (qi < nq) out[((size_t)b * nch + c) * Q + q_base + qi] = s_out[c * (kLookupQ + 1) + qi];
    }
}

// ------------------------------------------------------------------------------------------------
// Window lookup for the tensor-core update operator: output NHWC fp16 hi/lo planes, (L*(2r+1)^2) contiguous channels per
// query -- no transposition is needed, so one WARP owns one query: all L*(2r+2)^2 neighbourhood loads of the query are issued
// back to back (up to 13 per lane in flight: a single memory round trip per query instead of one per level), staged in a
// private shared-memory patch, then the lanes evaluate two adjacent channels each and store packed fp16 pairs.
// Coordinate arithmetic and interpolation order are those of corr_lookup_kernel (bit-identical values before the split).
// ------------------------------------------------------------------------------------------------
constexpr int kLkWarps = 8;
constexpr int kLkMaxElems = 4 * kMaxWin * kMaxWin;     // 400 patch values per query
constexpr int kLkRounds = (kLkMaxElems + 31) / 32;     // 13

template <int kRadius>      // compile-time radius: the index arithmetic below is all divisions by window sizes
__global__ void __launch_bounds__(32 * kLkWarps) corr_lookup_nhwc_kernel(const float *__restrict__ pyr, PyrDims d,
                                                                         const float *__restrict__ coords,
                                                                         plane_t *__restrict__ out_hi,
                                                                         plane_t *__restrict__ out_lo, int ld, int Q,
                                                                         long long total_q, int radius_rt) {
    __shared__ float s_patch[kLkWarps][kLkMaxElems];
    __shared__ float s_frac[kLkWarps][8];               // fx, fy per level
    __shared__ int s_org[kLkWarps][8];                  // x0, y0 per level
    __shared__ size_t s_off[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int radius = kRadius > 0 ? kRadius : radius_rt;
    const int n = 2 * radius + 1, win = n + 1, ww = win * win, nn = n * n;
    const int n_elems = d.levels * ww, nch = d.levels * nn;
    const int h0 = d.h[0], w0 = d.w[0];                 // level l is (h0 >> l) x (w0 >> l)
    if (threadIdx.x < 4) s_off[threadIdx.x] = d.off[threadIdx.x];
    __syncthreads();
    float *patch = s_patch[warp];
    for (long long gq = (long long)blockIdx.x * kLkWarps + warp; gq < total_q; gq += (long long)gridDim.x * kLkWarps) {
        const int b = (int)(gq / Q), q = (int)(gq - (long long)b * Q);
        if (lane < d.levels) {
            const int l = lane;
            const float cx = __ldg(coords + ((size_t)b * 2 + 0) * Q + q);
            const float cy = __ldg(coords + ((size_t)b * 2 + 1) * Q + q);
            const int hl = h0 >> l, wl = w0 >> l;
            const float inv = 1.0f / (float)(1 << l);
            const float xc = __fmul_rn(cx, inv), yc = __fmul_rn(cy, inv);
            const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, xc), (float)(wl - 1)), 1.0f);
            const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, yc), (float)(hl - 1)), 1.0f);
            const float ix = __fmul_rn(__fadd_rn(gx, 1.0f), (float)(wl - 1) * 0.5f);
            const float iy = __fmul_rn(__fadd_rn(gy, 1.0f), (float)(hl - 1) * 0.5f);
            const float x0f = floorf(ix), y0f = floorf(iy);
            const float lim = 1.0e6f;
            s_frac[warp][2 * l] = ix - x0f;
            s_frac[warp][2 * l + 1] = iy - y0f;
            s_org[warp][2 * l] = (int)fminf(fmaxf(x0f, -lim), lim) - radius;
            s_org[warp][2 * l + 1] = (int)fminf(fmaxf(y0f, -lim), lim) - radius;
        }
        __syncwarp();
        float v[kLkRounds];
#pragma unroll
        for (int k = 0; k < kLkRounds; ++k) {
            const int e = lane + 32 * k;
            v[k] = 0.0f;
            if (e < n_elems) {
                const int l = e / ww, r = e - l * ww;
                const int ry = r / win, rx = r - ry * win;
                const int hl = h0 >> l, wl = w0 >> l;
                const int yy = s_org[warp][2 * l + 1] + ry, xx = s_org[warp][2 * l] + rx;
                if (yy >= 0 && yy < hl && xx >= 0 && xx < wl)
                    v[k] = __ldg(pyr + s_off[l] + (size_t)gq * (size_t)(hl * wl) + yy * wl + xx);
            }
        }
#pragma unroll
        for (int k = 0; k < kLkRounds; ++k) {
            const int e = lane + 32 * k;
            if (e < n_elems) patch[e] = v[k];
        }
        __syncwarp();
        const size_t obase = (size_t)gq * ld;
        for (int c = 2 * lane; c < nch; c += 64) {
            float o[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int cc = c + t;
                o[t] = 0.0f;
                if (cc < nch) {
                    const int l = cc / nn, e = cc - l * nn;
                    const int i = e / n, j = e - i * n;          // i: x-offset index (slow), j: y-offset index (fast)
                    const float fx = s_frac[warp][2 * l], fy = s_frac[warp][2 * l + 1];
                    const float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx, w10 = fy * (1.0f - fx), w11 = fy * fx;
                    const float *pp = patch + l * ww + j * win + i;
                    o[t] = pp[0] * w00 + pp[1] * w01 + pp[win] * w10 + pp[win + 1] * w11;
                }
            }
            const plane_t h0 = to_plane(o[0]), h1 = to_plane(o[1]);
            const plane2_t hh = __halves2half2(h0, h1);
            const plane2_t ll = __halves2half2(to_plane_lo(o[0] - plane_to_float(h0)),
                                                         to_plane_lo(o[1] - plane_to_float(h1)));
            if (c + 1 < nch) {
                *reinterpret_cast<plane2_t *>(out_hi + obase + c) = hh;
                *reinterpret_cast<plane2_t *>(out_lo + obase + c) = ll;
            } else {
                out_hi[obase + c] = h0;
                out_lo[obase + c] = __low2half(ll);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Radius-4, 4-level specialisation of the NHWC lookup (the shape RAFT uses): same arithmetic, a third of the instructions.
// The first version spent more issue slots on index arithmetic (divisions by the window size, level selection per element)
// than on memory: ~1000 warp instructions per query against ~50 cache lines.  Here every loop is level-uniform and unrolled:
//   load     level l, round j: lane = r * 10 + c (r < 3, c < 10) fetches window row 3j + r, column c  -- 4 rounds per level, all
//            16 loads of a lane in flight before the first use; (r, c) are per-lane constants;
//   compute  level l, pass m: lane evaluates output e = lane + 32 m (< 81) from the staged 10x10 patch; its patch offset
//            (e % 9) * 10 + e / 9 is a per-lane constant as well (channel e = x-offset * 9 + y-offset).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kLkWarps) corr_lookup_nhwc_r4_kernel(const float *__restrict__ pyr, PyrDims d,
                                                                            const float *__restrict__ coords, plane_t *__restrict__ out_hi,
                                                                            plane_t *__restrict__ out_lo, int ld, int Q, long long total_q) {
    __shared__ float s_patch[kLkWarps][4][104];          // [level][10 x 10] (+4 floats: keeps the levels 16-byte apart)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane / 10, c = lane - 10 * r;          // window row (within a round of 3 rows) and column fetched by this lane
    const bool loader = lane < 30;
    int poff[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const int e = lane + 32 * m;                     // e = i * 9 + j: i = x offset (slow), j = y offset (fast)
        const int i = e / 9, j = e - 9 * i;
        poff[m] = j * 10 + i;
    }
    const int h0 = d.h[0], w0 = d.w[0];
    float(*patch)[104] = s_patch[warp];
    for (long long gq = (long long)blockIdx.x * kLkWarps + warp; gq < total_q; gq += (long long)gridDim.x * kLkWarps) {
        const int b = (int)(gq / Q), q = (int)(gq - (long long)b * Q);
        const float cx = __ldg(coords + ((size_t)b * 2 + 0) * Q + q);
        const float cy = __ldg(coords + ((size_t)b * 2 + 1) * Q + q);
        float fx[4], fy[4], v[4][4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int hl = h0 >> l, wl = w0 >> l;
            const float inv = 1.0f / (float)(1 << l);
            // reference: x = cx / 2^l + dx ; g = 2 x / (wl - 1) - 1 ; ix = (g + 1) * (wl - 1) / 2  (fp32 round trip, corr_lookup_kernel)
            const float xc = __fmul_rn(cx, inv), yc = __fmul_rn(cy, inv);
            const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, xc), (float)(wl - 1)), 1.0f);
            const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, yc), (float)(hl - 1)), 1.0f);
            const float ix = __fmul_rn(__fadd_rn(gx, 1.0f), (float)(wl - 1) * 0.5f);
            const float iy = __fmul_rn(__fadd_rn(gy, 1.0f), (float)(hl - 1) * 0.5f);
            const float x0f = floorf(ix), y0f = floorf(iy);
            fx[l] = ix - x0f, fy[l] = iy - y0f;
            const float lim = 1.0e6f;
            const int xx = (int)fminf(fmaxf(x0f, -lim), lim) - 4 + c;
            const int y0 = (int)fminf(fmaxf(y0f, -lim), lim) - 4 + r;
            const float *slab = pyr + d.off[l] + (size_t)gq * (size_t)(hl * wl);
            const bool xin = loader && xx >= 0 && xx < wl;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = y0 + 3 * j;
                v[l][j] = 0.0f;
                if (xin && yy >= 0 && yy < hl && (j < 3 || r == 0)) v[l][j] = __ldg(slab + yy * wl + xx);
            }
        }
        if (loader) {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
#pragma unroll
                for (int j = 0; j < 3; ++j) patch[l][(3 * j + r) * 10 + c] = v[l][j];
                if (r == 0) patch[l][90 + c] = v[l][3];
            }
        }
        __syncwarp();
        const size_t obase = (size_t)gq * ld;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const float w00 = (1.0f - fy[l]) * (1.0f - fx[l]), w01 = (1.0f - fy[l]) * fx[l], w10 = fy[l] * (1.0f - fx[l]), w11 = fy[l] * fx[l];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                if (m < 2 || lane < 81 - 64) {
                    const float *pp = patch[l] + poff[m];
                    const float o = pp[0] * w00 + pp[1] * w01 + pp[10] * w10 + pp[11] * w11;
                    const plane_t hh = to_plane(o);
                    const size_t oo = obase + l * 81 + lane + 32 * m;
                    out_hi[oo] = hh;
                    out_lo[oo] = to_plane_lo(o - plane_to_float(hh));
                }
            }
        }
        __syncwarp();
    }
}

// ================================================================================================
// Host side
// ================================================================================================
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int load_encode() {
    if (g_encode) return RPE_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return cuda_fail(e == cudaSuccess ? cudaErrorUnknown : e);
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return RPE_OK;
}

// 3-D fp32 tensor (inner, rows, batch) with a (box_inner x box_rows x 1) box, SWIZZLE_128B.
static int make_map(CUtensorMap *m, const void *base, uint64_t inner, uint64_t rows, uint64_t batch, uint32_t box_inner,
                    uint32_t box_rows) {
    const uint64_t es = 4;
    cuuint64_t dims[3] = {inner, rows, batch};
    cuuint64_t strides[2] = {inner * es, inner * rows * es};
    cuuint32_t box[3] = {box_inner, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = 100000 + (int)r;
        return RPE_ERR_CUDA;
    }
    return RPE_OK;
}

static PyrDims pyr_dims(int B, int h, int w, int levels) {
    PyrDims d;
    size_t off = 0;
    d.levels = levels;
    for (int l = 0; l < 4; ++l) {
        d.h[l] = l < levels ? (h >> l) : 0;
        d.w[l] = l < levels ? (w >> l) : 0;
        d.off[l] = off;
        if (l < levels) {
            size_t n = (size_t)B * h * w * d.h[l] * d.w[l];
            off += (n + 63) & ~(size_t)63;      // keep every level 256-byte aligned
        }
    }
    return d;
}

}  // namespace rpe

extern "C" {

size_t rpe_corr_level_offset(int B, int h, int w, int level) {
    if (level < 0 || level > 3) return 0;
    return rpe::pyr_dims(B, h, w, 4).off[level] * sizeof(float);
}

size_t rpe_corr_pyramid_bytes(int B, int h, int w, int num_levels) {
    if (num_levels < 1 || num_levels > 4) return 0;
    rpe::PyrDims d = rpe::pyr_dims(B, h, w, num_levels);
    const int l = num_levels - 1;
    size_t n = (size_t)B * h * w * d.h[l] * d.w[l];
    return (d.off[l] + ((n + 63) & ~(size_t)63)) * sizeof(float);
}

size_t rpe_corr_workspace_bytes(int B, int C, int h, int w, int precision) {
    if (precision == RPE_CORR_F16X3)      // four NHWC fp16 planes (hi / lo of both feature maps)
        return 4 * (((size_t)B * h * w * C * 2 + 1023) & ~(size_t)1023);
    const size_t Kp = precision == RPE_CORR_TF32 ? (size_t)C : 3 * (size_t)C;
    return 2 * (((size_t)B * h * w * Kp * sizeof(float) + 1023) & ~(size_t)1023);
}

int rpe_corr_build(const float *fmap1, const float *fmap2, float *pyramid, int B, int C, int h, int w, int num_levels,
                   int precision, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace rpe;
    if (!fmap1 || !fmap2 || !pyramid || !workspace) return RPE_ERR_INVALID_ARG;
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || num_levels < 1 || num_levels > 4) return RPE_ERR_INVALID_ARG;
    if (C % 32 != 0) return RPE_ERR_INVALID_ARG;
    if (precision != RPE_CORR_TF32 && precision != RPE_CORR_TF32X3 && precision != RPE_CORR_F16X3) return RPE_ERR_INVALID_ARG;
    const bool f16 = precision == RPE_CORR_F16X3;
    if (f16 && C % 64 != 0) return RPE_ERR_INVALID_ARG;
    if ((h >> (num_levels - 1)) < 1 || (w >> (num_levels - 1)) < 1) return RPE_ERR_INVALID_ARG;
    const int Q = h * w;
    if (Q % 4 != 0) return RPE_ERR_INVALID_ARG;     // TMA global strides must be multiples of 16 bytes
    if (workspace_bytes < rpe_corr_workspace_bytes(B, C, h, w, precision)) return RPE_ERR_WORKSPACE;
    if (!aligned16(pyramid) || (reinterpret_cast<uintptr_t>(workspace) & 1023u)) return RPE_ERR_ALIGNMENT;
    int rc = load_encode();
    if (rc != RPE_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (f16) {
        // fp16x3: NHWC hi/lo planes of both feature maps, then the fused volume + pyramid kernel (conv.cu kind 7): level 0 and
        // the three pooled levels leave the tensor-memory accumulators in one pass, no second read of the volume
        const size_t plane = (((size_t)B * Q * C * 2) + 1023) & ~(size_t)1023;
        char *ws = reinterpret_cast<char *>(workspace);
        void *f1h = ws, *f1l = ws + plane, *f2h = ws + 2 * plane, *f2l = ws + 3 * plane;
        if ((rc = rpe_nchw_to_nhwc_split(fmap1, f1h, f1l, nullptr, B, C, h, w, C, 0, 0, 0, stream)) != RPE_OK) return rc;
        if ((rc = rpe_nchw_to_nhwc_split(fmap2, f2h, f2l, nullptr, B, C, h, w, C, 0, 0, 0, stream)) != RPE_OK) return rc;
        return rpe_corr_build_planes(f1h, f1l, f2h, f2l, pyramid, B, C, h, w, num_levels, 0, 0, stream);
    }
    const int split = precision != RPE_CORR_TF32;
    const int Kp = split ? 3 * C : C;
    float *opA = reinterpret_cast<float *>(workspace);
    float *opB = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + rpe_corr_workspace_bytes(B, C, h, w, precision) / 2);

    dim3 pgrid((Q + 31) / 32, C / 32, B);
    corr_prep_kernel<<<pgrid, 256, 0, st>>>(fmap1, opA, C, Q, split, 0);
    RPE_LAUNCH_CHECK();
    corr_prep_kernel<<<pgrid, 256, 0, st>>>(fmap2, opB, C, Q, split, 1);
    RPE_LAUNCH_CHECK();

    PyrDims d = pyr_dims(B, h, w, num_levels);
    CUtensorMap ma, mb, mc;
    if ((rc = make_map(&ma, opA, Kp, Q, B, kBK, kBM)) != RPE_OK) return rc;
    if ((rc = make_map(&mb, opB, Kp, Q, B, kBK, kBN)) != RPE_OK) return rc;
    if ((rc = make_map(&mc, pyramid + d.off[0], Q, Q, B, kEpiCols, kBM)) != RPE_OK) return rc;

    GemmShape s;
    s.B = B, s.Q = Q, s.Kp = Kp;
    s.m_tiles = (Q + kBM - 1) / kBM;
    s.n_tiles = (Q + kBN - 1) / kBN;
    s.scale = 1.0f / sqrtf((float)C);
    // (function attributes are per device: set on every call -- a host-side no-op after the first)
    RPE_CUDA_TRY(cudaFuncSetAttribute(corr_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    int grid = sm_count();
    const int tiles = s.B * s.m_tiles * s.n_tiles;
    if (grid > tiles) grid = tiles;
    corr_gemm_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(ma, mb, mc, s);
    RPE_LAUNCH_CHECK();

    if (num_levels > 1) {
        size_t smem = 0;
        for (int l = 0; l < num_levels; ++l) smem += (size_t)d.h[l] * d.w[l] * sizeof(float);
        if (smem > 200 * 1024) return RPE_ERR_INVALID_ARG;
        if (smem > 48 * 1024) RPE_CUDA_TRY(cudaFuncSetAttribute(corr_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        corr_pool_kernel<<<B * Q, 256, smem, st>>>(pyramid, d);
        RPE_LAUNCH_CHECK();
    }
    return RPE_OK;
}

static int corr_lookup_impl(const float *pyramid, const float *coords, float *out, void *out_hi, void *out_lo, int nhwc_ld, int B,
                            int h, int w, int num_levels, int radius, void *stream) {
    using namespace rpe;
    if (!pyramid || !coords || (!out && !out_hi)) return RPE_ERR_INVALID_ARG;
    if (B <= 0 || h <= 0 || w <= 0 || num_levels < 1 || num_levels > 4 || radius < 1 || 2 * radius + 2 > kMaxWin)
        return RPE_ERR_INVALID_ARG;
    const int Q = h * w;
    PyrDims d = pyr_dims(B, h, w, num_levels);
    const int n = 2 * radius + 1;
    const int nch = num_levels * n * n;
    if (out_hi && (nhwc_ld % 2) == 0 && (reinterpret_cast<uintptr_t>(out_hi) & 3u) == 0 && (reinterpret_cast<uintptr_t>(out_lo) & 3u) == 0) {
        const long long total_q = (long long)B * Q;
        long long blocks = (total_q + kLkWarps - 1) / kLkWarps;
        const long long cap = (long long)sm_count() * 8 * 4;
        if (blocks > cap) blocks = cap;
        if (radius == 4 && num_levels == 4)
            corr_lookup_nhwc_r4_kernel<<<(unsigned)blocks, 32 * kLkWarps, 0, (cudaStream_t)stream>>>(
                pyramid, d, coords, reinterpret_cast<rpe::plane_t *>(out_hi), reinterpret_cast<rpe::plane_t *>(out_lo), nhwc_ld, Q, total_q);
        else if (radius == 4)
            corr_lookup_nhwc_kernel<4><<<(unsigned)blocks, 32 * kLkWarps, 0, (cudaStream_t)stream>>>(
                pyramid, d, coords, reinterpret_cast<rpe::plane_t *>(out_hi), reinterpret_cast<rpe::plane_t *>(out_lo), nhwc_ld, Q,
                total_q, radius);
        else
            corr_lookup_nhwc_kernel<0><<<(unsigned)blocks, 32 * kLkWarps, 0, (cudaStream_t)stream>>>(
                pyramid, d, coords, reinterpret_cast<rpe::plane_t *>(out_hi), reinterpret_cast<rpe::plane_t *>(out_lo), nhwc_ld, Q,
                total_q, radius);
        RPE_LAUNCH_CHECK();
        return RPE_OK;
    }
    const size_t smem = ((size_t)nch * (kLookupQ + 1) + 8 * kMaxWin * kMaxWin) * sizeof(float);
    if (smem > 48 * 1024) RPE_CUDA_TRY(cudaFuncSetAttribute(corr_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((Q + kLookupQ - 1) / kLookupQ, B);
    corr_lookup_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(pyramid, d, coords, out, Q, radius,
                                                                  reinterpret_cast<rpe::plane_t *>(out_hi),
                                                                  reinterpret_cast<rpe::plane_t *>(out_lo), nhwc_ld);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_corr_lookup(const float *pyramid, const float *coords, float *out, int B, int h, int w, int num_levels, int radius,
                    void *stream) {
    return corr_lookup_impl(pyramid, coords, out, nullptr, nullptr, 0, B, h, w, num_levels, radius, stream);
}

int rpe_corr_lookup_nhwc_split(const float *pyramid, const float *coords, void *out_hi, void *out_lo, int ld, int B, int h, int w,
                              int num_levels, int radius, void *stream) {
    const int n = 2 * radius + 1;
    if (!out_hi || !out_lo || ld < num_levels * n * n) return RPE_ERR_INVALID_ARG;
    return corr_lookup_impl(pyramid, coords, nullptr, out_hi, out_lo, ld, B, h, w, num_levels, radius, stream);
}

}  // extern "C"
