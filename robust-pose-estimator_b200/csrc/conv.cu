// Implicit-GEMM 2-D convolution on tcgen05 tensor cores: the convolutional trunk of RAFT (SURVEY.md section 8f-1;
// reference: /root/reference/core/RAFT/core/update.py:79-136 update operator, core/RAFT/core/extractor.py:118-192 encoders,
// which the reference runs through cuDNN).
//
//   v[n, y, x, co]   = bias[co] + pre[n, y, x, co] + sum_s sum_(ky,kx) sum_ci  A_s[n, y*st+ky-ph, x*st+kx-pw, ci] * W_s[ky*kw+kx][co][ci]
//   out[n, y, x, co] = act(v) * scale                       (then  out = relu(out + res[n, y, x, co])  when a residual is given)
//
// Activations are NHWC fp16 planes.  The contraction runs over a LIST of sources (activation tensor + its weight slice), so
// concatenated inputs are never materialised, and every source may carry two planes hi = fp16(v), lo = fp16(v - hi): the
// "fp16x3" arithmetic evaluates hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (22 mantissa bits per operand; measured
// deviation from fp32 convolutions ~1e-5 relative).  Stride 1 or 2, "same"-style zero padding (k/2).
//
// Kernel: persistent, warp-specialised.  The output tile is 128 pixels = 16 (major) x 8 (minor) positions of one image; `orient`
// picks which image axis is the minor one.  Shared memory holds an activation SLAB of (16 + kmaj - 1) x 8 positions x 64
// channels per plane, laid out [major][minor][64 ch] with the 128-byte swizzle, so the UMMA descriptor of a filter tap that is
// shifted along the major axis is the same slab advanced by 1024 bytes: one TMA load serves all kmaj taps of a filter column
// (3x3: 3 loads instead of 9; 1x5 / 5x1: 1 load instead of 5).  Out-of-image coordinates are zero-filled by the TMA unit, which
// IS the zero padding.  The hi and lo planes of a stage are loaded once and feed all three products.  Weights stream through
// their own ring, one [bn][64] tile per tap (or stay resident in shared memory when the whole set fits).  Roles: warp 0 =
// activation TMA producer, warp 3 = weight TMA producer, warp 1 = tcgen05.mma issuer (the whole warp walks the loop nest so that
// descriptors live in uniform registers; MMAs / commits are predicated on one elected lane), warp 2 = TMEM allocator, warps
// 4-11 = epilogue (tcgen05.ld -> smem transposition -> bias / addend / activation / residual / GRU arithmetic -> coalesced fp32
// and/or fp16 hi/lo NHWC stores) overlapped with the next tile through two TMEM accumulator stages.  CTA pairs (cta_group::2)
// share every weight tile.  The kernel is instantiated once per epilogue KIND (kK* below) so that each instance carries only
// the code of the tensors it touches: the epilogue warps share issue slots with the MMA issuer and the instruction cache with
// both producers (profiles/README.md: the issuer, not the tensor pipe, was the bottleneck of the first version).
#include <cuda_fp16.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"

namespace rpe {

// RPE_CONV_DEBUG (probe switches: 1 no loads, 2 no MMAs, 4 no epilogue memory traffic, 8 no fp32 store of the GRU state) exists only in builds made with
// -DRPE_CONV_DEBUG_BUILD; in the production library every test of it folds to a constant.
#ifdef RPE_CONV_DEBUG_BUILD
#define RPE_CV_DBG(P) ((P).dbg)
#else
#define RPE_CV_DBG(P) 0
#endif

constexpr int kCvBK = 64;                             // fp16 channels per K block (one 128-byte swizzle row)
constexpr int kCvAcc = 2;                             // TMEM accumulator stages
constexpr int kCvEpiWarps = 8;                        // 2 per SM sub-partition: column halves of the same TMEM lanes
constexpr int kCvThreads = 128 + 32 * kCvEpiWarps;
constexpr int kCvMaxSrc = 4;
constexpr int kCvMaxBN = 256;
constexpr int kCvMaxAStages = 4, kCvMaxBStages = 16;
constexpr int kCvSmemData = 205 * 1024;               // budget for the two rings
constexpr int kCvMaxCout = 1024;                      // bias staged in shared memory
constexpr uint32_t kCvGroupMaxTile = 8 * 1024;          // weight plane tiles up to this size are grouped per ring entry (see the plan)
constexpr int kCvStageBytes = kCvEpiWarps * 32 * 16 * 4;   // epilogue transpose tiles: one [32 rows][16 floats] per warp
constexpr int kCvSmem = kCvSmemData + 1024 + 512 + kCvMaxCout * 4 + kCvStageBytes;   // + alignment slack + barriers + bias + staging

struct alignas(64) ConvParams {
    CUtensorMap amap[kCvMaxSrc][2];
    CUtensorMap wmap[kCvMaxSrc][2];
    CUtensorMap wmapx[kCvMaxSrc][2];                  // merged-N mode: the same weight planes with a box of bn (not bn / 2) rows
    int cblocks[kCvMaxSrc];
    int ksteps_last[kCvMaxSrc];                       // 16-channel MMA steps of the last K block (1..4)
    int n_src, n_planes, w_planes;                    // planes of the activations (1 or 2) and of the weights (1 or 2)
    int N, OH, OW, tiles_min, tiles_maj;
    int orient;                                       // 0: minor axis = x, 1: minor axis = y
    int kmin, kmaj, pad_min, pad_maj, stride, kw;
    int reuse;                                        // 1: one slab per filter column serves all kmaj taps (stride 1)
    uint32_t a_plane_bytes, a_stage_bytes, b_plane_bytes, b_stage_bytes;
    int n_a_stages, n_b_stages;
    int side_prefetch;                                // GRU kinds: the idle warp prefetches the next unit's side inputs into L2
    int merge_n;                                      // 1: hi and lo weight planes side by side as ONE N = 2 bn operand (see conv_body)
    int b_group;                                      // plane tiles per weight-ring entry: 1, w_planes (one tap) or kmaj * w_planes (one filter column)
    int resident_b;                                   // 1: the whole weight set stays in shared memory (loaded once per CTA)
    int cb_base[kCvMaxSrc];                           // first K block of each source in the resident weight array
    int cout, bn, n_blocks;
    const float *bias;
    const float *pre;
    int pre_ld;
    const float *res;
    const plane_t *res_hi, *res_lo;                   // the residual as split planes (channel pitch res_ld) instead of fp32
    int res_ld;
    int act;
    float scale;
    float acc_scale;                                  // accumulators are multiplied by this first (weights packed pre-scaled by its inverse)
    float *out_f32;
    int f32_ld, f32_off;
    plane_t *out_hi, *out_lo;
    int bf_ld, bf_off;
    int mode;                                         // 0 plain, 1 GRU z|r gates, 2 GRU candidate + state update
    float *aux;                                       // mode 1/2: hidden state h (fp32 NHWC, updated in place by mode 2)
    int aux_ld;
    const float *aux2;                                // mode 2: update gate z (fp32 NHWC)
    int aux2_ld;
    float *stat_part;                                 // kind 6: per (pixel tile, lane quarter) sums of out and out^2, [slot][cout_pad][2]
    int dbg;                                          // RPE_CONV_DEBUG bits (probe only): 1 no loads, 2 no MMAs, 4 no epilogue memory traffic
    // kind 7 (all-pairs correlation volume + pooled pyramid, corr.py:12-27,52-60): the "weights" are the second feature map of
    // the SAME image, read through a 4-D map as 16x16 target blocks; level l of the pyramid is (N*OH*OW, corr_h >> l, corr_w >> l)
    float *corr_lvl[4];
    int corr_h, corr_w, corr_nbx, corr_levels;
    int corr_vec;                                     // bit l: level l rows may be written with vector stores
    int corr_a_wrap, corr_a_sub;                      // query-side image of sample s: s < wrap ? s : s - sub (target side: image s)
    int resident_a;                                   // kind 7: the query tile (all K blocks) stays in shared memory while the CTA walks
                                                      // every target block of its image; only the target tiles stream
};

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
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

__device__ __forceinline__ float cv_activate_rt(float v, int act) {
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) return 1.0f / (1.0f + expf(-v));
    if (act == 3) return tanhf(v);
    return v;
}

// sigmoid / tanh on the hardware exponential and reciprocal (ex2.approx, rcp.approx): saturate correctly at +-inf
__device__ __forceinline__ float cv_sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float cv_tanh_fast(float v) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v)); }

// Kernel kinds (template parameter kK): the epilogue of each is compiled for exactly the tensors it touches, because the epilogue
// warps share their issue slots with the MMA issuer and the instruction cache with both producers.
//   0 generic mode 0 (every optional tensor tested at run time)      1 GRU z|r gates (mode 1)      2 GRU candidate + state (mode 2)
//   3 tap projection (mode 3)      4 mode 0, split planes only (no addend / residual / fp32 copy)      5 mode 0, fp32 only
//   6 = 5 + instance-norm partial sums: every epilogue warp reduces its 32 pixels and writes sum / sum of squares per channel
//   7 all-pairs correlation: N tile = one 16x16 block of TARGET positions of the same image; the epilogue scales by 1/sqrt(C), writes
//     level 0 and derives the 2x2 / 4x4 / 8x8 mean-pooled levels of the pyramid from the accumulators in registers
constexpr int kKGeneric = 0, kKGates = 1, kKState = 2, kKProj = 3, kKPlanes = 4, kKF32 = 5, kKF32Stats = 6, kKCorr = 7;

// fp32 x4 -> fp16 hi / lo planes (hi = fp16(v), lo = fp16(v - hi)), packed saturating conversions
template <bool kLo>
__device__ __forceinline__ void cv_store_split4(const ConvParams &P, uint32_t o, const float *v) {
    const plane2_t h01 = to_plane2(v[0], v[1]), h23 = to_plane2(v[2], v[3]);
    uint2 hv;
    hv.x = *reinterpret_cast<const uint32_t *>(&h01);
    hv.y = *reinterpret_cast<const uint32_t *>(&h23);
    *reinterpret_cast<uint2 *>(P.out_hi + o) = hv;
    if (kLo) {
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const plane2_t l01 = to_plane2_lo(v[0] - f01.x, v[1] - f01.y), l23 = to_plane2_lo(v[2] - f23.x, v[3] - f23.y);
        uint2 lv;
        lv.x = *reinterpret_cast<const uint32_t *>(&l01);
        lv.y = *reinterpret_cast<const uint32_t *>(&l23);
        *reinterpret_cast<uint2 *>(P.out_lo + o) = lv;
    }
}

// Ragged tail (1..3 channels) of an output-channel count that is not a multiple of 4 (mode 0 only).
__device__ __noinline__ void cv_epilogue_tail(const ConvParams &P, int act, float a0, float a1, float a2, const float *sbias, int co,
                                              uint32_t pix) {
    const float acc[3] = {a0, a1, a2};
    for (int k = 0; k < 3 && co + k < P.cout; ++k) {
        float r = fmaf(acc[k], P.acc_scale, sbias[co + k]);
        if (P.pre) r += __ldg(P.pre + (size_t)pix * P.pre_ld + co + k);
        r = cv_activate_rt(r, act) * P.scale;
        if (P.res) r = fmaxf(r + __ldg(P.res + (size_t)pix * P.res_ld + co + k), 0.0f);
        if (P.res_hi)
            r = fmaxf(r + (plane_to_float(P.res_hi[(size_t)pix * P.res_ld + co + k]) + plane_to_float(P.res_lo[(size_t)pix * P.res_ld + co + k])), 0.0f);
        if (P.out_f32) P.out_f32[(size_t)pix * P.f32_ld + P.f32_off + co + k] = r;
        if (P.out_hi) {
            const plane_t h = to_plane(r);
            P.out_hi[(size_t)pix * P.bf_ld + P.bf_off + co + k] = h;
            if (P.out_lo) P.out_lo[(size_t)pix * P.bf_ld + P.bf_off + co + k] = to_plane_lo(r - plane_to_float(h));
        }
    }
}

// Side inputs of one group of 4 channels (addend, residual / hidden state, update gate).  They do not depend on the
// accumulators, so they are requested one 16-column chunk ahead.  Every element is read at most once before any write to it.
// Element offsets are 32-bit (the plan rejects tensors of 2^32 elements or more).
struct CvSide {
    float4 pre, a, z;
};

template <int kK>
__device__ __forceinline__ void cv_side_load(const ConvParams &P, CvSide &sd, int co, uint32_t pix) {
    if (kK == kKPlanes || kK == kKF32 || kK == kKF32Stats || kK == kKProj || kK == kKCorr) return;
    if (co + 3 >= P.cout) return;
    if (kK == kKGeneric) {
        if (P.pre) sd.pre = __ldg(reinterpret_cast<const float4 *>(P.pre + (pix * (uint32_t)P.pre_ld + co)));
        if (P.res) sd.a = __ldg(reinterpret_cast<const float4 *>(P.res + (pix * (uint32_t)P.res_ld + co)));
        if (P.res_hi) {      // (plain loads: the residual planes may be the output planes of this very convolution)
            const uint2 h = *reinterpret_cast<const uint2 *>(P.res_hi + (pix * (uint32_t)P.res_ld + co));
            const uint2 l = *reinterpret_cast<const uint2 *>(P.res_lo + (pix * (uint32_t)P.res_ld + co));
            const float2 h01 = __half22float2(*reinterpret_cast<const plane2_t *>(&h.x)), h23 = __half22float2(*reinterpret_cast<const plane2_t *>(&h.y));
            const float2 l01 = __half22float2(*reinterpret_cast<const plane2_t *>(&l.x)), l23 = __half22float2(*reinterpret_cast<const plane2_t *>(&l.y));
            sd.a = make_float4(h01.x + l01.x, h01.y + l01.y, h23.x + l23.x, h23.y + l23.y);
        }
    } else if (kK == kKGates) {
        sd.pre = __ldg(reinterpret_cast<const float4 *>(P.pre + (pix * (uint32_t)P.pre_ld + co)));
        const int half = P.cout >> 1;
        if (co >= half) sd.a = __ldg(reinterpret_cast<const float4 *>(P.aux + (pix * (uint32_t)P.aux_ld + (co - half))));
    } else {
        sd.pre = __ldg(reinterpret_cast<const float4 *>(P.pre + (pix * (uint32_t)P.pre_ld + co)));
        sd.z = __ldg(reinterpret_cast<const float4 *>(P.aux2 + (pix * (uint32_t)P.aux2_ld + co)));
        sd.a = *reinterpret_cast<const float4 *>(P.aux + (pix * (uint32_t)P.aux_ld + co));        // (plain load: the state is updated in place by this kernel)
    }
}

// One group of 4 consecutive output channels of one pixel: bias / addend / activation / scale / residual / stores.
template <int kK>
__device__ __forceinline__ void cv_epilogue_group(const ConvParams &P, const float4 acc, const CvSide &sd, const float4 b, const float *sbias,
                                                  int co, uint32_t pix) {
    if (co >= P.cout) return;
    float o[4];
    if (co + 3 < P.cout) {
        o[0] = fmaf(acc.x, P.acc_scale, b.x), o[1] = fmaf(acc.y, P.acc_scale, b.y), o[2] = fmaf(acc.z, P.acc_scale, b.z), o[3] = fmaf(acc.w, P.acc_scale, b.w);
        if (kK == kKGates || kK == kKState || (kK == kKGeneric && P.pre)) o[0] += sd.pre.x, o[1] += sd.pre.y, o[2] += sd.pre.z, o[3] += sd.pre.w;
        // GRU kinds fix the activation (z|r: sigmoid, q: tanh) and use the hardware exponential / reciprocal (abs. error
        // < 5e-7, far below the 2^-16 of the split operands); planes / fp32 kinds only know relu / none (checked by the plan)
        if (kK == kKGates) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = cv_sigmoid_fast(o[k]);
        } else if (kK == kKState) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = cv_tanh_fast(o[k]);
        } else if (P.act == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.0f);
        } else if (kK == kKGeneric && P.act == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = 1.0f / (1.0f + expf(-o[k]));
        } else if (kK == kKGeneric && P.act == 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = tanhf(o[k]);
        }
        if ((kK == kKGeneric || kK == kKF32) && P.scale != 1.0f) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] *= P.scale;
        }
        if (kK == kKGeneric) {
            if (P.res || P.res_hi) {
                o[0] = fmaxf(o[0] + sd.a.x, 0.0f), o[1] = fmaxf(o[1] + sd.a.y, 0.0f);
                o[2] = fmaxf(o[2] + sd.a.z, 0.0f), o[3] = fmaxf(o[3] + sd.a.w, 0.0f);
            }
            if (P.out_f32) *reinterpret_cast<float4 *>(P.out_f32 + (pix * (uint32_t)P.f32_ld + P.f32_off + co)) = make_float4(o[0], o[1], o[2], o[3]);
            if (P.out_hi) {
                if (P.out_lo) cv_store_split4<true>(P, pix * (uint32_t)P.bf_ld + P.bf_off + co, o);
                else cv_store_split4<false>(P, pix * (uint32_t)P.bf_ld + P.bf_off + co, o);
            }
        } else if (kK == kKPlanes) {
            cv_store_split4<true>(P, pix * (uint32_t)P.bf_ld + P.bf_off + co, o);
        } else if (kK == kKF32) {
            *reinterpret_cast<float4 *>(P.out_f32 + (pix * (uint32_t)P.f32_ld + P.f32_off + co)) = make_float4(o[0], o[1], o[2], o[3]);
        } else if (kK == kKGates) {
            // SepConvGRU gates (update.py:45-50, 53-58): channels [0, cout/2) = z -> fp32; [cout/2, cout) = r -> planes of r * h
            const int half = P.cout >> 1;
            if (co < half) {
                *reinterpret_cast<float4 *>(P.out_f32 + (pix * (uint32_t)P.f32_ld + P.f32_off + co)) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
                o[0] *= sd.a.x, o[1] *= sd.a.y, o[2] *= sd.a.z, o[3] *= sd.a.w;
                cv_store_split4<true>(P, pix * (uint32_t)P.bf_ld + P.bf_off + (co - half), o);
            }
        } else {
            // candidate state q = tanh(.) and the state update h = (1 - z) * h + z * q, fp32 in place + planes
            o[0] = (1.0f - sd.z.x) * sd.a.x + sd.z.x * o[0];
            o[1] = (1.0f - sd.z.y) * sd.a.y + sd.z.y * o[1];
            o[2] = (1.0f - sd.z.z) * sd.a.z + sd.z.z * o[2];
            o[3] = (1.0f - sd.z.w) * sd.a.w + sd.z.w * o[3];
            if (!(RPE_CV_DBG(P) & 8))        // (probe bit 8: what would a state kept in planes only save?  measured: see DESIGN)
                *reinterpret_cast<float4 *>(P.aux + (pix * (uint32_t)P.aux_ld + co)) = make_float4(o[0], o[1], o[2], o[3]);
            cv_store_split4<true>(P, pix * (uint32_t)P.bf_ld + P.bf_off + co, o);
        }
    } else if (kK == kKGeneric || kK == kKPlanes || kK == kKF32) {      // ragged tail of a channel count that is not a multiple of 4
        cv_epilogue_tail(P, P.act, acc.x, acc.y, acc.z, sbias, co, pix);
    }
}

// Side inputs of the warp's 32 pixels for one 16-column chunk (issued a chunk ahead of their use).
template <int kK>
__device__ __forceinline__ void cv_side_load4(const ConvParams &P, CvSide *sd, int co, const uint32_t *pix, uint32_t inside_mask) {
#pragma unroll
    for (int it = 0; it < 4; ++it)
        if ((inside_mask >> it) & 1u) cv_side_load<kK>(P, sd[it], co, pix[it]);
}

// 16 accumulator columns of the warp's 32 pixels, already transposed through shared memory: lane = (pixel row it*8 + lane/4,
// channel group lane%4), i.e. 4 lanes cover 64 contiguous bytes of one pixel and a warp instruction touches 8 pixels.
// stage: [32 rows][4 chunks of 16 B], chunk k of row r stored at k ^ ((r >> 1) & 3) (conflict-free both ways).
template <int kK>
__device__ __forceinline__ void cv_epilogue_half(const ConvParams &P, const float *stage, const float *sbias, int co, const uint32_t *pix,
                                                 uint32_t inside_mask, int lane, const CvSide *sd) {
    float4 acc[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        acc[it] = *reinterpret_cast<const float4 *>(stage + r * 16 + (((lane & 3) ^ ((r >> 1) & 3)) << 2));
    }
    const float4 b = *reinterpret_cast<const float4 *>(sbias + co);          // the bias area is padded to bn * n_blocks floats
#pragma unroll
    for (int it = 0; it < 4; ++it)
        if ((inside_mask >> it) & 1u) cv_epilogue_group<kK>(P, acc[it], sd[it], b, sbias, co, pix[it]);
}

// fp32-only output + instance-norm partial sums (kind 6; cout is a multiple of 16, activation none / relu, no scale):
// the lane sums its 4 pixels per channel, the 8 lanes that hold the same channel group are combined with shuffles, and lanes
// 0..3 write (sum, sum of squares) of 4 channels each into this warp's slot -- no atomics, deterministic.
__device__ __forceinline__ void cv_epilogue_half_stats(const ConvParams &P, const float *stage, const float *sbias, int co, const uint32_t *pix,
                                                       uint32_t inside_mask, int lane, float *slot) {
    const float4 b = *reinterpret_cast<const float4 *>(sbias + co);
    float s[4] = {0.0f, 0.0f, 0.0f, 0.0f}, q[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        const float4 acc = *reinterpret_cast<const float4 *>(stage + r * 16 + (((lane & 3) ^ ((r >> 1) & 3)) << 2));
        if (((inside_mask >> it) & 1u) && co < P.cout) {
            float o[4] = {fmaf(acc.x, P.acc_scale, b.x), fmaf(acc.y, P.acc_scale, b.y), fmaf(acc.z, P.acc_scale, b.z), fmaf(acc.w, P.acc_scale, b.w)};
            if (P.act == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.0f);
            }
            *reinterpret_cast<float4 *>(P.out_f32 + (pix[it] * (uint32_t)P.f32_ld + P.f32_off + co)) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) s[k] += o[k], q[k] = fmaf(o[k], o[k], q[k]);
        }
    }
#pragma unroll
    for (int m = 4; m < 32; m <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s[k] += __shfl_xor_sync(0xffffffffu, s[k], m);
            q[k] += __shfl_xor_sync(0xffffffffu, q[k], m);
        }
    }
    if (lane < 4 && co < P.cout) {
        float4 *dst = reinterpret_cast<float4 *>(slot + 2 * co);
        dst[0] = make_float4(s[0], q[0], s[1], q[1]);
        dst[1] = make_float4(s[2], q[2], s[3], q[3]);
    }
}

// mode 3 ("tap projection"): the activated outputs y[c] of a pixel are not stored; instead the epilogue evaluates the
// per-pixel part of a FOLLOWING 3x3 convolution with few output channels, p[t] = sum_c y[c] * w2[c][t] (t = tap * 2 + o,
// 18 values), in fp32 on the CUDA cores under the shadow of the next tile's MMAs.  rpe_tap_gather3x3 then adds the nine
// shifted p maps.  Used for RAFT's flow head (update.py:6-13: conv2(relu(conv1(x))), 256 -> 2 channels), whose second
// convolution would otherwise run as an N = 16 tensor-core GEMM at ~1 % utilisation.
constexpr int kCvProj = 18;
__device__ __forceinline__ void cv_project_chunk(const ConvParams &P, const uint32_t *v, const float *sbias, const float *w2s, int co,
                                                 float *acc) {
    // w2s: [256][16] floats (taps 0..15, read as four 16-byte broadcasts) followed by [256][2] (taps 16, 17)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float y = cv_activate_rt(fmaf(__uint_as_float(v[j]), P.acc_scale, sbias[co + j]), P.act) * P.scale;
        const float4 *w = reinterpret_cast<const float4 *>(w2s + (co + j) * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 wk = w[k];
            acc[4 * k] = fmaf(y, wk.x, acc[4 * k]);
            acc[4 * k + 1] = fmaf(y, wk.y, acc[4 * k + 1]);
            acc[4 * k + 2] = fmaf(y, wk.z, acc[4 * k + 2]);
            acc[4 * k + 3] = fmaf(y, wk.w, acc[4 * k + 3]);
        }
        const float2 wt = *reinterpret_cast<const float2 *>(w2s + 256 * 16 + (co + j) * 2);
        acc[16] = fmaf(y, wt.x, acc[16]);
        acc[17] = fmaf(y, wt.y, acc[17]);
    }
}

// kind 7: one chunk = 16 accumulator columns = target row `c` (0..15) of the unit's 16x16 target block (by, bx); after the
// transposition a lane holds, for 4 query pixels (it), the 4 targets x = bx*16 + 4*(lane & 3) .. +3 of that row.  Level 0 is
// written as it arrives; rows are paired in registers for the pooled levels with the reference's summation order
// avg_pool2d: ((a00 + a01) + a10 + a11) * 0.25, each level from the ROUNDED previous one (corr.py:25-27).  The x-neighbour of a
// level-2 value lives in lane ^ 1.  Every warp drains 8 consecutive rows, i.e. complete 8x8 groups.
struct CvCorrState {
    float4 prev[4];
    float2 l1p[4];
    float l2p[4];
};
template <typename T>
__device__ __forceinline__ void cv_corr_st(T *p, const T v, bool streaming) {
    if (streaming) __stcs(p, v);
    else *p = v;
}
__device__ __forceinline__ void cv_epilogue_corr(const ConvParams &P, const float *stage, int c, int by, int bx, const uint32_t *pix,
                                                 uint32_t inside_mask, int lane, CvCorrState &st) {
    const int h0 = P.corr_h, w0 = P.corr_w;
    const bool cs = (P.corr_vec & 16) != 0;
    const int gy = by * 16 + c, gx = bx * 16 + ((lane & 3) << 2);
    float4 o[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        const float4 a = *reinterpret_cast<const float4 *>(stage + r * 16 + (((lane & 3) ^ ((r >> 1) & 3)) << 2));
        o[it] = make_float4(a.x * P.scale, a.y * P.scale, a.z * P.scale, a.w * P.scale);
    }
    if (gy < h0 && gx < w0) {
        const size_t n0 = (size_t)h0 * w0;
#pragma unroll
        for (int it = 0; it < 4; ++it)
            if ((inside_mask >> it) & 1u) {
                float *p = P.corr_lvl[0] + (size_t)pix[it] * n0 + (size_t)gy * w0 + gx;
                // streaming stores (evict-first): the 139 MB per sample written here must not push the feature-map operands, which
                // every CTA re-reads 20 times, out of the L2
                if (P.corr_vec & 1) {
                    cv_corr_st(reinterpret_cast<float4 *>(p), o[it], cs);
                } else {
                    cv_corr_st(p, o[it].x, cs);
                    if (gx + 1 < w0) cv_corr_st(p + 1, o[it].y, cs);
                    if (gx + 2 < w0) cv_corr_st(p + 2, o[it].z, cs);
                    if (gx + 3 < w0) cv_corr_st(p + 3, o[it].w, cs);
                }
            }
    }
    if (!(c & 1)) {
#pragma unroll
        for (int it = 0; it < 4; ++it) st.prev[it] = o[it];
        return;
    }
    if (P.corr_levels < 2) return;
    const int h1 = h0 >> 1, w1 = w0 >> 1, y1 = gy >> 1, x1 = gx >> 1;
    float2 l1[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        l1[it].x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(st.prev[it].x, st.prev[it].y), o[it].x), o[it].y), 0.25f);
        l1[it].y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(st.prev[it].z, st.prev[it].w), o[it].z), o[it].w), 0.25f);
    }
    if (y1 < h1 && x1 < w1) {
        const size_t n1 = (size_t)h1 * w1;
#pragma unroll
        for (int it = 0; it < 4; ++it)
            if ((inside_mask >> it) & 1u) {
                float *p = P.corr_lvl[1] + (size_t)pix[it] * n1 + (size_t)y1 * w1 + x1;
                if (P.corr_vec & 2) {
                    cv_corr_st(reinterpret_cast<float2 *>(p), l1[it], cs);
                } else {
                    cv_corr_st(p, l1[it].x, cs);
                    if (x1 + 1 < w1) cv_corr_st(p + 1, l1[it].y, cs);
                }
            }
    }
    if ((c & 3) == 1) {
#pragma unroll
        for (int it = 0; it < 4; ++it) st.l1p[it] = l1[it];
        return;
    }
    if (P.corr_levels < 3) return;
    const int h2 = h0 >> 2, w2 = w0 >> 2, y2 = gy >> 2, x2 = gx >> 2;
    float l2[4];
#pragma unroll
    for (int it = 0; it < 4; ++it)
        l2[it] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(st.l1p[it].x, st.l1p[it].y), l1[it].x), l1[it].y), 0.25f);
    if (y2 < h2 && x2 < w2) {
        const size_t n2 = (size_t)h2 * w2;
#pragma unroll
        for (int it = 0; it < 4; ++it)
            if ((inside_mask >> it) & 1u) cv_corr_st(P.corr_lvl[2] + (size_t)pix[it] * n2 + (size_t)y2 * w2 + x2, l2[it], cs);
    }
    if ((c & 7) == 3) {
#pragma unroll
        for (int it = 0; it < 4; ++it) st.l2p[it] = l2[it];
        return;
    }
    if (P.corr_levels < 4) return;
    const int h3 = h0 >> 3, w3 = w0 >> 3, y3 = gy >> 3, x3 = gx >> 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const float top_r = __shfl_xor_sync(0xffffffffu, st.l2p[it], 1), bot_r = __shfl_xor_sync(0xffffffffu, l2[it], 1);
        if (!(lane & 1) && y3 < h3 && x3 < w3 && ((inside_mask >> it) & 1u)) {
            const float v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(st.l2p[it], top_r), l2[it]), bot_r), 0.25f);
            cv_corr_st(P.corr_lvl[3] + (size_t)pix[it] * ((size_t)h3 * w3) + (size_t)y3 * w3 + x3, v, cs);
        }
    }
}

struct CvTile {
    int nb, img, omin0, omaj0, lin;                    // lin: pixel-tile index over the whole batch
    bool ghost;                                        // pair mode: padding tile of an odd tile count (computed, never stored)
};
// Work unit u of CTA `rank`: a single tile, or (pair mode) one of two adjacent pixel tiles that share the weight block nb.
template <bool kPair>
__device__ __forceinline__ CvTile cv_decode(const ConvParams &P, int u, int rank) {
    CvTile t;
    t.nb = u % P.n_blocks;
    int t2 = u / P.n_blocks;
    const int per_img = P.tiles_min * P.tiles_maj;
    t.ghost = false;
    if (kPair) {
        t2 = 2 * t2 + rank;
        if (t2 >= P.N * per_img) t2 = P.N * per_img - 1, t.ghost = true;
    }
    t.lin = t2;
    t.img = t2 / per_img;
    const int tr = t2 - t.img * per_img;
    const int tj = tr / P.tiles_min;
    t.omin0 = (tr - tj * P.tiles_min) * 8;
    t.omaj0 = tj * 16;
    return t;
}

template <bool kPair>
__device__ __forceinline__ void cv_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (kPair) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        umma_f16(tmem_d, desc_a, desc_b, idesc, accumulate);
    }
}
// ---- issuer-side helpers on raw shared-memory addresses (no generic-pointer arithmetic on the critical path) ----
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    const long long t0 = clock64();
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) {
            printf("rpe_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    } while (!ok);
}
template <bool kPair>
__device__ __forceinline__ void cv_commit_u32(uint32_t bar) {
    if (kPair)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                     "h"((uint16_t)3)
                     : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart: constant upper word of the shared-memory descriptor
constexpr uint32_t kCvDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t cv_desc(uint32_t lo) { return ((uint64_t)kCvDescHi << 32) | (uint64_t)lo; }
// `ksteps` (1..4) MMAs of 16 channels each over one 64-channel K block; consecutive K steps are 32 bytes apart
template <bool kPair>
__device__ __forceinline__ void cv_mma_k(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, int ksteps) {
    if (ksteps == kCvBK / 16) {
        cv_mma<kPair>(d_tmem, cv_desc(a_lo), cv_desc(b_lo), idesc, accumulate);
        cv_mma<kPair>(d_tmem, cv_desc(a_lo + 2u), cv_desc(b_lo + 2u), idesc, 1u);
        cv_mma<kPair>(d_tmem, cv_desc(a_lo + 4u), cv_desc(b_lo + 4u), idesc, 1u);
        cv_mma<kPair>(d_tmem, cv_desc(a_lo + 6u), cv_desc(b_lo + 6u), idesc, 1u);
    } else {
        for (int k = 0; k < ksteps; ++k) {
            cv_mma<kPair>(d_tmem, cv_desc(a_lo + 2u * k), cv_desc(b_lo + 2u * k), idesc, accumulate);
            accumulate = 1u;
        }
    }
}
template <bool kPair>
__device__ __forceinline__ void cv_commit(uint64_t *bar) {
    if (kPair) umma_commit_pair(bar, 3);
    else umma_commit(bar);
}

__device__ __forceinline__ void cv_tmem_load16(uint32_t *v, uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    tmem_ld_wait();
}
// merged-N mode: the accumulator is the sum of two column ranges (a_hi * w_hi + a_lo * w_hi | a_hi * w_lo)
__device__ __forceinline__ void cv_tmem_load16_sum(uint32_t *v, uint32_t taddr, uint32_t second) {
    uint32_t w[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
          "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
        : "r"(taddr + second)
        : "memory");
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __uint_as_float(w[k]));
}
// the warp has read everything it needs from the accumulator stage: hand it back to the MMA issuer
template <bool kPair>
__device__ __forceinline__ void cv_release_acc(uint64_t *bar, uint32_t cluster_addr, int lane) {
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (kPair) mbar_arrive_cluster(cluster_addr);
        else mbar_arrive(bar);
    }
}

// k-th work unit of this CTA (pair).  Default: units first, first + step, ...  Correlation kind with a resident query tile:
// the CTA owns query-tile pairs first, first + step, ... and visits all n_blocks target blocks of each one consecutively.
template <int kK>
__device__ __forceinline__ int cv_unit(const ConvParams &P, int k, int first, int step) {
    if (kK == kKCorr && P.resident_a) return (first + (k / P.n_blocks) * step) * P.n_blocks + (k % P.n_blocks);
    return first + k * step;
}

// In pair mode (cta_group::2) two CTAs of a cluster own two adjacent pixel tiles: each loads its own activations and HALF of
// the weight tile, the leader (rank 0) issues M = 256 MMAs that read both shared memories and write both tensor memories,
// and every TMA load of either CTA counts its bytes on the leader's "full" barrier; the MMA commits are multicast to the
// "empty" / "accumulator full" barriers of both CTAs.
template <bool kPair, int kK>
__device__ __forceinline__ void conv_body(const ConvParams &P) {
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
    float *sstage = reinterpret_cast<float *>(smem + kCvSmemData + 512 + kCvMaxCout * 4);
    if (kK != kKCorr)      // (the correlation kind has no bias; its bn * n_blocks = number of target positions exceeds the staging area)
        for (int i = threadIdx.x; i < P.bn * P.n_blocks; i += kCvThreads) sbias[i] = (P.bias != nullptr && i < P.cout) ? P.bias[i] : 0.0f;
    // mode 3: projection weights [cout <= 256][18] fp32 behind the bias (rest of the bias area + the unused staging tiles)
    float *w2s = sbias + 256;
    if (kK == kKProj)      // [cout][18] in global -> [256][16] | [256][2] (see cv_project_chunk)
        for (int i = threadIdx.x; i < P.cout * kCvProj; i += kCvThreads) {
            const int c = i / kCvProj, k = i - c * kCvProj;
            w2s[k < 16 ? c * 16 + k : 256 * 16 + c * 2 + (k - 16)] = P.aux2[i];
        }

    // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int rank = kPair ? (int)cluster_ctarank() : 0;
    const int px_tiles = P.N * P.tiles_min * P.tiles_maj;
    const int num_units = (kPair ? (px_tiles + 1) / 2 : px_tiles) * P.n_blocks;
    const int u_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int u_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int a_units = P.reuse ? 1 : P.kmaj;        // slabs per filter column
    const uint32_t load_mult = kPair ? 2u : 1u;      // the leader's barriers count the bytes of both CTAs

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P.n_src; ++s)
            for (int p = 0; p < 2; ++p) {
                if (p < P.n_planes) prefetch_tmap(&P.amap[s][p]);
                if (p < P.w_planes) prefetch_tmap(&P.wmap[s][p]);
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
            mbar_init(&tmem_empty[i], kCvEpiWarps * load_mult);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (kPair) {
            tmem_alloc_pair(tmem_ptr, 512);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr, 512);
            tmem_relinquish();
        }
    }
    tcgen05_fence_before();
    __syncthreads();                     // CTA scope: barrier initialisation, bias staging and the TMEM base address written by tcgen05.alloc
    if (kPair) cluster_sync_all();       // cluster scope: the peer's barriers are initialised before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== activation producer =====================
        if (!(RPE_CV_DBG(P) & 1) && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0, u; (u = cv_unit<kK>(P, k, u_first, u_step)) < num_units; ++k) {
                const CvTile t = cv_decode<kPair>(P, u, rank);
                if (kK == kKCorr && P.resident_a && t.nb != 0) continue;        // the query tile is already resident
                for (int s = 0; s < P.n_src; ++s)
                    for (int cb = 0; cb < P.cblocks[s]; ++cb)
                        for (int tm = 0; tm < P.kmin; ++tm)
                            for (int au = 0; au < a_units; ++au) {
                                mbar_wait(&a_empty[stage], phase ^ 1);
                                if (rank == 0) mbar_expect_tx(&a_full[stage], P.a_plane_bytes * P.n_planes * load_mult);
                                const int cmin = t.omin0 * P.stride + tm - P.pad_min;
                                const int cmaj = t.omaj0 * P.stride + (P.reuse ? 0 : au) - P.pad_maj;
                                uint8_t *dst = sA + (size_t)stage * P.a_stage_bytes;
                                const int img_a = (kK == kKCorr && t.img >= P.corr_a_wrap) ? t.img - P.corr_a_sub : t.img;
                                for (int p = 0; p < P.n_planes; ++p) {
                                    if (kPair)
                                        tma_load_4d_pair(dst + p * P.a_plane_bytes, &P.amap[s][p], mapa_shared(smem_u32(&a_full[stage]), 0),
                                                         cb * kCvBK, cmin, cmaj, img_a);
                                    else
                                        tma_load_4d(dst + p * P.a_plane_bytes, &P.amap[s][p], &a_full[stage], cb * kCvBK, cmin, cmaj, img_a);
                                }
                                if (++stage == P.n_a_stages) stage = 0, phase ^= 1;
                            }
            }
        }
    } else if (warp == 3) {
        // ===================== weight producer =====================
        if (!(RPE_CV_DBG(P) & 1) && elect_one()) {
            const int taps = P.kmin * P.kmaj;
            const int brow = kPair ? rank * (P.bn >> 1) : 0;          // this CTA's half of the weight rows
            if (kPair && P.merge_n) {
                // Merged-N (pair mode only): per (K block, tap) this CTA holds X = all bn rows of ONE weight plane (rank 0: hi,
                // rank 1: lo) -- together the N = 2 bn operand [w_hi ; w_lo] of the a_hi product -- followed by Y = its half of
                // the hi plane, the N = bn operand of the a_lo product.
                const uint32_t x_bytes = (uint32_t)P.bn * 128u, tile_bytes = x_bytes + x_bytes / 2;
                if (P.resident_b) {
                    uint32_t total = 0;
                    for (int s = 0; s < P.n_src; ++s) total += (uint32_t)P.cblocks[s] * taps * tile_bytes;
                    if (u_first < num_units) {
                        if (rank == 0) mbar_expect_tx(&b_full[0], total * 2u);
                        for (int s = 0; s < P.n_src; ++s)
                            for (int cb = 0; cb < P.cblocks[s]; ++cb)
                                for (int tap = 0; tap < taps; ++tap) {
                                    uint8_t *dst = sB + (size_t)((P.cb_base[s] + cb) * taps + tap) * tile_bytes;
                                    const uint32_t bar = mapa_shared(smem_u32(&b_full[0]), 0);
                                    tma_load_3d_pair(dst, &P.wmapx[s][rank], bar, cb * kCvBK, 0, tap);
                                    tma_load_3d_pair(dst + x_bytes, &P.wmap[s][0], bar, cb * kCvBK, brow, tap);
                                }
                    }
                } else {
                    int stage = 0, in_entry = 0;
                    uint32_t phase = 0;
                    const uint32_t entry_bytes = tile_bytes * (uint32_t)P.b_group;
                    for (int k = 0, u; (u = cv_unit<kK>(P, k, u_first, u_step)) < num_units; ++k)
                        for (int s = 0; s < P.n_src; ++s)
                            for (int cb = 0; cb < P.cblocks[s]; ++cb)
                                for (int tm = 0; tm < P.kmin; ++tm)
                                    for (int tj = 0; tj < P.kmaj; ++tj) {
                                        const int tap = P.orient == 0 ? tj * P.kw + tm : tm * P.kw + tj;
                                        if (in_entry == 0) {
                                            mbar_wait(&b_empty[stage], phase ^ 1);
                                            if (rank == 0) mbar_expect_tx(&b_full[stage], entry_bytes * 2u);
                                        }
                                        uint8_t *dst = sB + (size_t)stage * entry_bytes + (size_t)in_entry * tile_bytes;
                                        const uint32_t bar = mapa_shared(smem_u32(&b_full[stage]), 0);
                                        tma_load_3d_pair(dst, &P.wmapx[s][rank], bar, cb * kCvBK, 0, tap);
                                        tma_load_3d_pair(dst + x_bytes, &P.wmap[s][0], bar, cb * kCvBK, brow, tap);
                                        if (++in_entry == P.b_group) {
                                            in_entry = 0;
                                            if (++stage == P.n_b_stages) stage = 0, phase ^= 1;
                                        }
                                    }
                }
            } else if (P.resident_b) {
                // the whole weight set fits: one load per CTA, [source K block][tap][plane] tiles
                uint32_t total = 0;
                for (int s = 0; s < P.n_src; ++s) total += (uint32_t)P.cblocks[s] * taps * P.w_planes * P.b_plane_bytes;
                if (u_first < num_units) {
                    if (rank == 0) mbar_expect_tx(&b_full[0], total * load_mult);
                    for (int s = 0; s < P.n_src; ++s)
                        for (int cb = 0; cb < P.cblocks[s]; ++cb)
                            for (int tap = 0; tap < taps; ++tap)
                                for (int p = 0; p < P.w_planes; ++p) {
                                    uint8_t *dst = sB + (size_t)(((P.cb_base[s] + cb) * taps + tap) * P.w_planes + p) * P.b_plane_bytes;
                                    if (kPair) tma_load_3d_pair(dst, &P.wmap[s][p], mapa_shared(smem_u32(&b_full[0]), 0), cb * kCvBK, brow, tap);
                                    else tma_load_3d(dst, &P.wmap[s][p], &b_full[0], cb * kCvBK, 0, tap);
                                }
                }
            } else {
                // ring entries of b_group plane tiles (one barrier round trip per entry): a single tile for the wide layers, the
                // planes of a tap or of a whole filter column for the narrow ones, whose MMAs are too short to hide per-tile barriers
                int stage = 0, in_entry = 0;
                uint32_t phase = 0;
                const uint32_t entry_bytes = P.b_plane_bytes * (uint32_t)P.b_group;
                for (int k = 0, u; (u = cv_unit<kK>(P, k, u_first, u_step)) < num_units; ++k) {
                    const int nb = u % P.n_blocks;
                    const int img = kK == kKCorr ? cv_decode<kPair>(P, u, rank).img : 0;
                    const int cby = kK == kKCorr ? nb / P.corr_nbx : 0, cbx = kK == kKCorr ? nb - cby * P.corr_nbx : 0;
                    for (int s = 0; s < P.n_src; ++s)
                        for (int cb = 0; cb < P.cblocks[s]; ++cb)
                            for (int tm = 0; tm < P.kmin; ++tm)
                                for (int tj = 0; tj < P.kmaj; ++tj) {
                                    const int tap = P.orient == 0 ? tj * P.kw + tm : tm * P.kw + tj;
                                    for (int p = 0; p < P.w_planes; ++p) {
                                        if (in_entry == 0) {
                                            mbar_wait(&b_empty[stage], phase ^ 1);
                                            if (rank == 0) mbar_expect_tx(&b_full[stage], entry_bytes * load_mult);
                                        }
                                        uint8_t *dst = sB + (size_t)stage * entry_bytes + (size_t)in_entry * P.b_plane_bytes;
                                        if (kK == kKCorr) {
                                            // 16x16 block of target positions of image `img` (this CTA's 8 rows of it in pair mode): the
                                            // box lands as [row][x][64 ch] = accumulator column row * 16 + x; outside the map reads zero
                                            if (kPair)
                                                tma_load_4d_pair(dst, &P.wmap[s][p], mapa_shared(smem_u32(&b_full[stage]), 0), cb * kCvBK, cbx * 16,
                                                                 cby * 16 + rank * 8, img);
                                            else
                                                tma_load_4d(dst, &P.wmap[s][p], &b_full[stage], cb * kCvBK, cbx * 16, cby * 16, img);
                                        } else if (kPair)
                                            tma_load_3d_pair(dst, &P.wmap[s][p], mapa_shared(smem_u32(&b_full[stage]), 0), cb * kCvBK,
                                                             nb * P.bn + brow, tap);
                                        else
                                            tma_load_3d(dst, &P.wmap[s][p], &b_full[stage], cb * kCvBK, nb * P.bn, tap);
                                        if (++in_entry == P.b_group) {
                                            in_entry = 0;
                                            if (++stage == P.n_b_stages) stage = 0, phase ^= 1;
                                        }
                                    }
                                }
                }
            }
        }
    } else if (warp == 2 && (kK == kKState || kK == kKGates) && P.side_prefetch) {
        // ===================== side-input prefetcher (GRU epilogues) =====================
        // The GRU epilogues read 1.0-1.5 KB of fp32 side inputs per pixel (hoisted context term, update gate, hidden state) that were
        // written by earlier kernels and left the L2 since.  When the accumulator of unit k is complete (its epilogue starts), this
        // otherwise idle warp asks the L2 for the side inputs of unit k + 1: one bulk prefetch per pixel row of the tile and tensor.
        const int rows = P.orient == 0 ? 16 : 8, cols = P.orient == 0 ? 8 : 16;       // image rows / contiguous pixels of a tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int k = 0, u; (u = cv_unit<kK>(P, k, u_first, u_step)) < num_units; ++k) {
            const int un = cv_unit<kK>(P, k + 1, u_first, u_step);
            mbar_wait(&tmem_full[acc], acc_phase);
            if (++acc == kCvAcc) acc = 0, acc_phase ^= 1;
            if (un >= num_units) break;
            const CvTile t = cv_decode<kPair>(P, un, rank);
            if (t.ghost || lane >= rows) continue;
            const int y = (P.orient == 0 ? t.omaj0 : t.omin0) + lane, x0 = P.orient == 0 ? t.omin0 : t.omaj0;
            if (y >= P.OH || x0 >= P.OW) continue;
            const int npx = min(cols, P.OW - x0);
            const size_t pix = ((size_t)t.img * P.OH + y) * P.OW + x0;
            prefetch_l2_bulk(P.pre + pix * P.pre_ld, (uint32_t)(npx * P.pre_ld * 4));
            prefetch_l2_bulk(P.aux + pix * P.aux_ld, (uint32_t)(npx * P.aux_ld * 4));
            if (kK == kKState) prefetch_l2_bulk(P.aux2 + pix * P.aux2_ld, (uint32_t)(npx * P.aux2_ld * 4));
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (the leader CTA in pair mode) =====================
        // The WHOLE warp walks the loop nest, so that every address / descriptor / barrier computation is warp-uniform (uniform
        // registers, which is where tcgen05.mma takes its descriptors from); only the MMAs and commits are predicated on one
        // elected lane.  The issuer's instruction stream is the critical path of the kernel: one 128-cycle MMA per ~20 issue slots.
        const bool leader = elect_one();
        // instruction descriptor: D fp32 (c_format 1), A/B fp16 (a/b_format 0), both K-major, N = bn, M = 128 (256 across a CTA pair)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(P.bn >> 3) << 17) |
                               ((uint32_t)((kPair ? 256 : 128) >> 4) << 24);
        const bool no_load = (RPE_CV_DBG(P) & 1) != 0, no_mma = (RPE_CV_DBG(P) & 2) != 0;
        // products per tap: a_hi * w_hi, + a_lo * w_hi with two activation planes, + a_hi * w_lo with two weight planes
        const bool a2 = P.n_planes == 2, w2 = P.w_planes == 2, reuse = P.reuse != 0, resident = P.resident_b != 0;
        // shared-memory descriptors = constant upper word | (address >> 4); ring positions advance the low word only
        const uint32_t a_base16 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16), b_base16 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t a_stage16 = P.a_stage_bytes >> 4, a_plane16 = P.a_plane_bytes >> 4, b_plane16 = P.b_plane_bytes >> 4;
        const uint32_t a_full_u = smem_u32(a_full), a_empty_u = smem_u32(a_empty), b_full_u = smem_u32(b_full), b_empty_u = smem_u32(b_empty);
        const uint32_t t_full_u = smem_u32(tmem_full), t_empty_u = smem_u32(tmem_empty);
        const int n_a = P.n_a_stages, n_b = P.n_b_stages, kmin = P.kmin, kmaj = P.kmaj, taps = P.kmin * P.kmaj, b_group = P.b_group;
        const uint32_t b_entry16 = b_plane16 * (uint32_t)P.b_group;
        const bool merged = kPair && P.merge_n != 0;
        const uint32_t m_x16 = ((uint32_t)P.bn * 128u) >> 4, m_tile16 = m_x16 + (m_x16 >> 1);         // merged-N tile: X (bn rows) | Y (bn / 2 rows)
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)(P.bn >> 2) << 17) | ((uint32_t)((kPair ? 256 : 128) >> 4) << 24);   // N = 2 bn
        uint32_t sa = 0, sb = 0, pa = 0, pb = 0, acc = 0, acc_phase = 0;
        bool b_ready = no_load;
        for (int k = 0, u; (u = cv_unit<kK>(P, k, u_first, u_step)) < num_units; ++k) {
            // resident query tile (kind 7): its stages are released (and the parity advances) only after the last target block
            const bool a_release = !(kK == kKCorr && P.resident_a) || (u % P.n_blocks) == P.n_blocks - 1;
            mbar_wait_u32(t_empty_u + acc * 8u, acc_phase ^ 1u);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * (uint32_t)kCvMaxBN;
            uint32_t accumulate = 0;
            for (int s = 0; s < P.n_src; ++s) {
                const int ncb = P.cblocks[s];
                for (int cb = 0; cb < ncb; ++cb) {
                    const int ksteps = (cb == ncb - 1) ? P.ksteps_last[s] : kCvBK / 16;
                    for (int tm = 0; tm < kmin; ++tm)
                        for (int tj = 0; tj < kmaj; ++tj) {
                            if ((!reuse || tj == 0) && !no_load) mbar_wait_u32(a_full_u + sa * 8u, pa);
                            const uint32_t a_hi = a_base16 + sa * a_stage16 + (reuse ? (uint32_t)tj * 64u : 0u);   // tap shift: 1024 B
                            if (merged) {
                                // two MMAs per K step instead of three: a_hi x [w_hi ; w_lo] (N = 2 bn, columns [0, bn) and [bn, 2 bn))
                                // and a_lo x w_hi (N = bn, columns [0, bn)); the epilogue adds the two column ranges
                                uint32_t b_x;
                                if (resident) {
                                    if (!b_ready) {
                                        mbar_wait_u32(b_full_u, 0);
                                        b_ready = true;
                                    }
                                    const int tap = P.orient == 0 ? tj * P.kw + tm : tm * P.kw + tj;
                                    b_x = b_base16 + (uint32_t)((P.cb_base[s] + cb) * taps + tap) * m_tile16;
                                } else {
                                    const bool col = b_group != 1;
                                    if ((!col || tj == 0) && !no_load) mbar_wait_u32(b_full_u + sb * 8u, pb);
                                    b_x = b_base16 + sb * m_tile16 * (uint32_t)b_group + (col ? (uint32_t)tj * m_tile16 : 0u);
                                }
                                tcgen05_fence_after();
                                if (leader && !no_mma) {
                                    cv_mma_k<kPair>(d_tmem, a_hi, b_x, idesc2, accumulate, ksteps);
                                    cv_mma_k<kPair>(d_tmem, a_hi + a_plane16, b_x + m_x16, idesc, 1u, ksteps);
                                }
                                if (!resident && (b_group == 1 || tj == kmaj - 1)) {
                                    if (leader && !no_load) cv_commit_u32<kPair>(b_empty_u + sb * 8u);
                                    if (++sb == (uint32_t)n_b) sb = 0, pb ^= 1u;
                                }
                            } else if (resident) {
                                if (!b_ready) {
                                    mbar_wait_u32(b_full_u, 0);
                                    b_ready = true;
                                }
                                tcgen05_fence_after();
                                const int tap = P.orient == 0 ? tj * P.kw + tm : tm * P.kw + tj;
                                const uint32_t b_hi = b_base16 + (uint32_t)(((P.cb_base[s] + cb) * taps + tap) * P.w_planes) * b_plane16;
                                if (leader && !no_mma) {
                                    cv_mma_k<kPair>(d_tmem, a_hi, b_hi, idesc, accumulate, ksteps);
                                    if (a2) cv_mma_k<kPair>(d_tmem, a_hi + a_plane16, b_hi, idesc, 1u, ksteps);
                                    if (w2) cv_mma_k<kPair>(d_tmem, a_hi, b_hi + b_plane16, idesc, 1u, ksteps);
                                }
                            } else if (b_group == 1) {
                                if (!no_load) mbar_wait_u32(b_full_u + sb * 8u, pb);
                                tcgen05_fence_after();
                                const uint32_t b_hi = b_base16 + sb * b_plane16;
                                if (leader) {
                                    if (!no_mma) {
                                        cv_mma_k<kPair>(d_tmem, a_hi, b_hi, idesc, accumulate, ksteps);
                                        if (a2) cv_mma_k<kPair>(d_tmem, a_hi + a_plane16, b_hi, idesc, 1u, ksteps);
                                    }
                                    if (!no_load) cv_commit_u32<kPair>(b_empty_u + sb * 8u);
                                }
                                if (++sb == (uint32_t)n_b) sb = 0, pb ^= 1u;
                                if (w2) {
                                    if (!no_load) mbar_wait_u32(b_full_u + sb * 8u, pb);
                                    tcgen05_fence_after();
                                    if (leader) {
                                        if (!no_mma) cv_mma_k<kPair>(d_tmem, a_hi, b_base16 + sb * b_plane16, idesc, 1u, ksteps);
                                        if (!no_load) cv_commit_u32<kPair>(b_empty_u + sb * 8u);
                                    }
                                    if (++sb == (uint32_t)n_b) sb = 0, pb ^= 1u;
                                }
                            } else {
                                // grouped entries: both planes of the tap (or of every tap of the filter column) behind one barrier
                                const bool col = b_group != (w2 ? 2 : 1);
                                if ((!col || tj == 0) && !no_load) mbar_wait_u32(b_full_u + sb * 8u, pb);
                                tcgen05_fence_after();
                                const uint32_t b_hi = b_base16 + sb * b_entry16 + (col ? (uint32_t)tj * (w2 ? 2u : 1u) * b_plane16 : 0u);
                                if (leader && !no_mma) {
                                    cv_mma_k<kPair>(d_tmem, a_hi, b_hi, idesc, accumulate, ksteps);
                                    if (a2) cv_mma_k<kPair>(d_tmem, a_hi + a_plane16, b_hi, idesc, 1u, ksteps);
                                    if (w2) cv_mma_k<kPair>(d_tmem, a_hi, b_hi + b_plane16, idesc, 1u, ksteps);
                                }
                                if (!col || tj == kmaj - 1) {
                                    if (leader && !no_load) cv_commit_u32<kPair>(b_empty_u + sb * 8u);
                                    if (++sb == (uint32_t)n_b) sb = 0, pb ^= 1u;
                                }
                            }
                            accumulate = 1;
                            if (!reuse || tj == kmaj - 1) {
                                if (leader && !no_load && a_release) cv_commit_u32<kPair>(a_empty_u + sa * 8u);
                                if (++sa == (uint32_t)n_a) {
                                    sa = 0;
                                    if (a_release) pa ^= 1u;
                                }
                            }
                        }
                }
            }
            if (leader) cv_commit_u32<kPair>(t_full_u + acc * 8u);
            if (++acc == (uint32_t)kCvAcc) acc = 0, acc_phase ^= 1u;
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // TMEM lane = pixel row of the tile.  Every 16 accumulator columns are transposed through a 2 KB per-warp staging
        // tile so that the global side loads and the stores are coalesced (64 contiguous bytes per pixel per instruction).
        const int wq = warp & 3;                               // TMEM lane quarter this warp may access
        const int chalf = (warp - 4) >> 2;                     // which half of the accumulator columns it drains
        float *stage = sstage + (warp - 4) * (32 * 16);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int k = 0, u; (u = cv_unit<kK>(P, k, u_first, u_step)) < num_units; ++k) {
            const CvTile t = cv_decode<kPair>(P, u, rank);
            uint32_t pix[4];
            uint32_t inside_mask = 0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int m = wq * 32 + it * 8 + (lane >> 2);      // accumulator row handled in the transposed mapping
                const int gi = m >> 3, mi = m & 7;                 // (major, minor) position inside the tile
                const int y = P.orient == 0 ? t.omaj0 + gi : t.omin0 + mi;
                const int x = P.orient == 0 ? t.omin0 + mi : t.omaj0 + gi;
                if (y < P.OH && x < P.OW && !t.ghost) inside_mask |= 1u << it;
                pix[it] = (uint32_t)((t.img * P.OH + y) * P.OW + x);
            }
            const int n_chunks = P.bn / 16;
            const int c_begin = chalf * ((n_chunks + 1) / 2), c_end = chalf == 0 ? (n_chunks + 1) / 2 : n_chunks;
            const uint32_t empty_addr = kPair ? mapa_shared(smem_u32(&tmem_empty[acc]), 0) : 0u;
            const uint32_t taddr0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * kCvMaxBN);
            if (kK == kKProj) {
                // ---- tap projection: lane = accumulator row = pixel, no transposition, nothing but 18 partial sums is stored
                float proj[kCvProj];
#pragma unroll
                for (int k = 0; k < kCvProj; ++k) proj[k] = 0.0f;
                mbar_wait(&tmem_full[acc], acc_phase);
                tcgen05_fence_after();
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t v[16];
                    cv_tmem_load16(v, taddr0 + (uint32_t)(c * 16));
                    if (c + 1 == c_end) cv_release_acc<kPair>(&tmem_empty[acc], empty_addr, lane);
                    if (!(RPE_CV_DBG(P) & 4)) cv_project_chunk(P, v, sbias, w2s, t.nb * P.bn + c * 16, proj);
                }
                const int m = wq * 32 + lane;
                const int gi = m >> 3, mi = m & 7;
                const int y = P.orient == 0 ? t.omaj0 + gi : t.omin0 + mi;
                const int x = P.orient == 0 ? t.omin0 + mi : t.omaj0 + gi;
                if (y < P.OH && x < P.OW && !t.ghost && !(RPE_CV_DBG(P) & 4)) {
                    float2 *dst = reinterpret_cast<float2 *>(P.out_f32 + (((size_t)t.img * P.OH + y) * P.OW + x) * P.f32_ld + P.f32_off +
                                                             chalf * kCvProj);
#pragma unroll
                    for (int k = 0; k < kCvProj / 2; ++k) dst[k] = make_float2(proj[2 * k], proj[2 * k + 1]);
                }
            } else {
                const int co_lane = t.nb * P.bn + ((lane & 3) << 2);
                const int cby = kK == kKCorr ? t.nb / P.corr_nbx : 0, cbx = kK == kKCorr ? t.nb - cby * P.corr_nbx : 0;
                CvCorrState cst;
                // side inputs of the first chunk are requested before the accumulator is even complete
                CvSide sd[4], sd_next[4];
                if (c_begin < c_end && !(RPE_CV_DBG(P) & 4)) cv_side_load4<kK>(P, sd, co_lane + c_begin * 16, pix, inside_mask);
                mbar_wait(&tmem_full[acc], acc_phase);
                tcgen05_fence_after();
                if (c_begin >= c_end) cv_release_acc<kPair>(&tmem_empty[acc], empty_addr, lane);   // nothing to drain (bn = 16)
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t v[16];
                    if (kPair && kK != kKCorr && kK != kKGates && kK != kKState && P.merge_n) cv_tmem_load16_sum(v, taddr0 + (uint32_t)(c * 16), (uint32_t)P.bn);
                    else cv_tmem_load16(v, taddr0 + (uint32_t)(c * 16));
                    if (c + 1 == c_end) cv_release_acc<kPair>(&tmem_empty[acc], empty_addr, lane);
                    if (RPE_CV_DBG(P) & 4) continue;
                    if (c + 1 < c_end) cv_side_load4<kK>(P, sd_next, co_lane + (c + 1) * 16, pix, inside_mask);
                    __syncwarp();                                  // previous chunk's reads of the staging tile are done
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<uint4 *>(stage + lane * 16 + ((k ^ ((lane >> 1) & 3)) << 2)) =
                            make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    __syncwarp();
                    if (kK == kKCorr) {
                        cv_epilogue_corr(P, stage, c, cby, cbx, pix, inside_mask, lane, cst);
                    } else if (kK == kKF32Stats) {
                        // ghost tiles (pair padding) have inside_mask = 0 and alias the last real tile: they must not write a slot
                        float *slot = P.stat_part + ((size_t)t.lin * 4 + wq) * (size_t)(2 * P.bn * P.n_blocks);
                        if (!t.ghost) cv_epilogue_half_stats(P, stage, sbias, co_lane + c * 16, pix, inside_mask, lane, slot);
                    } else {
                        cv_epilogue_half<kK>(P, stage, sbias, co_lane + c * 16, pix, inside_mask, lane, sd);
                    }
                    if (kK == kKGeneric || kK == kKGates || kK == kKState) {
#pragma unroll
                        for (int it = 0; it < 4; ++it) sd[it] = sd_next[it];
                    }
                }
            }
            if (++acc == kCvAcc) acc = 0, acc_phase ^= 1;
        }
    }
    tcgen05_fence_before();
    if (kPair) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        if (kPair) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// One instantiation per epilogue mode: each kernel carries only its own epilogue (the instruction footprint of the three
// concurrently running roles has to stay inside the instruction cache).
template <int kMode>
__global__ void __launch_bounds__(kCvThreads, 1) conv_f16x3_kernel(const __grid_constant__ ConvParams P) {
    conv_body<false, kMode>(P);
}

template <int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kCvThreads, 1) conv_f16x3_pair_kernel(const __grid_constant__ ConvParams P) {
    conv_body<true, kMode>(P);
}

template <int kMode>
static cudaError_t cv_launch(const ConvParams &p, int grid, bool pair, cudaStream_t stream, bool set_attr) {
    if (set_attr) {
        cudaError_t e = cudaFuncSetAttribute(conv_f16x3_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_f16x3_pair_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmem);
        return e;
    }
    if (pair) conv_f16x3_pair_kernel<kMode><<<grid, kCvThreads, kCvSmem, stream>>>(p);
    else conv_f16x3_kernel<kMode><<<grid, kCvThreads, kCvSmem, stream>>>(p);
    return cudaSuccess;
}
static cudaError_t cv_dispatch(int kind, const ConvParams &p, int grid, bool pair, cudaStream_t stream, bool set_attr) {
    switch (kind) {
        case kKGates: return cv_launch<kKGates>(p, grid, pair, stream, set_attr);
        case kKState: return cv_launch<kKState>(p, grid, pair, stream, set_attr);
        case kKProj: return cv_launch<kKProj>(p, grid, pair, stream, set_attr);
        case kKPlanes: return cv_launch<kKPlanes>(p, grid, pair, stream, set_attr);
        case kKF32: return cv_launch<kKF32>(p, grid, pair, stream, set_attr);
        case kKF32Stats: return cv_launch<kKF32Stats>(p, grid, pair, stream, set_attr);
        case kKCorr: return cv_launch<kKCorr>(p, grid, pair, stream, set_attr);
        default: return cv_launch<kKGeneric>(p, grid, pair, stream, set_attr);
    }
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

// Probe / A-B switches from the environment, read once per process (not on every plan creation).
struct CvEnv {
    bool no_pair, generic, corr_resident_a, corr_plain_stores;
    int a_stages, dbg, b_group, merge_max;
    bool side_prefetch;
    CvEnv() {
        const char *e = getenv("RPE_CONV_PAIR");
        no_pair = e && e[0] == '0';
        generic = getenv("RPE_CONV_GENERIC") != nullptr;
        corr_plain_stores = getenv("RPE_CORR_PLAIN_STORES") != nullptr;   // A-B switch: default cache operator for the pyramid stores
        corr_resident_a = getenv("RPE_CORR_RESIDENT_A") != nullptr;  // A-B switch: keep the query tile of the correlation kernel resident
        e = getenv("RPE_CONV_ASTAGES");
        a_stages = e ? atoi(e) : 0;
        e = getenv("RPE_CONV_DEBUG");
        dbg = e ? atoi(e) : 0;
        // A-B switch, off by default: L2 prefetch of the GRU epilogues' side inputs one unit ahead by the idle warp.  Measured on B200
        // (profiles/r2_y_conv_probe_side_prefetch.txt, 64 samples): q1 314 -> 337 us, zr1 460 -> 471 us -- the prefetches compete with
        // the operand loads for the same L2 / HBM bandwidth and the epilogue's own loads were not the exposed latency.
        e = getenv("RPE_CONV_SIDE_PREFETCH");
        side_prefetch = e && e[0] == '1';
        e = getenv("RPE_CONV_MERGE");           // A-B switch: largest bn that runs in merged-N mode (0 = off); default 128
        merge_max = e ? atoi(e) : 128;
        e = getenv("RPE_CONV_BGROUP");          // A-B switch: weight-ring entries of 1 = a plane tile, 2 = a tap, 3 = a filter column; unset = auto
        b_group = e ? atoi(e) : 0;
    }
};
static const CvEnv &cv_env() {
    static const CvEnv env;
    return env;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device and per kernel instantiation.
static cudaError_t cv_ensure_attr(int kind, const ConvParams &p) {
    static bool done[16][8] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 16 && done[dev][kind]) return cudaSuccess;
    e = cv_dispatch(kind, p, 0, false, nullptr, true);
    if (e == cudaSuccess && dev >= 0 && dev < 16) done[dev][kind] = true;
    return e;
}

struct ConvPlan {
    ConvParams p;
    int grid;
    bool pair;             // CTA-pair kernel (cta_group::2)
    int kind;              // kernel kind (kK*): which epilogue instantiation runs
    int tiles_per_image;   // 128-pixel tiles per image (instance-norm partial sums: 4 slots per tile)
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
    if ((d->res_hi == nullptr) != (d->res_lo == nullptr) || (d->res_hi && d->res)) return RPE_ERR_INVALID_ARG;
    if (d->res_hi && ((d->res_ld % 4) || (reinterpret_cast<uintptr_t>(d->res_hi) & 7u) || (reinterpret_cast<uintptr_t>(d->res_lo) & 7u)))
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
    p.w_planes = d->src[0].w_lo ? 2 : 1;              // (1 activation plane x 2 weight planes: exact 16-bit inputs such as raw uint8 frames)
    // CTA pairs (cta_group::2): each CTA of a 2-cluster loads half of every weight tile.  RPE_CONV_PAIR=0 disables.
    const int px_tiles = d->N * (((orient == 0 ? OW : OH) + 7) / 8) * (((orient == 0 ? OH : OW) + 15) / 16);
    pl->pair = !cv_env().no_pair && (bn % 32 == 0) && px_tiles >= 2 && sm_count() >= 2;
    const int b_rows = pl->pair ? bn / 2 : bn;
    p.a_plane_bytes = (uint32_t)slab_rows * 8 * 128;
    p.a_stage_bytes = p.a_plane_bytes * p.n_planes;
    p.b_plane_bytes = (uint32_t)b_rows * 128;
    p.b_stage_bytes = p.b_plane_bytes * p.w_planes;
    const int taps = d->kh * d->kw;
    // Merged-N (narrow layers with streamed weights, CTA pairs, plain epilogues): the hi and lo weight planes form ONE operand of
    // 2 bn columns, so that a K step costs two MMAs (N = 2 bn, N = bn) instead of three of N = bn: fewer, longer MMAs and fewer
    // activation-operand reads per flop.  Measured (profiles/r2_v_conv_probe_merged_n.txt): 96 -> 96 +18 %, 128 -> 64 +8 %,
    // 256 -> 126 +5 %.  RPE_CONV_MERGE=0 disables, =<n> sets the largest bn.  Per (K block, tap) a CTA then holds bn + bn / 2
    // weight rows instead of 2 * bn / 2.
    p.merge_n = (pl->pair && p.n_planes == 2 && p.w_planes == 2 && n_blocks == 1 && d->mode == 0 && bn <= cv_env().merge_max) ? 1 : 0;
    // Weights resident in shared memory when the whole set fits beside two activation stages (small layers: every tile would
    // otherwise re-stream them); else a ring whose entries are single plane tiles, so that many small loads are in flight.
    size_t k_tiles = 0;
    for (int s = 0; s < d->n_sources; ++s) k_tiles += (size_t)((d->src[s].c_count + kCvBK - 1) / kCvBK) * taps;
    const size_t room = (size_t)kCvSmemData - 2 * (size_t)p.a_stage_bytes;
    if (p.merge_n && k_tiles * 2 * p.b_plane_bytes <= room)
        p.merge_n = 0;            // layers whose plain weight set stays resident gain nothing (64 -> 64: 1061 vs 1038 TFLOP/s) and would lose an activation stage
    const size_t tile_b = p.merge_n ? 3 * (size_t)p.b_plane_bytes : (size_t)p.w_planes * p.b_plane_bytes;      // per (K block, tap)
    const size_t total_b = k_tiles * tile_b;
    p.resident_b = (n_blocks == 1 && total_b + 2 * (size_t)p.a_stage_bytes <= (size_t)kCvSmemData) ? 1 : 0;
    if (p.resident_b) {
        p.n_a_stages = (int)(((size_t)kCvSmemData - total_b) / p.a_stage_bytes);
        if (p.n_a_stages > kCvMaxAStages) p.n_a_stages = kCvMaxAStages;
        p.n_b_stages = 1;
        p.b_group = 1;
    } else {
        p.n_a_stages = 2;
        if (3 * (size_t)p.a_stage_bytes + 8 * (size_t)p.b_plane_bytes <= (size_t)kCvSmemData) p.n_a_stages = 3;
        if (cv_env().a_stages > 0) {      // probe only
            const int v = cv_env().a_stages;
            if (v >= 1 && v <= kCvMaxAStages && (size_t)v * p.a_stage_bytes + 2 * (size_t)p.b_plane_bytes <= (size_t)kCvSmemData) p.n_a_stages = v;
        }
        // Ring entries: narrow layers (short MMAs) cannot hide one barrier round trip per plane tile in the issuer's instruction
        // stream, so their entries hold both planes of a tap, or of all taps of a filter column when three such entries fit.
        const size_t ring = (size_t)kCvSmemData - (size_t)p.n_a_stages * p.a_stage_bytes;
        if (p.merge_n) {         // entries of whole (X | Y) tiles: a filter column when three of them fit, else a tap
            p.b_group = (cv_env().b_group != 1 && cv_env().b_group != 2 && ring / ((size_t)p.kmaj * tile_b) >= 3) ? p.kmaj : 1;
            p.n_b_stages = (int)(ring / ((size_t)p.b_group * tile_b));
        } else {
            const int tap_tiles = p.w_planes, col_tiles = p.kmaj * p.w_planes;
            int mode = cv_env().b_group;
            if (mode == 0) mode = p.b_plane_bytes <= kCvGroupMaxTile ? 3 : 1;
            if (mode == 3 && ring / ((size_t)col_tiles * p.b_plane_bytes) < 3) mode = 2;
            if (mode == 2 && ring / ((size_t)tap_tiles * p.b_plane_bytes) < 3) mode = 1;
            p.b_group = mode == 3 ? col_tiles : mode == 2 ? tap_tiles : 1;
            p.n_b_stages = (int)(ring / ((size_t)p.b_group * p.b_plane_bytes));
        }
        if (p.n_b_stages > kCvMaxBStages) p.n_b_stages = kCvMaxBStages;
        if (p.n_b_stages < 2) {
            delete pl;
            return RPE_ERR_INVALID_ARG;
        }
    }
    double macs = 0.0;
    for (int s = 0; s < d->n_sources; ++s) {
        const rpe_conv_source &sc = d->src[s];
        if (!sc.act_hi || !sc.w_hi || ((sc.act_lo != nullptr) != (p.n_planes == 2)) || ((sc.w_lo != nullptr) != (p.w_planes == 2)) ||
            sc.c_count <= 0 || (sc.c_count % 16) || (sc.c_offset % 8) || (sc.c_total % 8) || sc.c_offset + sc.c_count > sc.c_total ||
            sc.w_cstride < sc.c_count || (sc.w_cstride % 8)) {
            delete pl;
            return RPE_ERR_INVALID_ARG;
        }
        p.cblocks[s] = (sc.c_count + kCvBK - 1) / kCvBK;
        p.cb_base[s] = s == 0 ? 0 : p.cb_base[s - 1] + p.cblocks[s - 1];
        p.ksteps_last[s] = (sc.c_count - (p.cblocks[s] - 1) * kCvBK) / 16;
        macs += (double)sc.c_count * taps;
        for (int pln = 0; pln < 2; ++pln) {
            const void *act = pln == 0 ? sc.act_hi : sc.act_lo;
            const void *wgt = pln == 0 ? sc.w_hi : sc.w_lo;
            if ((reinterpret_cast<uintptr_t>(act) & 15u) || (reinterpret_cast<uintptr_t>(wgt) & 15u)) {
                delete pl;
                return RPE_ERR_ALIGNMENT;
            }
            if (pln < p.n_planes) {   // activations: (C, minor, major, N) fp16; the channel window starts at c_offset, channels beyond it read as zero
                const char *base = reinterpret_cast<const char *>(act) + (size_t)sc.c_offset * 2;
                const cuuint64_t pix = (cuuint64_t)sc.c_total * 2, row = pix * d->W;
                cuuint64_t dims[4] = {(cuuint64_t)sc.c_count, (cuuint64_t)(orient == 0 ? d->W : d->H),
                                      (cuuint64_t)(orient == 0 ? d->H : d->W), (cuuint64_t)d->N};
                cuuint64_t strides[3] = {orient == 0 ? pix : row, orient == 0 ? row : pix, row * d->H};
                cuuint32_t box[4] = {kCvBK, (cuuint32_t)(8 * stride), (cuuint32_t)(slab_rows * stride), 1};
                cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
                CUresult r = g_cv_encode(&p.amap[s][pln], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<char *>(base), dims, strides, box,
                                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    g_last_cuda_error = 100000 + (int)r;
                    delete pl;
                    return RPE_ERR_CUDA;
                }
            }
            if (pln < p.w_planes) {   // weights: (Cin_s, Cout_pad, taps) fp16 with row pitch w_cstride, box (64, bn, 1)
                cuuint64_t dims[3] = {(cuuint64_t)sc.c_count, (cuuint64_t)d->cout_pad, (cuuint64_t)taps};
                cuuint64_t strides[2] = {(cuuint64_t)sc.w_cstride * 2, (cuuint64_t)sc.w_cstride * 2 * d->cout_pad};
                cuuint32_t box[3] = {kCvBK, (cuuint32_t)b_rows, 1};
                cuuint32_t es[3] = {1, 1, 1};
                CUresult r = g_cv_encode(&p.wmap[s][pln], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(wgt), dims, strides, box,
                                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    g_last_cuda_error = 100000 + (int)r;
                    delete pl;
                    return RPE_ERR_CUDA;
                }
                if (p.merge_n) {      // the same plane with a box of all bn rows
                    cuuint32_t boxx[3] = {kCvBK, (cuuint32_t)bn, 1};
                    r = g_cv_encode(&p.wmapx[s][pln], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(wgt), dims, strides, boxx, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) {
                        g_last_cuda_error = 100000 + (int)r;
                        delete pl;
                        return RPE_ERR_CUDA;
                    }
                }
            }
        }
    }
    p.N = d->N, p.OH = OH, p.OW = OW;
    p.tiles_min = ((orient == 0 ? OW : OH) + 7) / 8;
    p.tiles_maj = ((orient == 0 ? OH : OW) + 15) / 16;
    p.cout = d->cout, p.bn = bn, p.n_blocks = n_blocks;
    p.bias = d->bias, p.act = d->activation, p.scale = d->out_scale;
    p.acc_scale = d->acc_scale != 0.0f ? d->acc_scale : 1.0f;
    p.pre = d->pre, p.pre_ld = d->pre_ld, p.res = d->res, p.res_ld = d->res_ld;
    p.res_hi = reinterpret_cast<const rpe::plane_t *>(d->res_hi), p.res_lo = reinterpret_cast<const rpe::plane_t *>(d->res_lo);
    p.out_f32 = d->out_f32, p.f32_ld = d->f32_ld, p.f32_off = d->f32_offset;
    p.out_hi = reinterpret_cast<rpe::plane_t *>(d->out_hi), p.out_lo = reinterpret_cast<rpe::plane_t *>(d->out_lo);
    p.bf_ld = d->bf_ld, p.bf_off = d->bf_offset;
    p.mode = d->mode, p.aux = d->aux, p.aux_ld = d->aux_ld, p.aux2 = d->aux2, p.aux2_ld = d->aux2_ld;
    p.dbg = cv_env().dbg;
    if (p.mode != 0) {
        // GRU epilogues: channel groups of 4 never straddle the z|r boundary; state tensors must be 16-byte addressable
        bool ok;
        if (p.mode == 3)        // tap projection: partial sums only, 2 x 18 floats per pixel; weights [cout][18] fp32 in aux2
            ok = d->aux2 && d->out_f32 && !d->out_hi && !d->pre && !d->res && n_blocks == 1 && d->cout <= 256 && bn >= 32 &&
                 d->cout == d->cout_pad && (d->f32_ld % 2) == 0 && (d->f32_offset % 2) == 0 && d->f32_ld >= d->f32_offset + 2 * kCvProj;
        else
            ok = (p.mode == 1 || p.mode == 2) && d->activation == (p.mode == 1 ? 2 : 3) && d->aux && aligned16(d->aux) && (d->aux_ld % 4) == 0 && d->out_hi && d->out_lo &&
                 (d->cout % 8) == 0 && n_blocks == 1 &&
                 (p.mode == 1 ? (d->out_f32 != nullptr) : (d->aux2 != nullptr && aligned16(d->aux2) && (d->aux2_ld % 4) == 0));
        if (!ok) {
            delete pl;
            return RPE_ERR_INVALID_ARG;
        }
    }
    pl->flops = 2.0 * macs * (double)d->cout * (double)d->N * OH * OW * (double)(1 + (p.n_planes == 2) + (p.w_planes == 2));
    if (pl->pair) {
        const int units = ((px_tiles + 1) / 2) * p.n_blocks;
        int clusters = sm_count() / 2;
        if (clusters > units) clusters = units;
        pl->grid = 2 * (clusters < 1 ? 1 : clusters);
    } else {
        const int tiles = px_tiles * p.n_blocks;
        pl->grid = sm_count() < tiles ? sm_count() : tiles;
        if (pl->grid < 1) pl->grid = 1;
    }
    if (p.mode < 0 || p.mode > 3) {
        delete pl;
        return RPE_ERR_INVALID_ARG;
    }
    {   // 32-bit element offsets in the epilogue: no tensor row index times leading dimension may reach 2^32
        int max_ld = d->f32_ld > d->bf_ld ? d->f32_ld : d->bf_ld;
        const int lds[4] = {d->pre ? d->pre_ld : 0, (d->res || d->res_hi) ? d->res_ld : 0, d->aux ? d->aux_ld : 0, (d->aux2 && p.mode == 2) ? d->aux2_ld : 0};
        for (int k = 0; k < 4; ++k) max_ld = lds[k] > max_ld ? lds[k] : max_ld;
        if ((double)d->N * OH * OW * (double)max_ld >= 4294967296.0) {
            delete pl;
            return RPE_ERR_INVALID_ARG;
        }
    }
    // kernel kind: the GRU / projection modes have their own epilogues; mode 0 runs a lean epilogue when it only writes split
    // planes or only fp32 (the convolutions of the motion encoder and of the instance-norm encoder), else the generic one
    pl->kind = p.mode;
    if (p.mode == 0 && !d->pre && !d->res && !d->res_hi && d->activation <= 1 && !cv_env().generic) {
        if (d->out_hi && d->out_lo && !d->out_f32 && d->out_scale == 1.0f) pl->kind = kKPlanes;
        else if (d->out_f32 && !d->out_hi) pl->kind = kKF32;
    }
    p.side_prefetch = ((p.mode == 1 || p.mode == 2) && cv_env().side_prefetch && n_blocks == 1) ? 1 : 0;
    p.stat_part = d->stat_partials;
    if (d->stat_partials) {      // instance-norm partial sums ride on the fp32-only epilogue
        if (p.mode != 0 || d->pre || d->res || d->activation > 1 || !d->out_f32 || d->out_hi || d->out_scale != 1.0f || (d->cout % 16) ||
            !aligned16(d->stat_partials)) {
            delete pl;
            return RPE_ERR_INVALID_ARG;
        }
        pl->kind = kKF32Stats;
    }
    pl->tiles_per_image = p.tiles_min * p.tiles_maj;
    {
        cudaError_t e = cv_ensure_attr(pl->kind, p);
        if (e != cudaSuccess) {
            delete pl;
            return cuda_fail(e);
        }
    }
    *plan_out = pl;
    return RPE_OK;
}

// All-pairs correlation volume + pooled pyramid on the convolution kernel (kind 7): level0[b, q, t] = <f1[b, q, :], f2[b, t, :]> /
// sqrt(C) as a 1x1 "convolution" of the first feature map whose weights are the second feature map of the same image.
int rpe_corr_build_planes(const void *f1_hi, const void *f1_lo, const void *f2_hi, const void *f2_lo, float *pyramid, int B, int C, int h,
                          int w, int num_levels, int f1_wrap, int f1_sub, void *stream) {
    using namespace rpe;
    if (!f1_hi || !f1_lo || !f2_hi || !f2_lo || !pyramid) return RPE_ERR_INVALID_ARG;
    if (B <= 0 || C <= 0 || (C % 64) || h <= 0 || w <= 0 || num_levels < 1 || num_levels > 4) return RPE_ERR_INVALID_ARG;
    if ((h >> (num_levels - 1)) < 1 || (w >> (num_levels - 1)) < 1) return RPE_ERR_INVALID_ARG;
    if ((double)B * h * w >= 4294967296.0) return RPE_ERR_INVALID_ARG;
    const void *ptrs[4] = {f1_hi, f1_lo, f2_hi, f2_lo};
    for (int k = 0; k < 4; ++k)
        if (reinterpret_cast<uintptr_t>(ptrs[k]) & 15u) return RPE_ERR_ALIGNMENT;
    if (!aligned16(pyramid)) return RPE_ERR_ALIGNMENT;
    int rc = cv_load_encode();
    if (rc != RPE_OK) return rc;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.n_src = 1, p.n_planes = 2, p.w_planes = 2;
    p.cblocks[0] = C / kCvBK, p.cb_base[0] = 0, p.ksteps_last[0] = kCvBK / 16;
    p.N = B, p.OH = h, p.OW = w;
    p.orient = 0, p.kmin = 1, p.kmaj = 1, p.pad_min = 0, p.pad_maj = 0, p.stride = 1, p.kw = 1, p.reuse = 1;
    p.tiles_min = (w + 7) / 8, p.tiles_maj = (h + 15) / 16;
    const int per_img = p.tiles_min * p.tiles_maj;
    // CTA pairs own two adjacent query tiles that must read the SAME target block, i.e. lie in the same image
    const bool pair = !cv_env().no_pair && (per_img % 2 == 0) && sm_count() >= 2;
    const int b_rows = pair ? 128 : 256;
    p.a_plane_bytes = 16 * 8 * 128, p.a_stage_bytes = 2 * p.a_plane_bytes;
    p.b_plane_bytes = (uint32_t)b_rows * 128, p.b_stage_bytes = 2 * p.b_plane_bytes;
    p.resident_b = 0;
    p.b_group = 1;
    p.n_a_stages = 2;
    if (3 * (size_t)p.a_stage_bytes + 8 * (size_t)p.b_plane_bytes <= (size_t)kCvSmemData) p.n_a_stages = 3;
    // Resident query tile (A-B switch RPE_CORR_RESIDENT_A, off by default): all K blocks of the CTA's 128 queries (C / 64 stages of
    // 32 KB) stay in shared memory while it visits the target blocks of the image, so that per unit only the target tiles stream
    // (half the L2 -> SM operand traffic).  Measured on B200 (64 samples, profiles/r2_corr_probe.txt): 4.59 ms against 2.95 ms
    // with both operands streaming -- the 128 KB tile leaves room for only 4 target-tile ring entries (two K blocks of
    // look-ahead), and the exposed TMA latency costs more than the saved bandwidth.
    p.resident_a = 0;
    if (pair && p.cblocks[0] <= kCvMaxAStages && cv_env().corr_resident_a &&
        (size_t)p.cblocks[0] * p.a_stage_bytes + 4 * (size_t)p.b_plane_bytes <= (size_t)kCvSmemData) {
        p.resident_a = 1;
        p.n_a_stages = p.cblocks[0];
    }
    p.n_b_stages = (int)((kCvSmemData - (size_t)p.n_a_stages * p.a_stage_bytes) / p.b_plane_bytes);
    if (p.n_b_stages > kCvMaxBStages) p.n_b_stages = kCvMaxBStages;
    p.corr_h = h, p.corr_w = w, p.corr_nbx = (w + 15) / 16, p.corr_levels = num_levels;
    if (f1_wrap <= 0) f1_wrap = B, f1_sub = 0;                  // identity: sample s correlates f1 image s with f2 image s
    if (f1_sub < 0 || f1_sub > f1_wrap) return RPE_ERR_INVALID_ARG;
    p.corr_a_wrap = f1_wrap, p.corr_a_sub = f1_sub;
    p.bn = 256, p.n_blocks = p.corr_nbx * ((h + 15) / 16), p.cout = p.bn * p.n_blocks;
    p.scale = 1.0f / sqrtf((float)C);
    p.acc_scale = 1.0f;
    {   // level bases (the layout of rpe_corr_level_offset) and which of them take vector stores
        size_t off = 0;
        for (int l = 0; l < num_levels; ++l) {
            p.corr_lvl[l] = pyramid + off;
            const size_t n = (size_t)B * h * w * (size_t)(h >> l) * (size_t)(w >> l);
            off += (n + 63) & ~(size_t)63;
        }
        p.corr_vec = 0;
        if ((w % 4) == 0 && (((size_t)h * w) % 4) == 0) p.corr_vec |= 1;
        if (num_levels > 1 && ((w >> 1) % 2) == 0 && (((size_t)(h >> 1) * (w >> 1)) % 2) == 0 && (w % 2) == 0) p.corr_vec |= 2;
        if (!cv_env().corr_plain_stores) p.corr_vec |= 16;         // streaming (evict-first) stores for the pyramid
    }
    for (int pln = 0; pln < 2; ++pln) {
        const cuuint64_t pix = (cuuint64_t)C * 2, row = pix * w;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
        cuuint64_t strides[3] = {pix, row, row * h};
        cuuint32_t es[4] = {1, 1, 1, 1};
        cuuint32_t abox[4] = {kCvBK, 8, 16, 1};
        cuuint32_t bbox[4] = {kCvBK, 16, (cuuint32_t)(b_rows / 16), 1};
        CUresult r = g_cv_encode(&p.amap[0][pln], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptrs[pln]), dims, strides, abox, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = g_cv_encode(&p.wmap[0][pln], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptrs[2 + pln]), dims, strides, bbox, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            g_last_cuda_error = 100000 + (int)r;
            return RPE_ERR_CUDA;
        }
    }
    RPE_CUDA_TRY(cv_ensure_attr(kKCorr, p));
    const int px_tiles = B * per_img;
    int grid;
    if (pair) {
        const long long units = (long long)(px_tiles / 2) * p.n_blocks;
        long long clusters = sm_count() / 2;
        if (clusters > units) clusters = units;
        grid = 2 * (int)(clusters < 1 ? 1 : clusters);
    } else {
        const long long tiles = (long long)px_tiles * p.n_blocks;
        grid = (int)(sm_count() < tiles ? sm_count() : tiles);
    }
    cv_dispatch(kKCorr, p, grid, pair, (cudaStream_t)stream, false);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_conv_plan_run(void *plan, void *stream) {
    using namespace rpe;
    if (!plan) return RPE_ERR_INVALID_ARG;
    ConvPlan *pl = reinterpret_cast<ConvPlan *>(plan);
    cv_dispatch(pl->kind, pl->p, pl->grid, pl->pair, (cudaStream_t)stream, false);
    RPE_LAUNCH_CHECK();
    return RPE_OK;
}

int rpe_conv_plan_tiles_per_image(void *plan) {
    if (!plan) return 0;
    return reinterpret_cast<rpe::ConvPlan *>(plan)->tiles_per_image;
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
