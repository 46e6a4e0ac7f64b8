// Shared helpers for the rpe_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rpe_b200.h"

namespace rpe {

extern int g_last_cuda_error;
extern long long g_launch_count;   // kernels launched by this library (bench.py reports it as gpu_launches)

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return RPE_ERR_CUDA;
}

#define RPE_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return ::rpe::cuda_fail(_e);  \
    } while (0)

#define RPE_LAUNCH_CHECK()                                   \
    do {                                                     \
        cudaError_t _e = cudaGetLastError();                 \
        if (_e != cudaSuccess) return ::rpe::cuda_fail(_e);  \
        ++::rpe::g_launch_count;                             \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- split planes -------------------------------------------------------------------------------------------------------
// Every tensor-core operand is the sum of two fp16 planes: hi = fp16(v), lo = fp16(v - hi): 22 significant bits while the lo
// plane is a normal number (|v| >= 2^-3), an absolute error of 2^-25 below that (the lo plane turns subnormal).  hi*hi + lo*hi +
// hi*lo on the kind::f16 MMAs with fp32 accumulation then tracks fp32 arithmetic to ~3e-7 relative, PROVIDED the operands are
// not small: activations of this network are O(0.1 .. 10); convolution weights (O(0.01)) are therefore packed pre-multiplied
// by a power of two and the accumulators are multiplied back in the epilogue (rpe_conv_desc.acc_scale) -- exact.
// Why not two bf16 planes (round 1): 16 bits per operand miss the pose gate on real texture (tools/precision_study.py, the
// tartan_air fixture pair).  Mixed formats (fp16 hi x bf16 lo) are not an option: tcgen05.mma kind::f16 takes ONE format for A
// and B per instruction (an f16 x bf16 descriptor raises an illegal-instruction fault on sm_100a; tried).
// Conversions saturate (cvt.satfinite) instead of producing infinities beyond +-65504.
#if defined(__CUDACC__)
typedef __half plane_t;
typedef __half2 plane2_t;
__device__ __forceinline__ plane_t to_plane(float v) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return __ushort_as_half(r);
}
__device__ __forceinline__ float plane_to_float(plane_t h) { return __half2float(h); }
// (a, b) -> packed pair with a in the low half
__device__ __forceinline__ plane2_t to_plane2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return *reinterpret_cast<plane2_t *>(&r);
}
// lo plane: the remainder, same format
__device__ __forceinline__ plane_t to_plane_lo(float v) { return to_plane(v); }
__device__ __forceinline__ plane2_t to_plane2_lo(float a, float b) { return to_plane2(a, b); }
#endif

int sm_count();

}  // namespace rpe
