// Shared helpers for the rpe_b200 kernels (sm_100a only).
#pragma once
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

int sm_count();

}  // namespace rpe
