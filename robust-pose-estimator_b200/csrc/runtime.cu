// Library-level entry points: version, status strings, device queries.
#include "common.cuh"

namespace rpe {
int g_last_cuda_error = 0;

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached = n;
    }
    return cached;
}
}  // namespace rpe

extern "C" {

int rpe_version(void) { return 100; }

const char *rpe_status_string(int status) {
    switch (status) {
        case RPE_OK: return "ok";
        case RPE_ERR_INVALID_ARG: return "invalid argument";
        case RPE_ERR_ALIGNMENT: return "pointer or stride misaligned";
        case RPE_ERR_WORKSPACE: return "workspace too small";
        case RPE_ERR_CUDA: return "CUDA call failed";
        case RPE_ERR_UNSUPPORTED_DEVICE: return "device is not sm_100";
        default: return "unknown status";
    }
}

int rpe_last_cuda_error(void) { return rpe::g_last_cuda_error; }

int rpe_device_sm_count(void) { return rpe::sm_count(); }

}  // extern "C"
