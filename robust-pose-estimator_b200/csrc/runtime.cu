// Library-level entry points: version, status strings, device queries.
#include <math.h>
#include "common.cuh"

namespace rpe {
int g_last_cuda_error = 0;
long long g_launch_count = 0;

int sm_count() {
    static int cached[64] = {};                      // per device
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
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

long long rpe_launch_count(void) { return rpe::g_launch_count; }

int rpe_l2_fetch_granularity(int bytes) {
    if (bytes > 0) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
        if (e != cudaSuccess) {
            rpe::g_last_cuda_error = (int)e;
            (void)cudaGetLastError();
            return RPE_ERR_CUDA;
        }
    }
    size_t v = 0;
    if (cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) != cudaSuccess) return RPE_ERR_CUDA;
    return (int)v;
}

// ---- host-side trajectory composition (fp32, same operation order as the reference's lietorch calls) ----
namespace {
inline void qrot(const float *q, const float *p, float *o) {
    const float uvx = 2.0f * (q[1] * p[2] - q[2] * p[1]);
    const float uvy = 2.0f * (q[2] * p[0] - q[0] * p[2]);
    const float uvz = 2.0f * (q[0] * p[1] - q[1] * p[0]);
    o[0] = p[0] + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
    o[1] = p[1] + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
    o[2] = p[2] + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}
}  // namespace

int rpe_compose_trajectory_host(const float *rel_host, const float *log_host, int n, const float *init_pose_host,
                                float inv_scale, float *abs_out_host, unsigned char *failed_out_host) {
    if (!init_pose_host || !abs_out_host || n < 0 || (n > 0 && (!rel_host || !log_host))) return RPE_ERR_INVALID_ARG;    // n = 0: a one-frame sequence
    float last[7];
    for (int k = 0; k < 7; ++k) last[k] = abs_out_host[k] = init_pose_host[k];
    for (int i = 0; i < n; ++i) {
        const float *r = rel_host + 7 * i, *lg = log_host + 6 * i;
        bool bad = false;
        for (int k = 0; k < 7; ++k) bad = bad || (r[k] != r[k]);
        for (int k = 0; k < 6; ++k) bad = bad || (fabsf(lg[k]) > 1.0e-1f);
        if (failed_out_host) failed_out_host[i] = bad ? 1 : 0;
        float t[3] = {0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 1.f};
        if (!bad) {
            for (int k = 0; k < 3; ++k) t[k] = r[k] * inv_scale;        // rel.scale(1 / scale)
            for (int k = 0; k < 4; ++k) q[k] = r[3 + k];
        }
        // inverse: q^-1, -(q^-1 t)
        const float qi[4] = {-q[0], -q[1], -q[2], q[3]};
        float ti[3];
        qrot(qi, t, ti);
        ti[0] = -ti[0], ti[1] = -ti[1], ti[2] = -ti[2];
        // last <- last * inv
        float rt[3];
        qrot(last + 3, ti, rt);
        const float ax = last[3], ay = last[4], az = last[5], aw = last[6];
        const float nq[4] = {aw * qi[0] + ax * qi[3] + ay * qi[2] - az * qi[1], aw * qi[1] - ax * qi[2] + ay * qi[3] + az * qi[0],
                             aw * qi[2] + ax * qi[1] - ay * qi[0] + az * qi[3], aw * qi[3] - ax * qi[0] - ay * qi[1] - az * qi[2]};
        for (int k = 0; k < 3; ++k) last[k] = last[k] + rt[k];
        for (int k = 0; k < 4; ++k) last[3 + k] = nq[k];
        for (int k = 0; k < 7; ++k) abs_out_host[7 * (i + 1) + k] = last[k];
    }
    return RPE_OK;
}

}  // extern "C"
