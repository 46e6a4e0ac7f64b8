// Stages 3 + 4: fused per-pixel residual / Jacobian reduction over SE(3) and the on-device solver.
//
// One persistent cooperative kernel per batch of pairs.  The grid is split into groups of CTAs;
// each group owns one pair at a time.  Per objective evaluation every CTA of the group
//   1. streams its pixel slices (42 B/px: flow 8 + pcl1 12 + pcl2w 12 + w1 4 + w2 4 + 2 mask bytes), evaluates the 2D
//      reprojection and 3D point-to-point residuals, the Lie-algebra Jacobians and the confidence weights in fp64, and
//      reduces f, grad (6) [and the 21 GN Hessian entries] with warp shuffles + a block-level tree,
//   2. publishes its partial sums, passes a group-wide barrier, and
//   3. re-sums ALL partials of the pair in a fixed order, so every CTA holds bit-identical totals
//      and advances an identical copy of the 6-dim solver state -- no broadcast, one barrier per
//      evaluation, no host round trip (the reference syncs on float(loss) every iteration).
// The pixels of a pair are dealt to kVirt = 128 VIRTUAL blocks by index alone and a CTA evaluates whole virtual blocks, so
// the summation tree -- and with it every bit of the result -- does not depend on how many CTAs or concurrent groups a
// launch uses (a pair solved alone, in a batch of 32 or on another rank of a sharded run gives the same pose).
// Measured (B200): the kernel is bound by the fp64 pipe and by the serial part of every evaluation (group barrier, re-summation,
// solver step), not by memory -- capping the concurrent groups so that their inputs stay L2-resident (4 groups of 64 CTAs) is
// slower than many small groups that stream from HBM.
//
// Reference semantics: /root/reference/core/pose/pose_head.py:12-79,
// core/geometry/pinhole_transforms.py:28-30,90-99, torch.optim.LBFGS.step (lr=1, no line search,
// tolerance_grad 1e-7, tolerance_change 1e-9, history 100, max_eval = 5/4 max_iter),
// clip_grad_norm_(y, 10), lietorch left retraction X <- Exp(t d) X.
#include <cooperative_groups.h>
#include <math.h>
#include "common.cuh"

namespace rpe {

constexpr int kPoseThreads = 256;
constexpr int kMaxHist = 100;      // torch.optim.LBFGS history_size default
constexpr int kAccGrad = 8;        // e2, e3, g[6]
constexpr int kAccHess = 8 + 21;   // + upper triangle of J^T W J
constexpr int kVirt = 128;         // virtual blocks per pair (fixed reduction tree)
constexpr int kMaxGroups = 64;

struct PoseParams {
    rpe_pose_problem p;
    int mode, max_iter, with_hessian;
    int blocks_per_group, n_groups;
    double *out;
    float *pose_f32, *log_f32;
    double *trace;
    int trace_cap;
    double *partials;          // [n_groups][2][kVirt][kAccHess]
    unsigned int *counters;    // [n_groups] (zeroed by the host before launch), 128 B apart
};

// ------------------------------------------------------------------------------------------------
// SE(3) arithmetic in fp64 (lietorch semantics)
// ------------------------------------------------------------------------------------------------
struct Pose {
    double t[3];
    double q[4];   // x y z w
};

__device__ __forceinline__ void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ void quat_rotate(const double *q, const double *p, double *o) {
    double uv[3], uuv[3];
    cross3(q, p, uv);
    uv[0] *= 2.0, uv[1] *= 2.0, uv[2] *= 2.0;
    cross3(q, uv, uuv);
    o[0] = p[0] + q[3] * uv[0] + uuv[0];
    o[1] = p[1] + q[3] * uv[1] + uuv[1];
    o[2] = p[2] + q[3] * uv[2] + uuv[2];
}

__device__ void se3_exp(const double *xi, Pose &o) {
    const double *tau = xi, *phi = xi + 3;
    const double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    const double th = sqrt(th2);
    double imag, real, c1, c2;
    if (th < 1e-6) {
        imag = 0.5 - th2 / 48.0 + th2 * th2 / 3840.0;
        real = 1.0 - th2 / 8.0 + th2 * th2 / 384.0;
        c1 = 0.5 - th2 / 24.0;
        c2 = 1.0 / 6.0 - th2 / 120.0;
    } else {
        imag = sin(0.5 * th) / th;
        real = cos(0.5 * th);
        c1 = (1.0 - cos(th)) / th2;
        c2 = (th - sin(th)) / (th2 * th);
    }
    double pxt[3], ppxt[3];
    cross3(phi, tau, pxt);
    cross3(phi, pxt, ppxt);
    for (int k = 0; k < 3; ++k) {
        o.t[k] = tau[k] + c1 * pxt[k] + c2 * ppxt[k];
        o.q[k] = imag * phi[k];
    }
    o.q[3] = real;
}

__device__ void se3_log(const Pose &X, double *xi) {
    const double n2 = X.q[0] * X.q[0] + X.q[1] * X.q[1] + X.q[2] * X.q[2];
    const double w = X.q[3];
    double s;
    if (n2 < 1e-12) {
        s = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
    } else {
        const double n = sqrt(n2);
        s = 2.0 * atan(n / w) / n;
    }
    double phi[3] = {s * X.q[0], s * X.q[1], s * X.q[2]};
    const double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    const double th = sqrt(th2);
    double c2;
    if (th < 1e-6) {
        c2 = 1.0 / 12.0;
    } else {
        c2 = (1.0 - th * cos(0.5 * th) / (2.0 * sin(0.5 * th))) / th2;
    }
    double pxt[3], ppxt[3];
    cross3(phi, X.t, pxt);
    cross3(phi, pxt, ppxt);
    for (int k = 0; k < 3; ++k) {
        xi[k] = X.t[k] - 0.5 * pxt[k] + c2 * ppxt[k];
        xi[3 + k] = phi[k];
    }
}

// X <- Exp(a) * X   (LieGroupParameter.add_ / retr)
__device__ void se3_retract(Pose &X, const double *a) {
    Pose E;
    se3_exp(a, E);
    double rt[3];
    quat_rotate(E.q, X.t, rt);
    const double ax = E.q[0], ay = E.q[1], az = E.q[2], aw = E.q[3];
    const double bx = X.q[0], by = X.q[1], bz = X.q[2], bw = X.q[3];
    X.q[0] = aw * bx + ax * bw + ay * bz - az * by;
    X.q[1] = aw * by - ax * bz + ay * bw + az * bx;
    X.q[2] = aw * bz + ax * by - ay * bx + az * bw;
    X.q[3] = aw * bw - ax * bx - ay * by - az * bz;
    for (int k = 0; k < 3; ++k) X.t[k] = E.t[k] + rt[k];
}

__device__ __forceinline__ void pose_to_rt(const Pose &X, double *R) {
    // rotation via the same uv/uuv form applied to the basis vectors -> R p == quat_rotate(q, p)
    const double x = X.q[0], y = X.q[1], z = X.q[2], w = X.q[3];
    R[0] = 1.0 - 2.0 * (y * y + z * z);
    R[1] = 2.0 * (x * y - z * w);
    R[2] = 2.0 * (x * z + y * w);
    R[3] = 2.0 * (x * y + z * w);
    R[4] = 1.0 - 2.0 * (x * x + z * z);
    R[5] = 2.0 * (y * z - x * w);
    R[6] = 2.0 * (x * z - y * w);
    R[7] = 2.0 * (y * z + x * w);
    R[8] = 1.0 - 2.0 * (x * x + y * y);
}

// ------------------------------------------------------------------------------------------------
// Solver state, one identical copy per CTA (shared memory, advanced by thread 0)
// ------------------------------------------------------------------------------------------------
struct SolverState {
    Pose X;
    double R[9];
    double g[6], prev_g[6], d[6];
    double loss, prev_loss, t, H_diag;
    double L2, L3, g_raw[6];
    double hess[21];
    double old_dirs[kMaxHist][6], old_stps[kMaxHist][6], ro[kMaxHist];
    int num_old, n_iter, evals, status;
    int done;
};

struct PixelConsts {
    double K[9];
    double s2, s3;   // lw[1] / N / (H W),  lw[0] / N
    int W, H, N;
};

// Accumulate one pixel.  acc: [0]=sum e2, [1]=sum e3, [2..7]=grad, [8..28]=Hessian upper triangle.
template <bool kHess>
__device__ __forceinline__ void accumulate_pixel(const PixelConsts &c, const double *R, const double *tr, int idx, float fx,
                                                 float fy, float p1x, float p1y, float p1z, float p2x, float p2y,
                                                 float p2z, float w1, float w2, bool m1, bool m2, double *acc) {
    const double px = (double)p1x, py = (double)p1y, pz = (double)p1z;
    const double X = R[0] * px + R[1] * py + R[2] * pz + tr[0];
    const double Y = R[3] * px + R[4] * py + R[5] * pz + tr[1];
    const double Z = R[6] * px + R[7] * py + R[8] * pz + tr[2];
    // ---- 3D point-to-point residual (pose_head.py:43-51)
    const double r3x = X - (double)p2x, r3y = Y - (double)p2y, r3z = Z - (double)p2z;
    const bool v3 = m1 && m2;
    const double e3 = (r3x * r3x + r3y * r3y + r3z * r3z) * (double)w2;
    const double c3 = v3 ? 2.0 * (double)w2 * c.s3 : 0.0;
    acc[1] += v3 ? e3 : 0.0;
    // ---- 2D reprojection residual (pose_head.py:18-29; project: pinhole_transforms.py:90-99)
    const double qx = c.K[0] * X + c.K[1] * Y + c.K[2] * Z;
    const double qy = c.K[3] * X + c.K[4] * Y + c.K[5] * Z;
    const double qz = c.K[6] * X + c.K[7] * Y + c.K[8] * Z;
    const double den = fmax(qz, 1e-12);
    const double pass = (qz >= 1e-12) ? 1.0 : 0.0;
    const double inv = 1.0 / den;                      // one fp64 division per pixel; q * inv differs from q / den by <= 1 ulp
    const double pix = qx * inv, piy = qy * inv;
    const int row = idx / c.W, col = idx - row * c.W;
    const double tx = ((double)col + 0.5) + (double)fx;
    const double ty = ((double)row + 0.5) + (double)fy;
    const double r2x = tx - pix, r2y = ty - piy;
    const double e2 = (r2x * r2x + r2y * r2y) * (double)w1;
    const bool inside = (tx > 0.0) && (ty > 0.0) && (tx < (double)c.W) && (ty < (double)c.H);
    const bool bad = isinf(e2) || isnan(e2) || !inside || !m1;
    acc[0] += bad ? 0.0 : e2;
    const double c2 = bad ? 0.0 : 2.0 * (double)w1 * c.s2;
    // ---- gradient wrt the left perturbation: J = [I | -[p']x]
    const double gpx = c2 * (-r2x), gpy = c2 * (-r2y);
    const double gqx = gpx * inv, gqy = gpy * inv, gqz = -(gpx * qx + gpy * qy) * inv * inv * pass;
    const double ax = c3 * r3x + (c.K[0] * gqx + c.K[3] * gqy + c.K[6] * gqz);
    const double ay = c3 * r3y + (c.K[1] * gqx + c.K[4] * gqy + c.K[7] * gqz);
    const double az = c3 * r3z + (c.K[2] * gqx + c.K[5] * gqy + c.K[8] * gqz);
    acc[2] += ax;
    acc[3] += ay;
    acc[4] += az;
    acc[5] += Y * az - Z * ay;
    acc[6] += Z * ax - X * az;
    acc[7] += X * ay - Y * ax;
    if (kHess) {
        // J rows (3x6): [1 0 0 0 Z -Y; 0 1 0 -Z 0 X; 0 0 1 Y -X 0]
        const double J[3][6] = {{1.0, 0.0, 0.0, 0.0, Z, -Y}, {0.0, 1.0, 0.0, -Z, 0.0, X}, {0.0, 0.0, 1.0, Y, -X, 0.0}};
        // dpi/dp' = dpi/dq K  (2x3)
        const double dq0[3] = {inv, 0.0, -qx * inv * inv * pass};
        const double dq1[3] = {0.0, inv, -qy * inv * inv * pass};
        double dp[2][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dp[0][k] = dq0[0] * c.K[k] + dq0[1] * c.K[3 + k] + dq0[2] * c.K[6 + k];
            dp[1][k] = dq1[0] * c.K[k] + dq1[1] * c.K[3 + k] + dq1[2] * c.K[6 + k];
        }
        double J2[2][6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            J2[0][a] = dp[0][0] * J[0][a] + dp[0][1] * J[1][a] + dp[0][2] * J[2][a];
            J2[1][a] = dp[1][0] * J[0][a] + dp[1][1] * J[1][a] + dp[1][2] * J[2][a];
        }
        int k = 8;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) {
                acc[k] += c3 * (J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b]) +
                          c2 * (J2[0][a] * J2[0][b] + J2[1][a] * J2[1][b]);
                ++k;
            }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Group-wide barrier on a monotonically increasing counter (all CTAs co-resident: cooperative launch).
__device__ __forceinline__ void group_barrier(unsigned int *counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

template <bool kHess>
__device__ void evaluate_group(const PoseParams &P, const PixelConsts &c, int pair, int blk, SolverState &S, double *s_red,
                               double *partial_base) {
    constexpr int NA = kHess ? kAccHess : kAccGrad;
    const int N = c.N;
    const size_t off = (size_t)pair * N;
    const float *flow = P.p.flow + off * 2;
    const float *p1 = P.p.pcl1 + off * 3;
    const float *p2 = P.p.pcl2 + off * 3;
    const float *w1 = P.p.w1 ? P.p.w1 + off : nullptr;
    const float *w2 = P.p.w2 ? P.p.w2 + off : nullptr;
    const uint8_t *m1 = P.p.m1 + off;
    const uint8_t *m2 = P.p.m2 + off;
    double R[9], tr[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = S.R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) tr[k] = S.X.t[k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = kVirt * kPoseThreads * 4;
    for (int vb = blk; vb < kVirt; vb += P.blocks_per_group) {
        double acc[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) acc[k] = 0.0;
        for (int i = (vb * kPoseThreads + threadIdx.x) * 4; i < N; i += stride) {
            const float4 fx = *reinterpret_cast<const float4 *>(flow + i);
            const float4 fy = *reinterpret_cast<const float4 *>(flow + N + i);
            const float4 ax = *reinterpret_cast<const float4 *>(p1 + i);
            const float4 ay = *reinterpret_cast<const float4 *>(p1 + N + i);
            const float4 az = *reinterpret_cast<const float4 *>(p1 + 2 * N + i);
            const float4 bx = *reinterpret_cast<const float4 *>(p2 + i);
            const float4 by = *reinterpret_cast<const float4 *>(p2 + N + i);
            const float4 bz = *reinterpret_cast<const float4 *>(p2 + 2 * N + i);
            const float4 c1 = w1 ? *reinterpret_cast<const float4 *>(w1 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 c2 = w2 ? *reinterpret_cast<const float4 *>(w2 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
            const uchar4 ma = *reinterpret_cast<const uchar4 *>(m1 + i);
            const uchar4 mb = *reinterpret_cast<const uchar4 *>(m2 + i);
            accumulate_pixel<kHess>(c, R, tr, i + 0, fx.x, fy.x, ax.x, ay.x, az.x, bx.x, by.x, bz.x, c1.x, c2.x, ma.x != 0, mb.x != 0, acc);
            accumulate_pixel<kHess>(c, R, tr, i + 1, fx.y, fy.y, ax.y, ay.y, az.y, bx.y, by.y, bz.y, c1.y, c2.y, ma.y != 0, mb.y != 0, acc);
            accumulate_pixel<kHess>(c, R, tr, i + 2, fx.z, fy.z, ax.z, ay.z, az.z, bx.z, by.z, bz.z, c1.z, c2.z, ma.z != 0, mb.z != 0, acc);
            accumulate_pixel<kHess>(c, R, tr, i + 3, fx.w, fy.w, ax.w, ay.w, az.w, bx.w, by.w, bz.w, c1.w, c2.w, ma.w != 0, mb.w != 0, acc);
        }
        // ---- block reduction of this virtual block: warp shuffles, then a tree over the 8 warp partials in shared memory
        __syncthreads();                                   // s_red of the previous virtual block has been consumed
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            const double v = warp_sum(acc[k]);
            if (lane == 0) s_red[warp * kAccHess + k] = v;
        }
        __syncthreads();
        if (threadIdx.x < NA) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kPoseThreads / 32; ++w) v += s_red[w * kAccHess + threadIdx.x];
            partial_base[(size_t)vb * kAccHess + threadIdx.x] = v;
        }
    }
}

// After the barrier: every CTA sums all partials of its group in the same order.
template <bool kHess>
__device__ void gather_totals(const PoseParams &P, const PixelConsts &c, const double *partials_eval, double lw0, double lw1,
                              SolverState &S) {
    constexpr int NA = kHess ? kAccHess : kAccGrad;
    __shared__ double s_tot[kAccHess];
    __shared__ double s_q[kAccHess][4];
    if (threadIdx.x < 4 * NA) {                            // four interleaved quarter sums per accumulator, then ((q0 + q1) + q2) + q3
        const int a = threadIdx.x >> 2, part = threadIdx.x & 3;
        double v = 0.0;
        for (int b = part; b < kVirt; b += 4) v += __ldcg(partials_eval + (size_t)b * kAccHess + a);
        s_q[a][part] = v;
    }
    __syncthreads();
    if (threadIdx.x < NA) s_tot[threadIdx.x] = ((s_q[threadIdx.x][0] + s_q[threadIdx.x][1]) + s_q[threadIdx.x][2]) + s_q[threadIdx.x][3];
    __syncthreads();
    if (threadIdx.x == 0) {
        const double N = (double)c.N;
        S.L2 = s_tot[0] / N / ((double)c.H * (double)c.W);
        S.L3 = s_tot[1] / N;
        S.loss = lw1 * S.L2 + lw0 * S.L3;
        for (int k = 0; k < 6; ++k) S.g_raw[k] = s_tot[2 + k];
        if (kHess)
            for (int k = 0; k < 21; ++k) S.hess[k] = s_tot[8 + k];
    }
    __syncthreads();
}

__device__ void write_trace(const PoseParams &P, int pair, const SolverState &S) {
    if (!P.trace || S.evals > P.trace_cap) return;
    double *t = P.trace + ((size_t)pair * P.trace_cap + (S.evals - 1)) * 16;
    for (int k = 0; k < 3; ++k) t[k] = S.X.t[k];
    for (int k = 0; k < 4; ++k) t[3 + k] = S.X.q[k];
    for (int k = 0; k < 6; ++k) t[7 + k] = S.g_raw[k];
    t[13] = S.loss, t[14] = S.L2, t[15] = S.L3;
}

// clip_grad_norm_(y, 10): g *= min(1, 10 / (||g||_2 + 1e-6))
__device__ void clip_gradient(SolverState &S) {
    double n2 = 0.0;
    for (int k = 0; k < 6; ++k) n2 += S.g_raw[k] * S.g_raw[k];
    double coef = 10.0 / (sqrt(n2) + 1e-6);
    if (coef > 1.0) coef = 1.0;
    for (int k = 0; k < 6; ++k) S.g[k] = S.g_raw[k] * coef;
}

__device__ __forceinline__ double dot6(const double *a, const double *b) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += a[k] * b[k];
    return s;
}

// tensor.abs().max() of torch: a NaN component makes the result NaN (fmax alone would drop it), so that every stopping test of
// torch.optim.LBFGS fails on a non-finite gradient exactly like the reference's and the NaN pose reaches the tracker's guard
__device__ __forceinline__ double absmax6(const double *a) {
    double m = 0.0;
    bool nan = false;
    for (int k = 0; k < 6; ++k) {
        const double v = fabs(a[k]);
        nan = nan || (v != v);
        m = fmax(m, v);
    }
    return nan ? __longlong_as_double(0x7ff8000000000000LL) : m;
}

// One pass of the body of torch.optim.LBFGS.step's while loop up to (and including) the parameter
// update.  Returns true when the loop must stop BEFORE another evaluation is needed.
__device__ bool lbfgs_direction_and_step(SolverState &S, int max_iter, bool &need_eval) {
    need_eval = false;
    S.n_iter += 1;
    if (S.n_iter == 1) {
        for (int k = 0; k < 6; ++k) S.d[k] = -S.g[k];
        S.num_old = 0;
        S.H_diag = 1.0;
    } else {
        double y[6], s[6];
        for (int k = 0; k < 6; ++k) {
            y[k] = S.g[k] - S.prev_g[k];
            s[k] = S.d[k] * S.t;
        }
        const double ys = dot6(y, s);
        if (ys > 1e-10) {
            if (S.num_old == kMaxHist) {   // shift history (limited memory)
                for (int i = 1; i < kMaxHist; ++i) {
                    for (int k = 0; k < 6; ++k) {
                        S.old_dirs[i - 1][k] = S.old_dirs[i][k];
                        S.old_stps[i - 1][k] = S.old_stps[i][k];
                    }
                    S.ro[i - 1] = S.ro[i];
                }
                S.num_old -= 1;
            }
            for (int k = 0; k < 6; ++k) {
                S.old_dirs[S.num_old][k] = y[k];
                S.old_stps[S.num_old][k] = s[k];
            }
            S.ro[S.num_old] = 1.0 / ys;
            S.num_old += 1;
            S.H_diag = ys / dot6(y, y);
        }
        double al[kMaxHist];
        double q[6];
        for (int k = 0; k < 6; ++k) q[k] = -S.g[k];
        for (int i = S.num_old - 1; i >= 0; --i) {
            al[i] = dot6(S.old_stps[i], q) * S.ro[i];
            for (int k = 0; k < 6; ++k) q[k] += S.old_dirs[i][k] * (-al[i]);
        }
        for (int k = 0; k < 6; ++k) S.d[k] = q[k] * S.H_diag;
        for (int i = 0; i < S.num_old; ++i) {
            const double be = dot6(S.old_dirs[i], S.d) * S.ro[i];
            for (int k = 0; k < 6; ++k) S.d[k] += S.old_stps[i][k] * (al[i] - be);
        }
    }
    for (int k = 0; k < 6; ++k) S.prev_g[k] = S.g[k];
    S.prev_loss = S.loss;
    if (S.n_iter == 1) {
        double l1 = 0.0;
        for (int k = 0; k < 6; ++k) l1 += fabs(S.g[k]);
        S.t = fmin(1.0, 1.0 / l1);
    } else {
        S.t = 1.0;
    }
    const double gtd = dot6(S.g, S.d);
    if (gtd > -1e-9) return true;
    double step[6];
    for (int k = 0; k < 6; ++k) step[k] = S.t * S.d[k];
    se3_retract(S.X, step);
    pose_to_rt(S.X, S.R);
    if (S.n_iter != max_iter) {
        need_eval = true;
        return false;
    }
    return true;   // n_iter == max_iter: no re-evaluation, loop ends
}

// Checks that follow the re-evaluation inside the while loop.
__device__ bool lbfgs_post_eval_stop(const SolverState &S, int max_iter) {
    const int max_eval = max_iter * 5 / 4;
    if (S.n_iter == max_iter) return true;
    if (S.evals >= max_eval) return true;
    if (absmax6(S.g) <= 1e-7) return true;
    double dt[6];
    for (int k = 0; k < 6; ++k) dt[k] = S.d[k] * S.t;
    if (absmax6(dt) <= 1e-9) return true;
    if (fabs(S.loss - S.prev_loss) < 1e-9) return true;
    return false;
}

// 6x6 Cholesky solve H x = -g (upper triangle in S.hess); returns false if not positive definite.
__device__ bool gn_step(const SolverState &S, double *step) {
    double A[6][6];
    int k = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) {
            A[a][b] = S.hess[k];
            A[b][a] = S.hess[k];
            ++k;
        }
    double L[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = A[i][j];
            for (int m = 0; m < j; ++m) s -= L[i][m] * L[j][m];
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[i][i] = sqrt(s);
            } else {
                L[i][j] = s / L[j][j];
            }
        }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = -S.g_raw[i];
        for (int m = 0; m < i; ++m) s -= L[i][m] * y[m];
        y[i] = s / L[i][i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int m = i + 1; m < 6; ++m) s -= L[m][i] * step[m];
        step[i] = s / L[i][i];
    }
    return true;
}

__device__ void write_result(const PoseParams &P, int pair, const SolverState &S) {
    double *o = P.out + (size_t)pair * RPE_POSE_OUT_STRIDE;
    double lg[6];
    se3_log(S.X, lg);
    for (int k = 0; k < 3; ++k) o[k] = S.X.t[k];
    for (int k = 0; k < 4; ++k) o[3 + k] = S.X.q[k];
    for (int k = 0; k < 6; ++k) o[7 + k] = lg[k];
    o[13] = S.loss, o[14] = S.L2, o[15] = S.L3;
    o[16] = (double)S.evals, o[17] = (double)S.n_iter, o[18] = (double)S.status;
    for (int k = 0; k < 6; ++k) o[19 + k] = S.g_raw[k];
    for (int k = 0; k < 21; ++k) o[25 + k] = S.hess[k];
    if (P.pose_f32) {
        for (int k = 0; k < 3; ++k) P.pose_f32[pair * 7 + k] = (float)S.X.t[k];
        for (int k = 0; k < 4; ++k) P.pose_f32[pair * 7 + 3 + k] = (float)S.X.q[k];
    }
    if (P.log_f32)
        for (int k = 0; k < 6; ++k) P.log_f32[pair * 6 + k] = (float)lg[k];
}

template <bool kHess>
__global__ void __launch_bounds__(kPoseThreads, kHess ? 1 : 2) pose_solve_kernel(PoseParams P) {
    __shared__ SolverState S;
    __shared__ double s_red[(kPoseThreads / 32) * kAccHess];
    __shared__ PixelConsts C;
    const int group = blockIdx.x / P.blocks_per_group;
    const int blk = blockIdx.x - group * P.blocks_per_group;
    unsigned int *counter = P.counters + group * 32;
    double *gpart = P.partials + (size_t)group * 2 * kVirt * kAccHess;
    unsigned int barrier_no = 0;

    for (int pair = group; pair < P.p.n; pair += P.n_groups) {
        if (threadIdx.x == 0) {
            C.W = P.p.W, C.H = P.p.H, C.N = P.p.W * P.p.H;
            for (int k = 0; k < 9; ++k) C.K[k] = (double)P.p.K[pair * 9 + k];
            const double N = (double)C.N;
            C.s3 = (double)P.p.lw[pair * 2 + 0] / N;
            C.s2 = (double)P.p.lw[pair * 2 + 1] / N / ((double)C.H * (double)C.W);
            if (P.p.init_pose) {
                for (int k = 0; k < 3; ++k) S.X.t[k] = P.p.init_pose[pair * 7 + k];
                for (int k = 0; k < 4; ++k) S.X.q[k] = P.p.init_pose[pair * 7 + 3 + k];
            } else {
                S.X.t[0] = S.X.t[1] = S.X.t[2] = 0.0;
                S.X.q[0] = S.X.q[1] = S.X.q[2] = 0.0;
                S.X.q[3] = 1.0;
            }
            pose_to_rt(S.X, S.R);
            S.num_old = 0, S.n_iter = 0, S.evals = 0, S.status = 0, S.done = 0;
            S.H_diag = 1.0, S.t = 1.0, S.loss = 0.0, S.prev_loss = 0.0;
            for (int k = 0; k < 21; ++k) S.hess[k] = 0.0;
        }
        __syncthreads();
        const double lw0 = (double)P.p.lw[pair * 2 + 0], lw1 = (double)P.p.lw[pair * 2 + 1];

        while (true) {
            // ---- one fused evaluation at S.X
            double *slot_base = gpart + (size_t)(barrier_no & 1u) * kVirt * kAccHess;
            evaluate_group<kHess>(P, C, pair, blk, S, s_red, slot_base);
            barrier_no += 1;
            group_barrier(counter, barrier_no * (unsigned int)P.blocks_per_group);
            gather_totals<kHess>(P, C, slot_base, lw0, lw1, S);
            // ---- identical solver step in every CTA
            if (threadIdx.x == 0) {
                S.evals += 1;
                if (blk == 0) write_trace(P, pair, S);
                if (P.mode == RPE_SOLVER_EVAL_ONLY) {
                    S.done = 1;
                } else if (P.mode == RPE_SOLVER_GN) {
                    double step[6];
                    if (S.n_iter >= P.max_iter) {
                        S.done = 1;
                    } else if (!gn_step(S, step)) {
                        S.status = 1;
                        S.done = 1;
                    } else {
                        S.n_iter += 1;
                        se3_retract(S.X, step);
                        pose_to_rt(S.X, S.R);
                        if (absmax6(step) < 1e-12) S.done = 1;
                    }
                } else {
                    clip_gradient(S);
                    bool stop;
                    if (S.evals == 1) {
                        stop = absmax6(S.g) <= 1e-7;   // initial optimality check
                    } else {
                        stop = lbfgs_post_eval_stop(S, P.max_iter);
                    }
                    if (!stop && S.n_iter >= P.max_iter) stop = true;
                    while (!stop) {
                        bool need_eval;
                        stop = lbfgs_direction_and_step(S, P.max_iter, need_eval);
                        if (need_eval) break;
                    }
                    if (stop) S.done = 1;
                }
            }
            __syncthreads();
            if (S.done) break;
        }
        if (blk == 0 && threadIdx.x == 0) write_result(P, pair, S);
        __syncthreads();
    }
}

static int g_pose_groups = 0;    // upper bound on concurrently solved pairs; 0 = as many as the CTA slots allow
static int g_pose_bpg = 16;      // CTAs per group in batch mode (power of two <= kVirt)

}  // namespace rpe

extern "C" {

int rpe_pose_set_groups(int groups) {
    if (groups < 1 || groups > rpe::kMaxGroups) return RPE_ERR_INVALID_ARG;
    rpe::g_pose_groups = groups;
    return RPE_OK;
}

int rpe_pose_set_group_size(int ctas) {
    if (ctas < 1 || ctas > rpe::kVirt || (ctas & (ctas - 1))) return RPE_ERR_INVALID_ARG;
    rpe::g_pose_bpg = ctas;
    return RPE_OK;
}

size_t rpe_pose_workspace_bytes(int n_pairs) {
    (void)n_pairs;
    // per group two buffers of kVirt partial-sum records + 128-byte-spaced barrier counters
    return (size_t)rpe::kMaxGroups * 2 * rpe::kVirt * rpe::kAccHess * sizeof(double) + rpe::kMaxGroups * 128;
}

int rpe_pose_solve(const rpe_pose_problem *pb, int mode, int max_iter, int with_hessian, double *out, float *pose_f32,
                   float *log_f32, double *trace, int trace_cap, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace rpe;
    if (!pb || !out || !workspace) return RPE_ERR_INVALID_ARG;
    if (!pb->flow || !pb->pcl1 || !pb->pcl2 || !pb->m1 || !pb->m2 || !pb->K || !pb->lw) return RPE_ERR_INVALID_ARG;
    if (pb->n <= 0 || pb->H <= 0 || pb->W <= 0 || max_iter < 0) return RPE_ERR_INVALID_ARG;
    if (mode != RPE_SOLVER_LBFGS_REF && mode != RPE_SOLVER_GN && mode != RPE_SOLVER_EVAL_ONLY) return RPE_ERR_INVALID_ARG;
    const long long N = (long long)pb->H * pb->W;
    if (N % 4 != 0) return RPE_ERR_INVALID_ARG;
    if (!aligned16(pb->flow) || !aligned16(pb->pcl1) || !aligned16(pb->pcl2) || (pb->w1 && !aligned16(pb->w1)) ||
        (pb->w2 && !aligned16(pb->w2)) || (reinterpret_cast<uintptr_t>(pb->m1) & 3u) || (reinterpret_cast<uintptr_t>(pb->m2) & 3u))
        return RPE_ERR_ALIGNMENT;
    if (workspace_bytes < rpe_pose_workspace_bytes(pb->n)) return RPE_ERR_WORKSPACE;
    const bool hess = (mode == RPE_SOLVER_GN) || with_hessian;
    cudaStream_t st = (cudaStream_t)stream;

    const void *kern = hess ? (const void *)pose_solve_kernel<true> : (const void *)pose_solve_kernel<false>;
    int per_sm = 0;
    RPE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPoseThreads, 0));
    if (per_sm < 1) return RPE_ERR_UNSUPPORTED_DEVICE;
    if (per_sm > 2) per_sm = 2;
    const int sms = sm_count();
    if (sms <= 0) return RPE_ERR_CUDA;
    int max_blocks = sms * per_sm;
    if (max_blocks > 1024) max_blocks = 1024;

    // Group sizing.  A pair is split into kVirt virtual blocks, so at most kVirt CTAs can work on it: one or two pairs take
    // kVirt CTAs each (latency path).  A batch runs many small groups (g_pose_bpg CTAs, several virtual blocks per CTA and
    // evaluation): the serial part of an evaluation (group barrier, re-summation, the solver step on one thread) is hidden by
    // the other groups' arithmetic.  Measured on B200, 32 pairs of 640x512 (tools/pose_probe.py, profiles/): see DESIGN.md.
    int bpg = pb->n <= 2 ? kVirt : g_pose_bpg;
    while (bpg > 1 && bpg > max_blocks) bpg >>= 1;
    int n_groups = max_blocks / bpg;
    if (g_pose_groups > 0 && n_groups > g_pose_groups) n_groups = g_pose_groups;
    if (n_groups > pb->n) n_groups = pb->n;
    if (n_groups > kMaxGroups) n_groups = kMaxGroups;
    if (n_groups < 1) n_groups = 1;
    // pairs are dealt round-robin: balance the group count against the number of rounds (11 pairs on 4 groups = 3 rounds of 4, 4, 3)
    {
        const int rounds = (pb->n + n_groups - 1) / n_groups;
        n_groups = (pb->n + rounds - 1) / rounds;
    }
    // with few groups and free CTA slots, give every pair more CTAs (powers of two up to kVirt)
    while (bpg < kVirt && n_groups * bpg * 2 <= max_blocks) bpg <<= 1;

    PoseParams P;
    P.p = *pb;
    P.mode = mode, P.max_iter = max_iter, P.with_hessian = hess ? 1 : 0;
    P.blocks_per_group = bpg, P.n_groups = n_groups;
    P.out = out, P.pose_f32 = pose_f32, P.log_f32 = log_f32, P.trace = trace, P.trace_cap = trace_cap;
    P.partials = reinterpret_cast<double *>(workspace);
    P.counters = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(workspace) + (size_t)kMaxGroups * 2 * kVirt * kAccHess * sizeof(double));
    RPE_CUDA_TRY(cudaMemsetAsync(P.counters, 0, kMaxGroups * 128, st));
    void *args[] = {&P};
    RPE_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(bpg * n_groups), dim3(kPoseThreads), args, 0, st));
    ++g_launch_count;
    return RPE_OK;
}

}  // extern "C"
