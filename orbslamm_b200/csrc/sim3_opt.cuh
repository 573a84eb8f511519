// sim3_opt.cuh -- Optimizer::OptimizeSim3 (Optimizer.cc:1348-1543) as one persistent CTA per keyframe pair.
// One VertexSim3Expmap against fixed points, EdgeSim3ProjectXYZ (x1 = S12 X2) and EdgeInverseSim3ProjectXYZ (x2 = S21 X1) with Huber
// kernels, g2o's Levenberg on a dense 7x7 system (BlockSolverX + LinearSolverDense).  The edges have no analytic Jacobian in this g2o
// (types_seven_dof_expmap.h:147,169): BaseBinaryEdge::linearizeOplus differentiates numerically, central differences with delta = 1e-9
// through oplusImpl (base_binary_edge.hpp:131-205).  The 14 perturbed transforms depend only on the current estimate, so they (and
// their inverses) are built once per linearisation in shared memory and every correspondence evaluates its two edges against them.
// Whole 5 + (5 | 10) iteration schedule on the device, no host round trip; fixed-order block reductions.
#pragma once
#include "common.cuh"
#include "pose_opt.cuh"

namespace orbs {

struct Sim3 { double q[4]; double t[3]; double s; };

__device__ inline void sim3_exp(const double u[7], Sim3 &out)                  // Sim3(const Vector7d&), types/sim3.h:45-104
{
    const double w[3] = {u[0], u[1], u[2]}, sigma = u[6];
    const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9], R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
    const double s = exp(sigma), eps = 0.00001;
    double A, B, C;
    if (fabs(sigma) < eps) {
        C = 1;
        if (theta < eps) {
            A = 1. / 2.; B = 1. / 6.;
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i];
        } else {
            const double theta2 = theta * theta;
            A = (1 - cos(theta)) / (theta2);
            B = (theta - sin(theta)) / (theta2 * theta);
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + sin(theta) / theta * O[i] + (1 - cos(theta)) / (theta * theta) * O2[i];
        }
    } else {
        C = (s - 1) / sigma;
        if (theta < eps) {
            const double sigma2 = sigma * sigma;
            A = ((sigma - 1) * s + 1) / sigma2;
            B = ((0.5 * sigma2 - sigma + 1) * s) / (sigma2 * sigma);
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i];
        } else {
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + sin(theta) / theta * O[i] + (1 - cos(theta)) / (theta * theta) * O2[i];
            const double a = s * sin(theta), b = s * cos(theta), theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
            A = (a * sigma + (1 - b) * theta) / (theta * c);
            B = (C - ((b - 1) * sigma + a * theta) / (c)) * 1. / (theta2);
        }
    }
    quat_from_R(R, out.q);
    for (int r = 0; r < 3; r++) {
        const double W0 = A * O[3 * r] + B * O2[3 * r] + C * (r == 0 ? 1.0 : 0.0), W1 = A * O[3 * r + 1] + B * O2[3 * r + 1] + C * (r == 1 ? 1.0 : 0.0),
                     W2 = A * O[3 * r + 2] + B * O2[3 * r + 2] + C * (r == 2 ? 1.0 : 0.0);
        out.t[r] = W0 * u[3] + W1 * u[4] + W2 * u[5];
    }
    out.s = s;
}

__device__ __forceinline__ void sim3_map(const Sim3 &S, const double x[3], double out[3])       // s*(r*xyz) + t
{
    double rx[3];
    quat_rotate(S.q, x, rx);
#pragma unroll
    for (int k = 0; k < 3; k++) out[k] = S.s * rx[k] + S.t[k];
}

__device__ inline void sim3_inverse(const Sim3 &S, Sim3 &out)                  // Sim3(r.conjugate(), r.conjugate()*((-1./s)*t), 1./s)
{
    const double qc[4] = {-S.q[0], -S.q[1], -S.q[2], S.q[3]};
    const double f = -1. / S.s, st[3] = {f * S.t[0], f * S.t[1], f * S.t[2]};
    quat_rotate(qc, st, out.t);
    out.q[0] = qc[0]; out.q[1] = qc[1]; out.q[2] = qc[2]; out.q[3] = qc[3];
    out.s = 1. / S.s;
}

__device__ inline void sim3_mul(const Sim3 &a, const Sim3 &b, Sim3 &o)         // operator*, sim3.h:214-220 (no renormalisation)
{
    Sim3 r;
    quat_mul(a.q, b.q, r.q);
    double rt[3];
    quat_rotate(a.q, b.t, rt);
    for (int k = 0; k < 3; k++) r.t[k] = a.s * rt[k] + a.t[k];
    r.s = a.s * b.s;
    o = r;
}

// VertexSim3Expmap::oplusImpl, types_seven_dof_expmap.h:54-63
__device__ inline void sim3_oplus(const Sim3 &S, const double *upd, bool fix_scale, Sim3 &out)
{
    double u[7];
    for (int k = 0; k < 7; k++) u[k] = upd[k];
    if (fix_scale) u[6] = 0;
    Sim3 d;
    sim3_exp(u, d);
    sim3_mul(d, S, out);
}

struct Sim3Args {
    int slab;
    double *sim3;                // [n_pairs, 8] in/out: r (x y z w), t, s
    const uint8_t *valid;        // [n_pairs*slab]
    const float *P1c, *P2c, *obs1, *obs2, *w1, *w2;
    const float *K1, *K2;        // [n_pairs, 4]
    const int *counts;
    float th2; int fix_scale;
    uint8_t *inlier;             // [n_pairs*slab] out
    int *n_inliers;              // [n_pairs] out
    double *err;                 // scratch [n_pairs*slab, 4]: stored _error of e12 / e21
    uint8_t *active;             // scratch [n_pairs*slab]
    int *stats;                  // optional [n_pairs, 2]: LM iterations, trials
};

constexpr int kSim3Threads = 128;
constexpr int kSim3NV = 36;      // 28 (upper H) + 7 (b) + 1 (chi2)

template <int N>
__device__ inline bool solve_ldlt(const double *Hu /*upper packed*/, const double *b, double lambda, double *x)
{
    double A[N * N];
    int p = 0;
    for (int r = 0; r < N; r++) for (int c = r; c < N; c++) { A[N * r + c] = Hu[p]; A[N * c + r] = Hu[p]; p++; }
    for (int i = 0; i < N; i++) A[(N + 1) * i] += lambda;
    for (int i = 0; i < N; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[N * i + j];
            for (int k = 0; k < j; k++) s -= A[N * i + k] * A[N * j + k] * A[(N + 1) * k];
            if (j < i) A[N * i + j] = s / A[(N + 1) * j];
            else { if (!(s > 0.0)) return false; A[(N + 1) * i] = s; }
        }
    for (int i = 0; i < N; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= A[N * i + k] * x[k]; x[i] = s; }
    for (int i = 0; i < N; i++) x[i] /= A[(N + 1) * i];
    for (int i = N - 1; i >= 0; i--) { const double xi = x[i]; for (int k = 0; k < i; k++) x[k] -= A[N * i + k] * xi; }
    return true;
}

__device__ __forceinline__ void sim3_edge_errors(const Sim3 &S, const Sim3 &Si, const double X1[3], const double X2[3], const double o[4], const double K1[4],
                                                 const double K2[4], double e[4])
{
    double p[3];
    sim3_map(S, X2, p);                                            // EdgeSim3ProjectXYZ::computeError
    e[0] = o[0] - (p[0] / p[2] * K1[0] + K1[2]);
    e[1] = o[1] - (p[1] / p[2] * K1[1] + K1[3]);
    sim3_map(Si, X1, p);                                           // EdgeInverseSim3ProjectXYZ::computeError
    e[2] = o[2] - (p[0] / p[2] * K2[0] + K2[2]);
    e[3] = o[3] - (p[1] / p[2] * K2[1] + K2[3]);
}

__device__ __forceinline__ double chi2_iso(double e0, double e1, double w) { return e0 * (w * e0 + 0.0 * e1) + e1 * (0.0 * e0 + w * e1); }

__global__ void __launch_bounds__(kSim3Threads)
k_optimize_sim3(const Sim3Args A)
{
    __shared__ double s_warp[(kSim3Threads / 32) * kSim3NV];
    __shared__ double s_red[kSim3NV];
    __shared__ Sim3 s_S, s_Si, s_backup, s_pert[14], s_perti[14];
    __shared__ double s_lambda, s_ni, s_cur_chi, s_x[7], s_rho, s_chi_out[1];
    __shared__ int s_nbad_lm, s_flag, s_ok2, s_cnt;
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int M = A.counts[f];
    const size_t o = (size_t)f * A.slab;
    const float *P1c = A.P1c + 3 * o, *P2c = A.P2c + 3 * o, *obs1 = A.obs1 + 2 * o, *obs2 = A.obs2 + 2 * o, *w1 = A.w1 + o, *w2 = A.w2 + o;
    uint8_t *active = A.active + o, *inlier = A.inlier + o;
    double *err = A.err + 4 * o;
    double K1[4], K2[4];
    for (int k = 0; k < 4; k++) { K1[k] = (double)A.K1[4 * f + k]; K2[k] = (double)A.K2[4 * f + k]; }
    const double th2 = (double)A.th2;
    const double delta = (double)sqrtf(A.th2), dsqr = (double)(float)(delta * delta);       // const float deltaHuber = sqrt(th2), Optimizer.cc:1395; delta^2 is a float member of RobustKernelHuber
    const bool fix_scale = A.fix_scale != 0;
    int my = 0;
    for (int i = tid; i < M; i += nt) { const uint8_t v = A.valid[o + i] ? 1 : 0; active[i] = v; inlier[i] = v; my += v; }
    if (tid == 0) {
        for (int k = 0; k < 4; k++) s_S.q[k] = A.sim3[8 * f + k];
        for (int k = 0; k < 3; k++) s_S.t[k] = A.sim3[8 * f + 4 + k];
        s_S.s = A.sim3[8 * f + 7];
        s_cnt = 0;
        if (A.stats) { A.stats[2 * f] = 0; A.stats[2 * f + 1] = 0; }
    }
    __syncthreads();
    if (my) atomicAdd(&s_cnt, my);
    __syncthreads();
    const int n_corr = s_cnt;
    int n_bad = 0, n_in = 0;

    for (int pass = 0; pass < 2; pass++) {
        const int iterations = pass == 0 ? 5 : (n_bad > 0 ? 10 : 5);
        const int n_act = n_corr - n_bad;
        if (pass == 1 && n_act < 10) break;                                   // Optimizer.cc:1497-1498: return 0, g2oS12 untouched
        if (n_act > 0) {
            for (int iter = 0; iter < iterations; iter++) {
                // the 14 perturbed estimates of the numeric Jacobian and the inverse of the current one
                if (tid < 14) {
                    double add[7] = {0, 0, 0, 0, 0, 0, 0};
                    add[tid >> 1] = (tid & 1) ? -1e-9 : 1e-9;
                    Sim3 sp, spi;
                    sim3_oplus(s_S, add, fix_scale, sp);
                    sim3_inverse(sp, spi);
                    s_pert[tid] = sp; s_perti[tid] = spi;
                } else if (tid == 32) { Sim3 si; sim3_inverse(s_S, si); s_Si = si; }
                __syncthreads();
                // computeActiveErrors + activeRobustChi2 + buildSystem
                double acc[kSim3NV];
#pragma unroll
                for (int v = 0; v < kSim3NV; v++) acc[v] = 0;
                for (int i = tid; i < M; i += nt) {
                    if (!active[i]) continue;
                    const double X1[3] = {(double)P1c[3 * i], (double)P1c[3 * i + 1], (double)P1c[3 * i + 2]};
                    const double X2[3] = {(double)P2c[3 * i], (double)P2c[3 * i + 1], (double)P2c[3 * i + 2]};
                    const double ob[4] = {(double)obs1[2 * i], (double)obs1[2 * i + 1], (double)obs2[2 * i], (double)obs2[2 * i + 1]};
                    double e[4], J[4][7];
                    sim3_edge_errors(s_S, s_Si, X1, X2, ob, K1, K2, e);
                    err[4 * i] = e[0]; err[4 * i + 1] = e[1]; err[4 * i + 2] = e[2]; err[4 * i + 3] = e[3];
                    const double scalar = 1.0 / (2 * 1e-9);
#pragma unroll
                    for (int d = 0; d < 7; d++) {
                        double ep[4], em[4];
                        sim3_edge_errors(s_pert[2 * d], s_perti[2 * d], X1, X2, ob, K1, K2, ep);
                        sim3_edge_errors(s_pert[2 * d + 1], s_perti[2 * d + 1], X1, X2, ob, K1, K2, em);
#pragma unroll
                        for (int r = 0; r < 4; r++) J[r][d] = scalar * (ep[r] - em[r]);
                    }
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const double w = (double)(k ? w2[i] : w1[i]);
                        const double e0 = e[2 * k], e1 = e[2 * k + 1];
                        const double c = chi2_iso(e0, e1, w);
                        double rho0 = c, rho1 = 1.0;
                        huber(c, delta, dsqr, rho0, rho1);
                        acc[35] += rho0;
                        // g2o's evaluation order (base_binary_edge.hpp:55-120, pinned against the reference's object code): omega_r = -(information * error),
                        // then *= rho'; the block entry that Eigen's LDLT reads is the LOWER one, (J_cc w) J_a
                        const double wo = rho1 * w, r0 = -(w * e0) * rho1, r1 = -(w * e1) * rho1;
                        int p = 0;
#pragma unroll
                        for (int a = 0; a < 7; a++) {
                            acc[28 + a] += J[2 * k][a] * r0 + J[2 * k + 1][a] * r1;
#pragma unroll
                            for (int cc = a; cc < 7; cc++) { acc[p] += J[2 * k][cc] * wo * J[2 * k][a] + J[2 * k + 1][cc] * wo * J[2 * k + 1][a]; p++; }
                        }
                    }
                }
                block_reduce_vec<kSim3NV>(acc, s_warp, s_red);
                if (tid == 0) {
                    s_cur_chi = s_red[35];
                    if (iter == 0) {                                              // computeLambdaInit
                        double mx = 0.;
                        int p = 0;
                        for (int a = 0; a < 7; a++) { mx = fmax(fabs(s_red[p]), mx); p += 7 - a; }
                        s_lambda = 1e-5 * mx; s_ni = 2; s_nbad_lm = 0;
                    }
                }
                __syncthreads();
                const double ini_chi = s_cur_chi;
                int qmax = 0;
                double rho = 0;
                do {
                    if (tid == 0) {
                        s_backup = s_S;
                        s_ok2 = solve_ldlt<7>(s_red, s_red + 28, s_lambda, s_x) ? 1 : 0;
                        Sim3 r, ri;
                        sim3_oplus(s_S, s_x, fix_scale, r);
                        sim3_inverse(r, ri);
                        s_S = r; s_Si = ri;
                    }
                    __syncthreads();
                    double chi[1] = {0};
                    for (int i = tid; i < M; i += nt) {
                        if (!active[i]) continue;
                        const double X1[3] = {(double)P1c[3 * i], (double)P1c[3 * i + 1], (double)P1c[3 * i + 2]};
                        const double X2[3] = {(double)P2c[3 * i], (double)P2c[3 * i + 1], (double)P2c[3 * i + 2]};
                        const double ob[4] = {(double)obs1[2 * i], (double)obs1[2 * i + 1], (double)obs2[2 * i], (double)obs2[2 * i + 1]};
                        double e[4];
                        sim3_edge_errors(s_S, s_Si, X1, X2, ob, K1, K2, e);
                        err[4 * i] = e[0]; err[4 * i + 1] = e[1]; err[4 * i + 2] = e[2]; err[4 * i + 3] = e[3];
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const double c = chi2_iso(e[2 * k], e[2 * k + 1], (double)(k ? w2[i] : w1[i]));
                            double rho0 = c, rho1 = 1.0;
                            huber(c, delta, dsqr, rho0, rho1);
                            chi[0] += rho0;
                        }
                    }
                    block_reduce_vec<1>(chi, s_warp, s_chi_out);
                    if (tid == 0) {
                        double temp_chi = s_chi_out[0];
                        if (!s_ok2) temp_chi = 1.7976931348623157e308;
                        double r = s_cur_chi - temp_chi;
                        double scale = 0.;
                        for (int j = 0; j < 7; j++) scale += s_x[j] * (s_lambda * s_x[j] + s_red[28 + j]);
                        scale += 1e-3;
                        r /= scale;
                        if (r > 0 && isfinite(temp_chi)) {
                            double alpha = 1. - pow((2 * r - 1), 3.0);
                            alpha = fmin(alpha, 2. / 3.);
                            s_lambda *= fmax(1. / 3., alpha);
                            s_ni = 2; s_cur_chi = temp_chi;
                        } else {
                            s_lambda *= s_ni; s_ni *= 2;
                            s_S = s_backup;
                        }
                        s_rho = r;
                        if (A.stats) A.stats[2 * f + 1]++;
                    }
                    __syncthreads();
                    rho = s_rho;
                    qmax++;
                    __syncthreads();
                } while (rho < 0 && qmax < 10);
                if (tid == 0 && A.stats) A.stats[2 * f]++;
                bool terminate = (qmax == 10 || rho == 0);
                if (!terminate) {
                    if (tid == 0) {
                        if ((ini_chi - s_cur_chi) * 1e3 < ini_chi) s_nbad_lm++; else s_nbad_lm = 0;
                        s_flag = s_nbad_lm >= 3;
                    }
                    __syncthreads();
                    terminate = s_flag != 0;
                    __syncthreads();
                }
                if (terminate) break;
            }
        }
        // inlier check on the stored errors (Optimizer.cc:1469-1486, 1506-1520)
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        int cnt = 0;
        for (int i = tid; i < M; i += nt) {
            if (!active[i]) continue;
            const bool bad = chi2_iso(err[4 * i], err[4 * i + 1], (double)w1[i]) > th2 || chi2_iso(err[4 * i + 2], err[4 * i + 3], (double)w2[i]) > th2;
            if (pass == 0) { if (bad) { active[i] = 0; inlier[i] = 0; cnt++; } }
            else { if (bad) inlier[i] = 0; else cnt++; }
        }
        if (cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (pass == 0) n_bad = s_cnt; else n_in = s_cnt;
        __syncthreads();
        if (pass == 1 && tid == 0) {
            for (int k = 0; k < 4; k++) A.sim3[8 * f + k] = s_S.q[k];
            for (int k = 0; k < 3; k++) A.sim3[8 * f + 4 + k] = s_S.t[k];
            A.sim3[8 * f + 7] = s_S.s;
        }
    }
    if (tid == 0) A.n_inliers[f] = n_in;
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// Sim3Solver (S/src/Sim3Solver.cc): the data-parallel part of the RANSAC.  ComputeSim3 (Horn's closed form on 3 points, a handful of OpenCV
// small-matrix calls) stays with the caller; what costs time in the reference is CheckInliers -- two cv::Mat projections per correspondence and
// hypothesis, each through freshly allocated 3x1 / 2x1 matrices (:392-417) -- and that is one thread per (hypothesis, correspondence) here.
struct Sim3CheckArgs {
    int n_hyp, N;
    const float *T12, *T21;            // [n_hyp, 16]
    const float *X1, *X2, *P1im1, *P2im2;
    const int *max_err1, *max_err2;    // the reference's std::vector<size_t>: 9.210 * sigma^2 truncated
    float K1[4], K2[4];
    uint8_t *inliers;                  // [n_hyp, N]
    int *n_inliers;                    // [n_hyp], zeroed
};

__device__ __forceinline__ void sim3_project_f32(const float *T, const float *K, const float *P, float &u, float &v)   // Project, :392-417
{
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float s = __fmul_rn(T[4 * r], P[0]);
        s = __fadd_rn(s, __fmul_rn(T[4 * r + 1], P[1]));
        s = __fadd_rn(s, __fmul_rn(T[4 * r + 2], P[2]));
        pc[r] = __fadd_rn(s, T[4 * r + 3]);
    }
    const float invz = __fdiv_rn(1.0f, pc[2]);
    u = __fadd_rn(__fmul_rn(K[0], __fmul_rn(pc[0], invz)), K[2]);
    v = __fadd_rn(__fmul_rn(K[1], __fmul_rn(pc[1], invz)), K[3]);
}

__global__ void __launch_bounds__(256)
k_sim3_check_inliers(const Sim3CheckArgs A)
{
    __shared__ float s_T[32];
    __shared__ int s_n;
    const int h = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x < 16) s_T[threadIdx.x] = A.T12[16 * h + threadIdx.x];
    else if (threadIdx.x < 32) s_T[threadIdx.x] = A.T21[16 * h + threadIdx.x - 16];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    bool in = false;
    if (i < A.N) {
        const float X1[3] = {A.X1[3 * i], A.X1[3 * i + 1], A.X1[3 * i + 2]}, X2[3] = {A.X2[3 * i], A.X2[3 * i + 1], A.X2[3 * i + 2]};
        float u21, v21, u12, v12;
        sim3_project_f32(s_T, A.K1, X2, u21, v21);                  // Project(mvX3Dc2, vP2im1, mT12i, mK1)
        sim3_project_f32(s_T + 16, A.K2, X1, u12, v12);             // Project(mvX3Dc1, vP1im2, mT21i, mK2)
        const float d1x = __fsub_rn(A.P1im1[2 * i], u21), d1y = __fsub_rn(A.P1im1[2 * i + 1], v21);
        const float d2x = __fsub_rn(u12, A.P2im2[2 * i]), d2y = __fsub_rn(v12, A.P2im2[2 * i + 1]);
        const float err1 = (float)__dadd_rn(__dmul_rn((double)d1x, (double)d1x), __dmul_rn((double)d1y, (double)d1y));   // Mat::dot: double accumulation
        const float err2 = (float)__dadd_rn(__dmul_rn((double)d2x, (double)d2x), __dmul_rn((double)d2y, (double)d2y));
        in = err1 < (float)A.max_err1[i] && err2 < (float)A.max_err2[i];
        A.inliers[(size_t)h * A.N + i] = in ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&s_n, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0 && s_n) atomicAdd(&A.n_inliers[h], s_n);
}

// the constructor's per-correspondence data: integer-truncated thresholds (:88-89) and FromCameraToImage (:419-437)
__global__ void __launch_bounds__(256)
k_sim3_prepare(int N, const float *__restrict__ X3Dc, const int *__restrict__ octave, const float *__restrict__ level_sigma2, int nlevels, float fx, float fy,
               float cx, float cy, int *__restrict__ max_err, float *__restrict__ p2d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float invz = __fdiv_rn(1.0f, X3Dc[3 * i + 2]);
    p2d[2 * i] = __fadd_rn(__fmul_rn(fx, __fmul_rn(X3Dc[3 * i], invz)), cx);
    p2d[2 * i + 1] = __fadd_rn(__fmul_rn(fy, __fmul_rn(X3Dc[3 * i + 1], invz)), cy);
    max_err[i] = (int)(unsigned long long)__dmul_rn(9.210, (double)level_sigma2[min(max(octave[i], 0), nlevels - 1)]);
}

// ---------------------------------------------------------------------------------------------------------------------
// Sim3Solver::ComputeSim3 (Sim3Solver.cc:226-338; Horn 1987, closed form on three point pairs) for a batch of RANSAC hypotheses, one thread each.
// The reference runs it through OpenCV float matrices; this restates the arithmetic with OpenCV's evaluation order and precisions where they are known:
//   centroid     cv::reduce(SUM) in float, then `C / 3` = float(double(C) * (1.0 / 3))                          (MatExpr scale)
//   M            Pr2 * Pr1.t(): a gemm with a transposed operand -> OpenCV's general path, double accumulation, rounded to float
//   N            doubles from the float entries of M, stored as float (Mat_<float> <<)
//   eigen        cv::eigen of a symmetric 4x4 float matrix = Jacobi rotations in float (JacobiImpl_: pivot = largest off-diagonal entry tracked per row /
//                column, 30 n^2 sweeps at most, eigenvalues sorted descending); the quaternion is the first eigenvector
//   angle-axis   norm() and atan2 in double; `2*ang*vec/norm(vec)` = float(double(v) * ((2 ang) * (1 / norm)))  (MatExpr scale)
//   Rodrigues    in double (theta, c, s, c1, r / theta), rounded to float
//   P3 = R Pr2   3x3 small-matrix gemm: float products summed left to right;  nom = Pr1.dot(P3) and den = sum(P3^2) accumulate in double
//   t12, T12, T21  scaled small gemms: float(double(float dot) * alpha)
// cv::eigen is build dependent (an OpenCV built with Eigen uses SelfAdjointEigenSolver instead of Jacobi) and the quaternion's sign is arbitrary, so this path is
// pinned to a float TOLERANCE against the cv2-backed restatement (tests/test_sim3_compute_gpu.py), not bit for bit; the drop-in keeps the reference's own host
// ComputeSim3 unless built with -DORBSLAMM_DEVICE_COMPUTE_SIM3.
__device__ __forceinline__ float s3_hypotf(float a, float b)
{
    a = fabsf(a); b = fabsf(b);                                            // OpenCV's hypot(): scaled form
    if (a > b) { b = __fdiv_rn(b, a); return __fmul_rn(a, sqrtf(__fadd_rn(1.f, __fmul_rn(b, b)))); }
    if (b > 0) { a = __fdiv_rn(a, b); return __fmul_rn(b, sqrtf(__fadd_rn(1.f, __fmul_rn(a, a)))); }
    return 0.f;
}

__device__ inline void s3_jacobi4(float (&A)[4][4], float (&W)[4], float (&V)[4][4])
{
    constexpr int n = 4;
    const float eps = 1.1920929e-07f;
    int indR[n], indC[n];
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) V[i][j] = 0.f; V[i][i] = 1.f; }
    for (int k = 0; k < n; k++) {
        W[k] = A[k][k];
        if (k < n - 1) { int m = k + 1; float mv = fabsf(A[k][m]); for (int i = k + 2; i < n; i++) { const float v = fabsf(A[k][i]); if (mv < v) { mv = v; m = i; } } indR[k] = m; }
        if (k > 0) { int m = 0; float mv = fabsf(A[0][k]); for (int i = 1; i < k; i++) { const float v = fabsf(A[i][k]); if (mv < v) { mv = v; m = i; } } indC[k] = m; }
    }
    for (int iters = 0; iters < n * n * 30; iters++) {
        int k = 0; float mv = fabsf(A[0][indR[0]]);
        for (int i = 1; i < n - 1; i++) { const float v = fabsf(A[i][indR[i]]); if (mv < v) { mv = v; k = i; } }
        int l = indR[k];
        for (int i = 1; i < n; i++) { const float v = fabsf(A[indC[i]][i]); if (mv < v) { mv = v; k = indC[i]; l = i; } }
        const float p = A[k][l];
        if (fabsf(p) <= eps) break;
        const float y = (float)(((double)__fsub_rn(W[l], W[k])) * 0.5);
        float t = __fadd_rn(fabsf(y), s3_hypotf(p, y));
        float sn = s3_hypotf(p, t);
        const float c = __fdiv_rn(t, sn);
        sn = __fdiv_rn(p, sn); t = __fmul_rn(__fdiv_rn(p, t), p);
        if (y < 0) { sn = -sn; t = -t; }
        A[k][l] = 0.f;
        W[k] = __fsub_rn(W[k], t); W[l] = __fadd_rn(W[l], t);
#define S3_ROT(v0, v1) { const float a0 = v0, b0 = v1; v0 = __fsub_rn(__fmul_rn(a0, c), __fmul_rn(b0, sn)); v1 = __fadd_rn(__fmul_rn(a0, sn), __fmul_rn(b0, c)); }
        for (int i = 0; i < k; i++) S3_ROT(A[i][k], A[i][l]);
        for (int i = k + 1; i < l; i++) S3_ROT(A[k][i], A[i][l]);
        for (int i = l + 1; i < n; i++) S3_ROT(A[k][i], A[l][i]);
        for (int i = 0; i < n; i++) S3_ROT(V[k][i], V[l][i]);
#undef S3_ROT
        for (int j = 0; j < 2; j++) {
            const int idx = j == 0 ? k : l;
            if (idx < n - 1) { int m = idx + 1; float mx = fabsf(A[idx][m]); for (int i = idx + 2; i < n; i++) { const float v = fabsf(A[idx][i]); if (mx < v) { mx = v; m = i; } } indR[idx] = m; }
            if (idx > 0) { int m = 0; float mx = fabsf(A[0][idx]); for (int i = 1; i < idx; i++) { const float v = fabsf(A[i][idx]); if (mx < v) { mx = v; m = i; } } indC[idx] = m; }
        }
    }
    for (int k = 0; k < n - 1; k++) {                                      // eigenvalues descending, eigenvectors in rows
        int m = k;
        for (int i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
        if (k != m) { const float w = W[m]; W[m] = W[k]; W[k] = w; for (int i = 0; i < n; i++) { const float v = V[m][i]; V[m][i] = V[k][i]; V[k][i] = v; } }
    }
}

// X1 / X2: f32[n_hyp, 3 points, 3 coordinates] (the min sets, camera-1 / camera-2 coordinates).  Out: T12, T21 f32[n_hyp, 16]; Rts f32[n_hyp, 13] = R12 (9), t12 (3), s12.
__global__ void __launch_bounds__(128)
k_sim3_compute(int n_hyp, const float *__restrict__ X1, const float *__restrict__ X2, int fix_scale, float *__restrict__ T12, float *__restrict__ T21, float *__restrict__ Rts)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hyp) return;
    float P1[3][3], P2[3][3], O1[3], O2[3];                                // [coordinate][point]
    for (int i = 0; i < 3; i++) for (int r = 0; r < 3; r++) { P1[r][i] = X1[9 * (size_t)h + 3 * i + r]; P2[r][i] = X2[9 * (size_t)h + 3 * i + r]; }
    for (int r = 0; r < 3; r++) {
        O1[r] = (float)__dmul_rn((double)__fadd_rn(__fadd_rn(P1[r][0], P1[r][1]), P1[r][2]), 1.0 / 3.0);
        O2[r] = (float)__dmul_rn((double)__fadd_rn(__fadd_rn(P2[r][0], P2[r][1]), P2[r][2]), 1.0 / 3.0);
        for (int i = 0; i < 3; i++) { P1[r][i] = __fsub_rn(P1[r][i], O1[r]); P2[r][i] = __fsub_rn(P2[r][i], O2[r]); }
    }
    float M[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double a = 0;
        for (int k = 0; k < 3; k++) a = __dadd_rn(a, __dmul_rn((double)P2[r][k], (double)P1[c][k]));
        M[r][c] = (float)a;
    }
    const double m00 = M[0][0], m01 = M[0][1], m02 = M[0][2], m10 = M[1][0], m11 = M[1][1], m12 = M[1][2], m20 = M[2][0], m21 = M[2][1], m22 = M[2][2];
    const float N11 = (float)(m00 + m11 + m22), N12 = (float)(m12 - m21), N13 = (float)(m20 - m02), N14 = (float)(m01 - m10), N22 = (float)(m00 - m11 - m22),
                N23 = (float)(m01 + m10), N24 = (float)(m20 + m02), N33 = (float)(-m00 + m11 - m22), N34 = (float)(m12 + m21), N44 = (float)(-m00 - m11 + m22);
    float A[4][4] = {{N11, N12, N13, N14}, {N12, N22, N23, N24}, {N13, N23, N33, N34}, {N14, N24, N34, N44}}, W[4], V[4][4];
    s3_jacobi4(A, W, V);
    const double vx = V[0][1], vy = V[0][2], vz = V[0][3];
    const double nrm = sqrt(vx * vx + vy * vy + vz * vz);
    const double ang = atan2(nrm, (double)V[0][0]);
    const double alpha = (2 * ang) * (1.0 / nrm);
    const float rv[3] = {(float)(vx * alpha), (float)(vy * alpha), (float)(vz * alpha)};
    float R[3][3];
    {   // cv::Rodrigues(vector -> matrix), double inside
        const double rx0 = rv[0], ry0 = rv[1], rz0 = rv[2];
        const double theta = sqrt(rx0 * rx0 + ry0 * ry0 + rz0 * rz0);
        if (!(theta >= 2.220446049250313e-16)) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r][c] = r == c ? 1.f : 0.f; }
        else {
            const double c = cos(theta), sn = sin(theta), c1 = 1. - c, it = 1. / theta;
            const double rx = rx0 * it, ry = ry0 * it, rz = rz0 * it;
            R[0][0] = (float)(c + c1 * rx * rx); R[0][1] = (float)(c1 * rx * ry - sn * rz); R[0][2] = (float)(c1 * rx * rz + sn * ry);
            R[1][0] = (float)(c1 * rx * ry + sn * rz); R[1][1] = (float)(c + c1 * ry * ry); R[1][2] = (float)(c1 * ry * rz - sn * rx);
            R[2][0] = (float)(c1 * rx * rz - sn * ry); R[2][1] = (float)(c1 * ry * rz + sn * rx); R[2][2] = (float)(c + c1 * rz * rz);
        }
    }
    float s12 = 1.0f;
    if (!fix_scale) {
        double nom = 0, den = 0;
        for (int r = 0; r < 3; r++) for (int i = 0; i < 3; i++) {
            const float p3 = __fadd_rn(__fadd_rn(__fmul_rn(R[r][0], P2[0][i]), __fmul_rn(R[r][1], P2[1][i])), __fmul_rn(R[r][2], P2[2][i]));
            nom = __dadd_rn(nom, __dmul_rn((double)P1[r][i], (double)p3));
            den = __dadd_rn(den, (double)__fmul_rn(p3, p3));
        }
        s12 = (float)(nom / den);
    }
    float t12[3];
    for (int r = 0; r < 3; r++) {
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(R[r][0], O2[0]), __fmul_rn(R[r][1], O2[1])), __fmul_rn(R[r][2], O2[2]));
        t12[r] = __fsub_rn(O1[r], (float)__dmul_rn((double)d, (double)s12));
    }
    float *o12 = T12 + 16 * (size_t)h, *o21 = T21 + 16 * (size_t)h, *ort = Rts + 13 * (size_t)h;
    float sRi[3][3];
    const double is = 1.0 / (double)s12;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) { o12[4 * r + c] = (float)__dmul_rn((double)R[r][c], (double)s12); sRi[r][c] = (float)__dmul_rn((double)R[c][r], is); o21[4 * r + c] = sRi[r][c]; ort[3 * r + c] = R[r][c]; }
        o12[4 * r + 3] = t12[r]; ort[9 + r] = t12[r];
    }
    for (int r = 0; r < 3; r++) {
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(sRi[r][0], t12[0]), __fmul_rn(sRi[r][1], t12[1])), __fmul_rn(sRi[r][2], t12[2]));
        o21[4 * r + 3] = (float)__dmul_rn((double)d, -1.0);
    }
    for (int c = 0; c < 3; c++) { o12[12 + c] = 0.f; o21[12 + c] = 0.f; }
    o12[15] = 1.f; o21[15] = 1.f; ort[12] = s12;
}

}  // namespace orbs
