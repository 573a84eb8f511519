// pose_opt.cuh -- Optimizer::PoseOptimization (Optimizer.cc:262-474) as one persistent CTA per frame.
// The whole 4 x 10 Levenberg-Marquardt schedule of the reference (g2o OptimizationAlgorithmLevenberg on a single
// VertexSE3Expmap with unary EdgeSE3ProjectXYZOnlyPose edges, LinearSolverDense 6x6) runs on the device with no
// host round trip; block-level reductions are fixed-order (deterministic).
#pragma once
#include "common.cuh"
#include "se3.cuh"

namespace orbs {

struct PoseArgs {
    int slab;
    double fx, fy, cx, cy;
    float *Tcw;                 // [n_frames, 16] in/out
    const float *Xw, *obs, *w;  // [n_frames*slab, 3|2|1]
    const int *counts;
    uint8_t *outlier;           // [n_frames*slab]
    int *n_inliers;             // [n_frames]
    double *err;                // scratch [n_frames*slab, 2]: the stored _error of every edge
};

constexpr int kPoseThreads = 256;
constexpr int kPoseNV = 28;      // 21 (upper H) + 6 (b) + 1 (chi2)

// fixed-order block reduction of NV doubles per thread; result valid in out[0..NV) for all threads after return
template <int NV>
__device__ __forceinline__ void block_reduce_vec(double (&acc)[NV], double *warp_buf /*[nwarps*NV]*/, double *out /*[NV]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        double x = acc[v];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
        if (lane == 0) warp_buf[wid * NV + v] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < nw; w++) s += warp_buf[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// LDL^T solve of the 6x6 system (H + lambda I) x = b; H given by its upper triangle (row-major packed 21).
// LinearSolverDense: fails unless the factorisation is positive (linear_solver_dense.h:107-112).
__device__ __forceinline__ bool solve6(const double *Hu, const double *b, double lambda, double *x)
{
    // every loop fully unrolled: the 6x6 matrix lives in registers (with runtime loop bounds it went to local memory, and this runs on ONE thread while
    // the other 255 wait -- it was the longest serial stretch of an LM trial)
    double A[6][6];
    {
        int p = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = r; c < 6; c++) { A[r][c] = Hu[p]; A[c][r] = Hu[p]; p++; }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) A[i][i] += lambda;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = A[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k] * A[k][k];
            if (j < i) A[i][j] = s / A[j][j];
            else { if (!(s > 0.0)) ok = false; A[i][i] = s; }
        }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { double s = b[i]; 
#pragma unroll
        for (int k = 0; k < i; k++) s -= A[i][k] * y[k]; y[i] = s; }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] /= A[i][i];
#pragma unroll
    for (int i = 5; i >= 0; i--) { const double xi = y[i]; 
#pragma unroll
        for (int k = 0; k < i; k++) y[k] -= A[i][k] * xi; }
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = y[i];
    return true;
}

__global__ void __launch_bounds__(kPoseThreads)
k_pose_optimization(const PoseArgs A)
{
    __shared__ double s_warp[(kPoseThreads / 32) * kPoseNV];
    __shared__ double s_red[kPoseNV], s_sys[kPoseNV];
    __shared__ Se3 s_pose, s_backup;
    __shared__ double s_lambda, s_ni, s_cur_chi, s_x[6];
    __shared__ int s_nbad_lm, s_flag, s_ok2, s_accept;
    __shared__ double s_rho;
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int M = A.counts[f];
    const size_t o = (size_t)f * A.slab;
    const float *Xw = A.Xw + 3 * o, *obs = A.obs + 2 * o, *wgt = A.w + o;
    uint8_t *outlier = A.outlier + o;
    double *err = A.err + 2 * o;
    const double in[4] = {A.fx, A.fy, A.cx, A.cy};
    const double delta = (double)(float)sqrt(5.991), dsqr = (double)(float)(delta * delta);   // const float deltaMono = sqrt(5.991), Optimizer.cc:295; delta^2 is a float member of RobustKernelHuber
    for (int i = tid; i < M; i += nt) outlier[i] = 0;
    if (tid < 6) s_x[tid] = 0.0;
    if (M < 3) { if (tid == 0) A.n_inliers[f] = 0; return; }                 // Optimizer.cc:387-388
    __syncthreads();
    int n_bad_total = 0;
    bool robust = true;
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        if (tid == 0) se3_from_Tcw(A.Tcw + 16 * f, s_pose);                  // reset to the initial pose, Optimizer.cc:400
        __syncthreads();
        // any active edge?  (initializeOptimization(0): level-0 edges only)
        int my_act = 0;
        for (int i = tid; i < M; i += nt) my_act += !outlier[i];
        const int n_act = __syncthreads_count(my_act > 0);
        if (n_act > 0) {
            // One pass over the edges per LM TRIAL: errors, robust chi2 AND the normal equations at the pose it is given (computeActiveErrors +
            // activeRobustChi2 + buildSystem).  g2o evaluates a trial with computeActiveErrors and, if it accepts, starts the next iteration with
            // computeActiveErrors + buildSystem at that same pose: the trial pass already holds exactly that system, so it is kept (s_sys) instead of being
            // rebuilt; a rejected trial keeps the previous system, as g2o does (only lambda changes).  Same loops, same summation order: bit-identical
            // to building twice, at about 60 % of the work.
            auto edge_pass = [&](const Se3 pose) {
                double acc[kPoseNV];
#pragma unroll
                for (int v = 0; v < kPoseNV; v++) acc[v] = 0;
                for (int i = tid; i < M; i += nt) {
                    if (outlier[i]) continue;
                    const double X[3] = {(double)Xw[3 * i], (double)Xw[3 * i + 1], (double)Xw[3 * i + 2]};
                    double Xc[3], e[2], Jp[12];
                    se3_map(pose, X, Xc);
                    reproj_error(Xc, in, (double)obs[2 * i], (double)obs[2 * i + 1], e);
                    err[2 * i] = e[0]; err[2 * i + 1] = e[1];
                    const double w = (double)wgt[i];
                    const double chi2 = e[0] * (w * e[0] + 0.0 * e[1]) + e[1] * (0.0 * e[0] + w * e[1]);
                    double rho0 = chi2, rho1 = 1.0;
                    if (robust) huber(chi2, delta, dsqr, rho0, rho1);
                    acc[27] += rho0;
                    jac_pose_only(Xc, A.fx, A.fy, Jp);
                    const double wo = rho1 * w;
                    const double r0 = rho1 * w * e[0], r1 = rho1 * w * e[1];         // b -= rho' A^T omega e
                    int p = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
                        acc[21 + a] -= Jp[a] * r0 + Jp[6 + a] * r1;
#pragma unroll
                        for (int c = a; c < 6; c++) { acc[p] += Jp[a] * wo * Jp[c] + Jp[6 + a] * wo * Jp[6 + c]; p++; }
                    }
                }
                block_reduce_vec<kPoseNV>(acc, s_warp, s_red);
            };
            edge_pass(s_pose);
            if (tid < kPoseNV) s_sys[tid] = s_red[tid];
            if (tid == 0) {
                s_cur_chi = s_red[27];
                double mx = 0.;                                                      // computeLambdaInit
                int p = 0;
                for (int a = 0; a < 6; a++) { mx = fmax(fabs(s_red[p]), mx); p += 6 - a; }
                s_lambda = 1e-5 * mx; s_ni = 2; s_nbad_lm = 0;
            }
            __syncthreads();
#pragma unroll 1
            for (int iter = 0; iter < 10; iter++) {
                const double ini_chi = s_cur_chi;
                int qmax = 0;
                double rho = 0;
                do {
                    if (tid == 0) {
                        s_backup = s_pose;
                        s_ok2 = solve6(s_sys, s_sys + 21, s_lambda, s_x) ? 1 : 0;
                        Se3 d, r;
                        se3_exp(s_x, d);
                        se3_mul(d, s_pose, r);
                        s_pose = r;
                    }
                    __syncthreads();
                    edge_pass(s_pose);
                    if (tid == 0) {
                        double temp_chi = s_red[27];
                        if (!s_ok2) temp_chi = 1.7976931348623157e308;
                        double r = s_cur_chi - temp_chi;
                        double scale = 0.;
                        for (int j = 0; j < 6; j++) scale += s_x[j] * (s_lambda * s_x[j] + s_sys[21 + j]);
                        scale += 1e-3;
                        r /= scale;
                        if (r > 0 && isfinite(temp_chi)) {
                            double alpha = 1. - pow((2 * r - 1), 3.0);
                            alpha = fmin(alpha, 2. / 3.);
                            s_lambda *= fmax(1. / 3., alpha);
                            s_ni = 2; s_cur_chi = temp_chi;
                            s_accept = 1;
                        } else {
                            s_lambda *= s_ni; s_ni *= 2;
                            s_pose = s_backup;
                            s_accept = 0;
                        }
                        s_rho = r;
                    }
                    __syncthreads();
                    rho = s_rho;
                    if (s_accept && tid < kPoseNV) s_sys[tid] = s_red[tid];          // the accepted trial's system is the next iteration's
                    qmax++;
                    __syncthreads();
                } while (rho < 0 && qmax < 10);
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
        // classify (Optimizer.cc:405-432)
        const Se3 pose = s_pose;
        int my_bad = 0;
        for (int i = tid; i < M; i += nt) {
            if (outlier[i]) {
                const double X[3] = {(double)Xw[3 * i], (double)Xw[3 * i + 1], (double)Xw[3 * i + 2]};
                double Xc[3], e[2];
                se3_map(pose, X, Xc);
                reproj_error(Xc, in, (double)obs[2 * i], (double)obs[2 * i + 1], e);
                err[2 * i] = e[0]; err[2 * i + 1] = e[1];
            }
            const double e0 = err[2 * i], e1 = err[2 * i + 1], w = (double)wgt[i];
            const float chi2 = (float)(e0 * (w * e0 + 0.0 * e1) + e1 * (0.0 * e0 + w * e1));
            if (chi2 > 5.991f) { outlier[i] = 1; my_bad++; } else outlier[i] = 0;
        }
        if (it == 2) robust = false;
        __shared__ int s_bad;
        if (tid == 0) s_bad = 0;
        __syncthreads();
        if (my_bad) atomicAdd(&s_bad, my_bad);
        __syncthreads();
        n_bad_total = s_bad;
        __syncthreads();
        if (M < 10) break;                                                            // optimizer.edges().size()<10
    }
    if (tid == 0) {
        se3_to_Tcw(s_pose, A.Tcw + 16 * f);
        A.n_inliers[f] = M - n_bad_total;
    }
}


// Frame -> edge list, in feature order like the loop of Optimizer.cc:303-384: every feature that holds a map point
// (feat_match >= 0) becomes one EdgeSE3ProjectXYZOnlyPose.  One CTA per frame, stable compaction by block scan.
struct PoseGatherArgs {
    int f_slab, q_slab, nlevels;
    const float2 *f_xy; const int *f_octave; const int *f_counts; const int *feat_match;
    const float *q_Xw; const int *q_counts; const float *inv_level_sigma2;
    float *Xw, *obs, *w; int *edge_feat; int *counts;      // outputs, slab = f_slab
};

__global__ void __launch_bounds__(256)
k_pose_gather(const PoseGatherArgs A)
{
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = A.f_counts[f], M = A.q_counts[f];
    const size_t fo = (size_t)f * A.f_slab, qo = (size_t)f * A.q_slab;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 256) {
        const int i = base + tid;
        int q = -1;
        if (i < N) { q = A.feat_match[fo + i]; if (q >= M) q = -1; }
        const unsigned bal = __ballot_sync(0xffffffffu, q >= 0);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < wid; w++) off += s_warp[w];
        if (q >= 0) {
            const int e = off + __popc(bal & ((1u << lane) - 1));
            const float2 p = A.f_xy[fo + i];
            A.obs[2 * (fo + e)] = p.x; A.obs[2 * (fo + e) + 1] = p.y;
            A.Xw[3 * (fo + e)] = A.q_Xw[3 * (qo + q)]; A.Xw[3 * (fo + e) + 1] = A.q_Xw[3 * (qo + q) + 1]; A.Xw[3 * (fo + e) + 2] = A.q_Xw[3 * (qo + q) + 2];
            A.w[fo + e] = A.inv_level_sigma2[min(max(A.f_octave[fo + i], 0), A.nlevels - 1)];
            A.edge_feat[fo + e] = i;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) A.counts[f] = s_base;
}

// edge outlier flags -> per-feature flags (Frame::mvbOutlier); f_outlier is zeroed beforehand
__global__ void __launch_bounds__(256)
k_pose_scatter(int f_slab, const int *__restrict__ e_counts, const int *__restrict__ edge_feat,
               const uint8_t *__restrict__ e_outlier, uint8_t *__restrict__ f_outlier)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t fo = (size_t)f * f_slab;
    if (i < e_counts[f]) f_outlier[fo + edge_feat[fo + i]] = e_outlier[fo + i];
}

}  // namespace orbs
