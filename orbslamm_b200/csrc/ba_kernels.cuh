// ba_kernels.cuh -- bundle-adjustment kernels (Optimizer::LocalBundleAdjustment / BundleAdjustment on g2o's
// BlockSolver_6_3 + Levenberg, restated).  All fp64, every reduction in a fixed order (bit-reproducible run to run).
//
// Edges are stored grouped by map point (CSR); a point's work (Hll, b_l, its Hpl blocks, back-substitution) is done by a
// group of 8 lanes (a point has ~6 observations).  The Schur complement is assembled per TARGET: the (pose, pose) pairs of
// every point are generated once per call, sorted by the 6x6 block they contribute to (tile slot, block, point order), and one
// warp sums a block's contributions in registers and writes it once -- no atomics, no 82 MB memset, only the structurally
// nonzero 64x64 tiles of the reduced system exist (tile_solver.cuh).  The Levenberg-Marquardt control flow lives in a device
// control block (LmCtl); the host enqueues "slots" and never waits for a decision.
//
//   k_ba_errors        computeActiveErrors + activeRobustChi2 partials      sparse_optimizer.cpp:61-114
//   k_ba_build_points  linearizeOplus + constructQuadraticForm (Hll, b_l, Hpl)   block_solver.hpp:506-564,
//   k_ba_build_poses   ... (Hpp, b_p)                                       base_binary_edge.hpp:55-120
//   k_lm_iter_begin    currentChi, computeLambdaInit                        levenberg.cpp:69-100, 166-180
//   k_ba_point_prep    Hll + lambda I = G G^T per point, Z_e = Hpl_e G^-T, g = G^-1 b_l      block_solver.hpp:371-398
//   k_ba_diag_init / k_ba_schur_seg   setLambda + Schur complement          block_solver.hpp:399-436, 568-593
//   k_rs_solve         sparse tiled Cholesky + triangular solves (tile_solver.cuh; replaces LinearSolverEigen)
//   k_ba_backsub                  x_l = Dinv (b_l - Hpl^T x_p) + computeScale   block_solver.hpp:463-487, levenberg.cpp:182-189
//   k_ba_update / k_ba_restore    push + oplus / pop                        sparse_optimizer.cpp:422-435
//   k_lm_reduce / k_lm_decide     rho, lambda update, accept / reject, stop rules   levenberg.cpp:102-161
#pragma once
#include "common.cuh"
#include "se3.cuh"
#include "tile_solver.cuh"

namespace orbs {

constexpr int kPosesPerTile = 10;       // 60 rows + 4 identity padding rows per 64-row tile

struct BaDev {
    int K, P, E;
    int nA;                       // free active poses (reduced system has 6 nA unknowns)
    int n;                        // 6 * nA
    int nt, ns;                   // tile rows, nonzero tiles
    // state
    Se3 *pose, *pose_bak;         // [K]
    double *pt, *pt_bak;          // [P*3]
    const double *intr;           // [K*4]
    // edges grouped by point
    const int *pt_start;          // [P+1]
    const int *e_kf;              // [E]
    const double *e_obs;          // [E*2]
    const double *e_w;            // [E]
    uint8_t *e_level;             // [E] 0 = active, 1 = outlier (setLevel(1))
    double *e_err;                // [E*2] stored _error
    double *e_W;                  // [E*18] Hpl block of the edge (6x3), valid for active edges with a free pose
    double *e_Z;                  // [E*18] Hpl G^-T for the current lambda
    // edges grouped by pose
    const int *pose_start;        // [K+1]
    const int *pose_edges;        // [E] -> edge index (point-grouped order)
    const int *e_point;           // [E] point of an edge (point-grouped order)
    // index mapping
    const int *pose_idx;          // [K] hessian index or -1 (fixed / inactive)
    const int *rowbase;           // [nA] first row of a free keyframe's 6x6 block in the tile-permuted reduced system
    const int *slot_of;           // [nt*nt]
    const int *tile_pose;         // [nt*kPosesPerTile] hessian index of block row b of tile t, or -1 (padding)
    uint8_t *pt_active;           // [P]
    // system
    double *Hpp, *bp;             // [nA*36], [nA*6]
    double *Hll, *bl;             // [P*9], [P*3]
    double *ptG, *ptg;            // [P*6] Cholesky factor of Hll + lambda I (g00 g10 g11 g20 g21 g22), [P*3] G^-1 b_l
    double *x;                    // [nt*64] pose solution (tile-permuted rows) -- then points [3P] in xl
    double *xl;                   // [3P]
    double *A, *bs;               // [ns*4096] packed tiles of the reduced system, [nt*64] its right-hand side (contiguous: one all-reduce)
    double *p_chi, *p_pt, *p_xp, *p_diag;   // per-block partial sums
    int n_chi, n_pt, n_xp, n_diag;
    double *scalars;              // [16]: 0 chi2, 1 scale (points), 2 scale (poses), 3 stop, 4 cholesky failure, 5 max diagonal; 8.. max-diagonal slots per rank
    int *flags;                   // [4]: 0 cholesky failure
    LmCtl *ctl;
    // Schur pairs sorted by target block
    const unsigned long long *pair_key, *pair_val;
    const int *seg_start;         // [nseg + 1]
    const int *n_seg;             // device scalar
    double delta, dsqr;
    int robust, lead, nranks, rank;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ double group8_sum(double v)
{
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ double edge_chi2(double e0, double e1, double w) { return e0 * (w * e0 + 0.0 * e1) + e1 * (0.0 * e0 + w * e1); }

// block-level sum of one double per thread in a fixed order -> partial[blockIdx.x]
__device__ __forceinline__ void block_partial(double v, double *partial)
{
    __shared__ double s_w[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += s_w[w]; partial[blockIdx.x] = s; }
}

// errors of all active edges at the current estimate + robust chi2 (per-block partial sums, fixed order)
__global__ void __launch_bounds__(256)
k_ba_errors(const BaDev B)
{
    if (B.ctl->state == LM_DONE) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double chi = 0;
    if (e < B.E && B.e_level[e] == 0) {
        const int kf = B.e_kf[e], p = B.e_point[e];
        double Xc[3], er[2];
        se3_map(B.pose[kf], &B.pt[3 * p], Xc);
        reproj_error(Xc, &B.intr[4 * kf], B.e_obs[2 * e], B.e_obs[2 * e + 1], er);
        B.e_err[2 * e] = er[0]; B.e_err[2 * e + 1] = er[1];
        const double c = edge_chi2(er[0], er[1], B.e_w[e]);
        double r0 = c, r1 = 1;
        if (B.robust) huber(c, B.delta, B.dsqr, r0, r1);
        chi = r0;
    }
    block_partial(chi, B.p_chi);
}

// Hll + lambda I = G G^T (lower Cholesky of the 3x3 point block); returns the reciprocals of the diagonal
__device__ __forceinline__ void point_chol(const double *H, double lambda, double (&G)[6], double (&inv)[3])
{
    G[0] = sqrt(H[0] + lambda); inv[0] = 1.0 / G[0];
    G[1] = H[3] * inv[0]; G[3] = H[6] * inv[0];
    G[2] = sqrt(H[4] + lambda - G[1] * G[1]); inv[1] = 1.0 / G[2];
    G[4] = (H[7] - G[3] * G[1]) * inv[1];
    G[5] = sqrt(H[8] + lambda - G[3] * G[3] - G[4] * G[4]); inv[2] = 1.0 / G[5];
}

// 8 lanes per point: Hll, b_l and the per-edge Hpl blocks.  When lambda is already known (every iteration but the first) and the
// point has at most 8 observations, the per-trial quantities of k_ba_point_prep (G, g, Z = Hpl G^-T) are produced here from registers.
__global__ void __launch_bounds__(256)
k_ba_build_points(const BaDev B)
{
    if (B.ctl->state != LM_BUILD) return;
    const int p = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    const bool live = p < B.P && B.pt_active[p];
    double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    double W[18];
    int my_e = -1;
    const int e_begin = live ? B.pt_start[p] : 0, e_end = live ? B.pt_start[p + 1] : 0;
    if (live) {
        const double X[3] = {B.pt[3 * p], B.pt[3 * p + 1], B.pt[3 * p + 2]};
        for (int e = e_begin + sub; e < e_end; e += 8) {
            if (B.e_level[e]) continue;
            const int kf = B.e_kf[e];
            const Se3 T = B.pose[kf];
            double Xc[3], R[9], Jl[6], Jp[12];
            se3_map(T, X, Xc);
            quat_to_R(T.q, R);
            jac_binary(Xc, R, B.intr[4 * kf], B.intr[4 * kf + 1], Jl, Jp);
            const double e0 = B.e_err[2 * e], e1 = B.e_err[2 * e + 1], w = B.e_w[e];
            double rw = 1.0, r0u;
            if (B.robust) huber(edge_chi2(e0, e1, w), B.delta, B.dsqr, r0u, rw);
            const double wo = rw * w;
            const double r0 = -w * e0 * rw, r1 = -w * e1 * rw;          // omega_r = -omega e rho'
            H[0] += Jl[0] * wo * Jl[0] + Jl[3] * wo * Jl[3];
            H[1] += Jl[0] * wo * Jl[1] + Jl[3] * wo * Jl[4];
            H[2] += Jl[0] * wo * Jl[2] + Jl[3] * wo * Jl[5];
            H[3] += Jl[1] * wo * Jl[1] + Jl[4] * wo * Jl[4];
            H[4] += Jl[1] * wo * Jl[2] + Jl[4] * wo * Jl[5];
            H[5] += Jl[2] * wo * Jl[2] + Jl[5] * wo * Jl[5];
            b[0] += Jl[0] * r0 + Jl[3] * r1; b[1] += Jl[1] * r0 + Jl[4] * r1; b[2] += Jl[2] * r0 + Jl[5] * r1;
            if (B.pose_idx[kf] >= 0) {
                double *Wg = &B.e_W[18 * (size_t)e];
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { W[3 * a + c] = Jp[a] * wo * Jl[c] + Jp[6 + a] * wo * Jl[3 + c]; Wg[3 * a + c] = W[3 * a + c]; }
                my_e = e;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) H[i] = group8_sum(H[i]);
#pragma unroll
    for (int i = 0; i < 3; i++) b[i] = group8_sum(b[i]);
    if (!live) return;
    if (sub == 0) {
        double *Ho = &B.Hll[9 * (size_t)p];
        Ho[0] = H[0]; Ho[1] = H[1]; Ho[2] = H[2]; Ho[3] = H[1]; Ho[4] = H[3]; Ho[5] = H[4]; Ho[6] = H[2]; Ho[7] = H[4]; Ho[8] = H[5];
        B.bl[3 * p] = b[0]; B.bl[3 * p + 1] = b[1]; B.bl[3 * p + 2] = b[2];
    }
    if (B.ctl->first || e_end - e_begin > 8) return;                    // lambda not known yet / more than one edge per lane: k_ba_point_prep
    const double Hf[9] = {H[0], H[1], H[2], H[1], H[3], H[4], H[2], H[4], H[5]};
    double G[6], inv[3];
    point_chol(Hf, B.ctl->lambda, G, inv);
    if (sub == 0) {
        double *Go = &B.ptG[6 * (size_t)p];
#pragma unroll
        for (int i = 0; i < 6; i++) Go[i] = G[i];
        const double y0 = b[0] * inv[0], y1 = (b[1] - G[1] * y0) * inv[1], y2 = (b[2] - G[3] * y0 - G[4] * y1) * inv[2];
        B.ptg[3 * p] = y0; B.ptg[3 * p + 1] = y1; B.ptg[3 * p + 2] = y2;
    }
    if (my_e >= 0) {
        double *Z = &B.e_Z[18 * (size_t)my_e];
#pragma unroll
        for (int a = 0; a < 6; a++) {
            const double z0 = W[3 * a] * inv[0], z1 = (W[3 * a + 1] - z0 * G[1]) * inv[1], z2 = (W[3 * a + 2] - z0 * G[3] - z1 * G[4]) * inv[2];
            Z[3 * a] = z0; Z[3 * a + 1] = z1; Z[3 * a + 2] = z2;
        }
    }
}

// one 128-thread CTA per keyframe: Hpp (full 6x6) and b_p over the keyframe's observations
__global__ void __launch_bounds__(128)
k_ba_build_poses(const BaDev B)
{
    if (B.ctl->state != LM_BUILD) return;
    __shared__ double s_acc[4][27];
    const int kf = blockIdx.x;
    const int ip = B.pose_idx[kf];
    if (ip < 0) return;
    double acc[27];
#pragma unroll
    for (int i = 0; i < 27; i++) acc[i] = 0;
    const Se3 T = B.pose[kf];
    const double fx = B.intr[4 * kf], fy = B.intr[4 * kf + 1];
    for (int j = B.pose_start[kf] + threadIdx.x; j < B.pose_start[kf + 1]; j += 128) {
        const int e = B.pose_edges[j];
        if (B.e_level[e]) continue;
        const int p = B.e_point[e];
        double Xc[3], Jp[12];
        se3_map(T, &B.pt[3 * p], Xc);
        {   // pose Jacobian of EdgeSE3ProjectXYZ (division form)
            const double x = Xc[0], y = Xc[1], z = Xc[2], z_2 = z * z;
            Jp[0] = x * y / z_2 * fx; Jp[1] = -(1 + (x * x / z_2)) * fx; Jp[2] = y / z * fx;
            Jp[3] = -1. / z * fx; Jp[4] = 0; Jp[5] = x / z_2 * fx;
            Jp[6] = (1 + y * y / z_2) * fy; Jp[7] = -x * y / z_2 * fy; Jp[8] = -x / z * fy;
            Jp[9] = 0; Jp[10] = -1. / z * fy; Jp[11] = y / z_2 * fy;
        }
        const double e0 = B.e_err[2 * e], e1 = B.e_err[2 * e + 1], w = B.e_w[e];
        double rw = 1.0, r0u;
        if (B.robust) huber(edge_chi2(e0, e1, w), B.delta, B.dsqr, r0u, rw);
        const double wo = rw * w;
        const double r0 = -w * e0 * rw, r1 = -w * e1 * rw;
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
            acc[21 + a] += Jp[a] * r0 + Jp[6 + a] * r1;
#pragma unroll
            for (int c = a; c < 6; c++) { acc[q] += Jp[a] * wo * Jp[c] + Jp[6 + a] * wo * Jp[6 + c]; q++; }
        }
    }
#pragma unroll
    for (int i = 0; i < 27; i++) acc[i] = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 27; i++) s_acc[threadIdx.x >> 5][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 27) {
        const double v = s_acc[0][threadIdx.x] + s_acc[1][threadIdx.x] + s_acc[2][threadIdx.x] + s_acc[3][threadIdx.x];
        if (threadIdx.x >= 21) B.bp[6 * ip + threadIdx.x - 21] = v;
        else {
            int a = 0, q = threadIdx.x;
            while (q >= 6 - a) { q -= 6 - a; a++; }
            const int c = a + q;
            double *H = &B.Hpp[36 * (size_t)ip];
            H[6 * a + c] = v; H[6 * c + a] = v;
        }
    }
}

// max |diagonal| of the free blocks (computeLambdaInit, levenberg.cpp:166-180): per-block partial maxima.  Sharded: the pose blocks
// are partial sums here, their diagonals are summed over the ranks first (host: all-reduce of diag6) and folded in by k_lm_iter_begin.
__global__ void __launch_bounds__(256)
k_ba_max_diag(const BaDev B, double *__restrict__ diag6)
{
    if (B.ctl->state != LM_BUILD || !B.ctl->first) return;
    __shared__ double s_w[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0;
    if (i < B.nA) for (int j = 0; j < 6; j++) { const double d = B.Hpp[36 * (size_t)i + 7 * j]; diag6[6 * i + j] = d; if (B.nranks == 1) m = fmax(m, fabs(d)); }
    const int p = i - B.nA;
    if (p >= 0 && p < B.P && B.pt_active[p]) for (int j = 0; j < 3; j++) m = fmax(m, fabs(B.Hll[9 * (size_t)p + 4 * j]));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; w++) s = fmax(s, s_w[w]); B.p_diag[blockIdx.x] = s; }
}

// this rank's maximum into its slot of the per-rank vector (sum all-reduce of zero-padded slots = all-gather)
__global__ void __launch_bounds__(256)
k_ba_max_diag_finish(const BaDev B, double *__restrict__ rank_slots)
{
    if (B.ctl->state != LM_BUILD || !B.ctl->first) return;
    __shared__ double s[256];
    double a = 0;
    for (int i = threadIdx.x; i < B.n_diag; i += 256) a = fmax(a, B.p_diag[i]);
    s[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) { if (threadIdx.x < d) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + d]); __syncthreads(); }
    if (threadIdx.x < B.nranks) rank_slots[threadIdx.x] = (threadIdx.x == B.rank) ? s[0] : 0.0;
}

// start of an LM iteration (levenberg.cpp:69-100): currentChi of the stored errors; lambda from the diagonal maximum on the first one
__global__ void k_lm_iter_begin(const BaDev B, const double *__restrict__ diag6, const double *__restrict__ rank_slots)
{
    LmCtl *c = B.ctl;
    if (c->state != LM_BUILD) return;
    if (c->first) {
        double m = 0;
        for (int r = 0; r < B.nranks; r++) m = fmax(m, rank_slots[r]);
        if (B.nranks > 1) for (int i = 0; i < B.n; i++) m = fmax(m, fabs(diag6[i]));
        c->currentChi = B.scalars[0];
        c->lambda = 1e-5 * m; c->ni = 2; c->nbad = 0;
        c->first = 0;
        c->fresh_first = 1;                                 // this slot linearised without lambda: k_ba_point_prep does all points
    } else c->fresh_first = 0;
    c->iniChi = c->currentChi;
    c->qmax = 0;
    c->rho = 0;
}

// per trial, 8 lanes per point: Hll + lambda I = G G^T, g = G^-1 b_l, and per edge Z = Hpl G^-T  (so that Hpl Dinv Hpl'^T = Z Z'^T).
// After an accepted trial k_ba_build_points has produced them already (same arithmetic); this kernel then only serves the first
// iteration, retries with a larger lambda and points with more than 8 observations.
__global__ void __launch_bounds__(256)
k_ba_point_prep(const BaDev B)
{
    const int state = B.ctl->state;
    if (state == LM_DONE) return;
    if (blockIdx.x == 0 && threadIdx.x < 4) B.flags[threadIdx.x] = 0;
    const int p = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    if (p >= B.P || !B.pt_active[p]) return;
    const int e_begin = B.pt_start[p], e_end = B.pt_start[p + 1];
    if (state == LM_BUILD && !B.ctl->fresh_first && e_end - e_begin <= 8) return;       // done by k_ba_build_points
    const double lambda = B.ctl->lambda;
    double G[6], inv[3];
    point_chol(&B.Hll[9 * (size_t)p], lambda, G, inv);
    if (sub == 0) {
        double *Go = &B.ptG[6 * (size_t)p];
#pragma unroll
        for (int i = 0; i < 6; i++) Go[i] = G[i];
        const double b0 = B.bl[3 * p], b1 = B.bl[3 * p + 1], b2 = B.bl[3 * p + 2];
        const double y0 = b0 * inv[0], y1 = (b1 - G[1] * y0) * inv[1], y2 = (b2 - G[3] * y0 - G[4] * y1) * inv[2];
        B.ptg[3 * p] = y0; B.ptg[3 * p + 1] = y1; B.ptg[3 * p + 2] = y2;
    }
    for (int e = e_begin + sub; e < e_end; e += 8) {
        if (B.e_level[e] || B.pose_idx[B.e_kf[e]] < 0) continue;
        const double *W = &B.e_W[18 * (size_t)e];
        double *Z = &B.e_Z[18 * (size_t)e];
#pragma unroll
        for (int a = 0; a < 6; a++) {
            const double z0 = W[3 * a] * inv[0], z1 = (W[3 * a + 1] - z0 * G[1]) * inv[1], z2 = (W[3 * a + 2] - z0 * G[3] - z1 * G[4]) * inv[2];
            Z[3 * a] = z0; Z[3 * a + 1] = z1; Z[3 * a + 2] = z2;
        }
    }
}

// reduced system before the Schur terms (A and bs were zeroed): lambda on the real diagonal entries, 1 on the padding rows (lead rank only)
__global__ void __launch_bounds__(256)
k_ba_diag_init(const BaDev B)
{
    if (B.ctl->state == LM_DONE || !B.lead) return;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= B.nt * TS) return;
    const int t = row >> 6, r = row & 63, br = r / 6;
    const bool real = br < kPosesPerTile && B.tile_pose[t * kPosesPerTile + br] >= 0;
    B.A[(size_t)t * TS2 + r * TS + r] = real ? B.ctl->lambda : 1.0;          // diagonal tile t is slot t
}

// ---- Schur complement, assembled per target 6x6 block from the sorted pair list ------------------------------------------
// A block's pairs ("segment") are cut into work items of <= kItemPairs pairs; one 128-thread CTA per item:
//   1. every thread loads the two Z records (6x3 fp64 = 144 B each) of ITS pair with 18 independent 16-byte loads and multiplies them
//      (36 products; diagonal blocks: + Z g for the right-hand side).  (A variant that staged the records through shared memory with
//      cooperative cp.async was slower: 203 vs 164 us -- the staging loop costs more instructions than the 32-line requests it saves.)
//   2. the warp folds its 32 x 42 partial values with a multi-value butterfly (43 shuffles instead of 210), the four warps are summed
//      in fixed order and the item's partial block goes to scratch.
// k_ba_schur_finish (one warp per block) adds a block's items in item order: block += [diag] Hpp - sum, bs += b_p - sum Z g.
constexpr int kItemPairs = 128;
constexpr int kRecPitch = 19;           // doubles per staged record (odd: conflict-free 64-bit reads with one record per lane)
constexpr int kSchurVals = 42;          // 36 block entries + 6 right-hand-side entries

template <int N>
__device__ __forceinline__ void mv_fold(double *v, int bit, int lane)
{
    constexpr int H = (N + 1) / 2;
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
        const double lo = v[i], hi = (i + H < N) ? v[i + H] : 0.0;
        const double send = upper ? lo : hi, keep = upper ? hi : lo;
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
}

__global__ void __launch_bounds__(128, 4)
k_ba_schur_items(const BaDev B, const int *__restrict__ seg_item0, double *__restrict__ item_part)
{
    if (B.ctl->state == LM_DONE) return;
    __shared__ int s_seg;
    __shared__ double s_part[4][kSchurVals];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, item = blockIdx.x;
    if (tid == 0) {                                            // the segment of this item: last seg with seg_item0[seg] <= item
        int lo = 0, hi = *B.n_seg - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (seg_item0[mid] <= item) lo = mid; else hi = mid - 1; }
        s_seg = lo;
    }
    __syncthreads();
    const int seg = s_seg;
    const int s0 = B.seg_start[seg] + (item - seg_item0[seg]) * kItemPairs, s1 = min(B.seg_start[seg + 1], s0 + kItemPairs);
    const unsigned tgt = (unsigned)(B.pair_key[B.seg_start[seg]] >> 32);
    const int slot = tgt >> 7, blk = tgt & 127;
    const bool diag = slot < B.nt && blk / kPosesPerTile == blk % kPosesPerTile;
    const int t = s0 + tid;
    int er = -1, ec = -1;
    if (t < s1) {
        const unsigned long long pv = B.pair_val[t];
        er = (int)(pv & 0xffffffffu); ec = (int)(pv >> 32);
        if (B.e_level[er] | B.e_level[ec]) { er = -1; ec = -1; }          // gated out after the robust stage
    }
    double v[kSchurVals];
#pragma unroll
    for (int i = 0; i < kSchurVals; i++) v[i] = 0.0;
    if (er >= 0) {
        // every thread multiplies its own pair: 9 + 9 independent 16-byte loads in flight per thread
        const double2 *pr = reinterpret_cast<const double2 *>(&B.e_Z[18 * (size_t)er]);
        const double2 *pc = reinterpret_cast<const double2 *>(&B.e_Z[18 * (size_t)ec]);
        double zr[18], zc[18];
#pragma unroll
        for (int i = 0; i < 9; i++) { const double2 q = pr[i]; zr[2 * i] = q.x; zr[2 * i + 1] = q.y; }
#pragma unroll
        for (int i = 0; i < 9; i++) { const double2 q = pc[i]; zc[2 * i] = q.x; zc[2 * i + 1] = q.y; }
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = 0; c < 6; c++) v[6 * a + c] = zr[3 * a] * zc[3 * c] + zr[3 * a + 1] * zc[3 * c + 1] + zr[3 * a + 2] * zc[3 * c + 2];
        if (diag) {
            if (er == ec) {
                const double *g = &B.ptg[3 * (size_t)B.e_point[er]];
                const double g0 = g[0], g1 = g[1], g2 = g[2];
#pragma unroll
                for (int a = 0; a < 6; a++) v[36 + a] = zr[3 * a] * g0 + zr[3 * a + 1] * g1 + zr[3 * a + 2] * g2;
            } else {
                // two observations of one point by the same keyframe (never produced by ORB-SLAM, legal for g2o): both cross terms
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int c = a + 1; c < 6; c++) { const double sm = v[6 * a + c] + v[6 * c + a]; v[6 * a + c] = sm; v[6 * c + a] = sm; }
#pragma unroll
                for (int a = 0; a < 6; a++) v[7 * a] *= 2.0;
            }
        }
    }
    // fold 32 lanes x 42 values: after the five steps a lane holds (at most) two fully summed values
    mv_fold<42>(v, 16, lane); mv_fold<21>(v, 8, lane); mv_fold<11>(v, 4, lane); mv_fold<6>(v, 2, lane); mv_fold<3>(v, 1, lane);
#pragma unroll
    for (int f = 0; f < 2; f++) {
        int i4 = f + ((lane & 1) ? 2 : 0); bool ok = i4 < 3;
        int i3 = i4 + ((lane & 2) ? 3 : 0); ok = ok && i3 < 6;
        int i2 = i3 + ((lane & 4) ? 6 : 0); ok = ok && i2 < 11;
        int i1 = i2 + ((lane & 8) ? 11 : 0); ok = ok && i1 < 21;
        int i0 = i1 + ((lane & 16) ? 21 : 0); ok = ok && i0 < 42;
        if (ok) s_part[warp][i0] = v[f];
    }
    __syncthreads();
    if (tid < kSchurVals) item_part[(size_t)item * kSchurVals + tid] = ((s_part[0][tid] + s_part[1][tid]) + s_part[2][tid]) + s_part[3][tid];
}

__global__ void __launch_bounds__(256)
k_ba_schur_finish(const BaDev B, const int *__restrict__ seg_item0, const double *__restrict__ item_part)
{
    if (B.ctl->state == LM_DONE) return;
    const int seg = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (seg >= *B.n_seg) return;
    const unsigned tgt = (unsigned)(B.pair_key[B.seg_start[seg]] >> 32);
    const int slot = tgt >> 7, blk = tgt & 127, br = blk / kPosesPerTile, bc = blk % kPosesPerTile;
    const bool diag = slot < B.nt && br == bc;
    double m1 = 0, m2 = 0;                                     // entries lane and lane + 32 (< 42)
    for (int it = seg_item0[seg]; it < seg_item0[seg + 1]; it++) {
        m1 += item_part[(size_t)it * kSchurVals + lane];
        if (lane < kSchurVals - 32) m2 += item_part[(size_t)it * kSchurVals + 32 + lane];
    }
    double *Ab = B.A + (size_t)slot * TS2 + (6 * br) * TS + 6 * bc;
    double h1 = 0, h2 = 0;
    if (diag) {
        const int ip = B.tile_pose[slot * kPosesPerTile + br];
        h1 = B.Hpp[36 * (size_t)ip + lane];
        if (lane < 4) h2 = B.Hpp[36 * (size_t)ip + 32 + lane];
        if (lane >= 4 && lane < 10) B.bs[slot * TS + 6 * br + lane - 4] += B.bp[6 * ip + lane - 4] - m2;
    }
    Ab[(lane / 6) * TS + lane % 6] += h1 - m1;
    if (lane < 4) Ab[((32 + lane) / 6) * TS + (32 + lane) % 6] += h2 - m2;
}

// work items per segment (k_ba_schur_items)
__global__ void __launch_bounds__(256)
k_ba_seg_items(const int *__restrict__ seg_start, const int *__restrict__ n_seg, int *__restrict__ cnt, int n)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    cnt[s] = s < *n_seg ? (seg_start[s + 1] - seg_start[s] + kItemPairs - 1) / kItemPairs : 0;
}

// 8 lanes per point: x_l = G^-T (g - sum_e Z_e^T x_p(e)); point part of computeScale (per-block partials)
__global__ void __launch_bounds__(256)
k_ba_backsub(const BaDev B)
{
    if (B.ctl->state == LM_DONE) return;
    const int p = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    const bool live = p < B.P && B.pt_active[p];
    double c[3] = {0, 0, 0}, sc = 0;
    if (live) {
        for (int e = B.pt_start[p] + sub; e < B.pt_start[p + 1]; e += 8) {
            if (B.e_level[e]) continue;
            const int i1 = B.pose_idx[B.e_kf[e]];
            if (i1 < 0) continue;
            const double *Z = &B.e_Z[18 * (size_t)e];
            const double *xp = &B.x[B.rowbase[i1]];
#pragma unroll
            for (int a = 0; a < 6; a++) { const double xa = xp[a]; c[0] += Z[3 * a] * xa; c[1] += Z[3 * a + 1] * xa; c[2] += Z[3 * a + 2] * xa; }
        }
    }
    c[0] = group8_sum(c[0]); c[1] = group8_sum(c[1]); c[2] = group8_sum(c[2]);
    if (live && sub == 0) {
        const double *G = &B.ptG[6 * (size_t)p];
        const double r0 = B.ptg[3 * p] - c[0], r1 = B.ptg[3 * p + 1] - c[1], r2 = B.ptg[3 * p + 2] - c[2];
        const double x2 = r2 / G[5], x1 = (r1 - G[4] * x2) / G[2], x0 = (r0 - G[1] * x1 - G[3] * x2) / G[0];
        const double lambda = B.ctl->lambda;
        B.xl[3 * (size_t)p] = x0; B.xl[3 * (size_t)p + 1] = x1; B.xl[3 * (size_t)p + 2] = x2;
        sc = x0 * (lambda * x0 + B.bl[3 * p]) + x1 * (lambda * x1 + B.bl[3 * p + 1]) + x2 * (lambda * x2 + B.bl[3 * p + 2]);
    }
    if ((int)blockIdx.x < B.n_pt) block_partial(sc, B.p_pt);
    // pose part of computeScale (x lives in tile-permuted rows of the solver output): one partial per block for the first n_xp blocks
    if ((int)blockIdx.x < B.n_xp) {
        __syncthreads();
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        double sx = 0;
        if (i < B.n) {
            const double x = B.x[B.rowbase[i / 6] + i % 6];
            sx = x * ((B.lead ? B.ctl->lambda * x : 0.0) + B.bp[i]);        // sharded: lambda x^2 once (lead rank), bp is this rank's partial sum
        }
        block_partial(sx, B.p_xp);
    }
}

// push + oplus
__global__ void __launch_bounds__(256)
k_ba_update(const BaDev B)
{
    if (B.ctl->state == LM_DONE) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.K) {
        const int ip = B.pose_idx[i];
        if (ip >= 0) {
            const Se3 T = B.pose[i];
            B.pose_bak[i] = T;
            Se3 d, r;
            se3_exp(&B.x[B.rowbase[ip]], d);
            se3_mul(d, T, r);
            B.pose[i] = r;
        }
    }
    const int p = i - B.K;
    if (p >= 0 && p < B.P && B.pt_active[p]) {
        for (int a = 0; a < 3; a++) { const double v = B.pt[3 * p + a]; B.pt_bak[3 * p + a] = v; B.pt[3 * p + a] = v + B.xl[3 * (size_t)p + a]; }
    }
}

// pop (only after a rejected trial)
__global__ void __launch_bounds__(256)
k_ba_restore(const BaDev B)
{
    if (!B.ctl->last_rejected) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.K && B.pose_idx[i] >= 0) B.pose[i] = B.pose_bak[i];
    const int p = i - B.K;
    if (p >= 0 && p < B.P && B.pt_active[p]) for (int a = 0; a < 3; a++) B.pt[3 * p + a] = B.pt_bak[3 * p + a];
}

__device__ __forceinline__ void lm_decide(const BaDev &B);

// fixed-order sums of the per-block partials of one trial -> scalars[0..4] (sharded: summed over the ranks before k_lm_decide)
__global__ void __launch_bounds__(256)
k_lm_reduce(const BaDev B, const volatile int *__restrict__ host_stop, int with_scale)
{
    if (B.ctl->state == LM_DONE) { if (with_scale == 2 && threadIdx.x == 0) B.ctl->last_rejected = 0; return; }
    __shared__ double s[3][256];
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = threadIdx.x; i < B.n_chi; i += 256) a0 += B.p_chi[i];
    if (with_scale) {
        for (int i = threadIdx.x; i < B.n_pt; i += 256) a1 += B.p_pt[i];
        for (int i = threadIdx.x; i < B.n_xp; i += 256) a2 += B.p_xp[i];
    }
    s[0][threadIdx.x] = a0; s[1][threadIdx.x] = a1; s[2][threadIdx.x] = a2;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) { s[0][threadIdx.x] += s[0][threadIdx.x + d]; s[1][threadIdx.x] += s[1][threadIdx.x + d]; s[2][threadIdx.x] += s[2][threadIdx.x + d]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        B.scalars[0] = s[0][0]; B.scalars[1] = s[1][0]; B.scalars[2] = s[2][0];
        B.scalars[3] = (host_stop && *host_stop) ? 1.0 : 0.0;
        B.scalars[4] = B.flags[0] ? 1.0 : 0.0;
        if (with_scale == 2) lm_decide(B);                                  // single GPU: nothing to exchange, decide in the same launch
    }
}

// one LM trial decided on the device (levenberg.cpp:102-161 + the iteration loop of SparseOptimizer::optimize)
__device__ __forceinline__ void lm_decide(const BaDev &B)
{
    LmCtl *c = B.ctl;
    c->last_rejected = 0;
    if (c->state == LM_DONE) return;
    const bool ok2 = B.scalars[4] == 0.0, stop = B.scalars[3] != 0.0;
    if (!ok2) c->chol_failures++;
    double tempChi = B.scalars[0];
    if (!ok2) tempChi = 1.7976931348623157e308;
    double rho = c->currentChi - tempChi;
    double scale = B.scalars[1] + B.scalars[2];
    scale += 1e-3;
    rho /= scale;
    c->tempChi = tempChi; c->scale = scale;
    if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3.0);
        alpha = fmin(alpha, 2. / 3.);
        c->lambda *= fmax(1. / 3., alpha);
        c->ni = 2;
        c->currentChi = tempChi;
    } else {
        c->lambda *= c->ni; c->ni *= 2;
        c->last_rejected = 1;
    }
    c->rho = rho;
    c->qmax++; c->lm_trials++;
    if (rho < 0 && c->qmax < 10 && !stop) { c->state = LM_RETRY; return; }
    // the iteration is over
    c->lm_iterations++; c->iteration++;
    if (c->qmax == 10 || rho == 0) { c->state = LM_DONE; return; }
    if ((c->iniChi - c->currentChi) * 1e3 < c->iniChi) c->nbad++; else c->nbad = 0;
    if (c->nbad >= 3 || c->iteration >= c->max_iterations || stop) { c->state = LM_DONE; return; }
    c->state = LM_BUILD;
    // start of the next iteration (levenberg.cpp:69-100 without the first-iteration lambda initialisation, which k_lm_iter_begin does in slot 0)
    c->fresh_first = 0;
    c->iniChi = c->currentChi;
    c->qmax = 0;
    c->rho = 0;
}

__global__ void k_lm_decide(const BaDev B) { lm_decide(B); }

// after the robust stage: edges with chi2 > 5.991 or non-positive depth leave the optimisation (setLevel(1), Optimizer.cc:691-705);
// also reports chi2 / depth of every edge for the caller's outlier handling (Optimizer.cc:734-766)
__global__ void __launch_bounds__(256)
k_ba_edge_check(const BaDev B, double *__restrict__ chi2, uint8_t *__restrict__ depth_ok, uint8_t *__restrict__ outlier, int gate)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B.E) return;
    const double c = edge_chi2(B.e_err[2 * e], B.e_err[2 * e + 1], B.e_w[e]);
    double Xc[3];
    se3_map(B.pose[B.e_kf[e]], &B.pt[3 * B.e_point[e]], Xc);
    const bool dok = Xc[2] > 0.0;
    chi2[e] = c; depth_ok[e] = dok;
    outlier[e] = (uint8_t)(c > 5.991 || !dok);
    if (gate && (c > 5.991 || !dok)) B.e_level[e] = 1;
}

// activity after a level change: pose_act[k] = the keyframe has an active edge, pt_active[p] likewise
__global__ void __launch_bounds__(256)
k_ba_activity(const BaDev B, int *__restrict__ pose_act)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.P) return;
    bool any = false;
    for (int e = B.pt_start[p]; e < B.pt_start[p + 1]; e++) if (!B.e_level[e]) { any = true; pose_act[B.e_kf[e]] = 1; }
    B.pt_active[p] = any;
    if (any) pose_act[B.K] = 1;
}

// tile adjacency of the reduced system: two free keyframes are coupled iff an active point is seen by both
__global__ void __launch_bounds__(256)
k_ba_tile_adj(const BaDev B, uint8_t *__restrict__ adj, int ng)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.P) return;
    const int e0 = B.pt_start[p], e1 = B.pt_start[p + 1];
    for (int u = e0; u < e1; u++) {
        if (B.e_level[u]) continue;
        const int iu = B.pose_idx[B.e_kf[u]];
        if (iu < 0) continue;
        const int gu = iu / kPosesPerTile;
        for (int v = u + 1; v < e1; v++) {
            if (B.e_level[v]) continue;
            const int iv = B.pose_idx[B.e_kf[v]];
            if (iv < 0) continue;
            const int gv = iv / kPosesPerTile;
            if (gu != gv) { adj[(size_t)gu * ng + gv] = 1; adj[(size_t)gv * ng + gu] = 1; }
        }
    }
}

// Schur pairs: count per point, then emit (key = target block << 32 | ordinal, value = row edge | col edge << 32)
__global__ void __launch_bounds__(256)
k_ba_pair_count(const BaDev B, int *__restrict__ cnt)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.P) return;
    int m = 0;
    for (int e = B.pt_start[p]; e < B.pt_start[p + 1]; e++) if (!B.e_level[e] && B.pose_idx[B.e_kf[e]] >= 0) m++;
    cnt[p] = m * (m + 1) / 2;
}

__global__ void __launch_bounds__(256)
k_ba_pair_emit(const BaDev B, const int *__restrict__ pair_start, unsigned long long *__restrict__ key, unsigned long long *__restrict__ val)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > B.P) return;
    if (p == B.P) return;                             // (the arrays are sized by an upper bound; the unused tail keeps its 0xff.. keys and sorts to the end)
    int t = pair_start[p];
    const int e0 = B.pt_start[p], e1 = B.pt_start[p + 1];
    for (int u = e0; u < e1; u++) {
        if (B.e_level[u]) continue;
        const int iu = B.pose_idx[B.e_kf[u]];
        if (iu < 0) continue;
        const int ru = B.rowbase[iu];
        for (int v = u; v < e1; v++) {
            if (B.e_level[v]) continue;
            const int iv = B.pose_idx[B.e_kf[v]];
            if (iv < 0) continue;
            const int rv = B.rowbase[iv];
            const bool swap = ru >= rv;                 // row = the later row of the permuted system
            const int rr = swap ? ru : rv, rc = swap ? rv : ru, er = swap ? u : v, ec = swap ? v : u;
            const int slot = B.slot_of[(size_t)(rr >> 6) * B.nt + (rc >> 6)];
            const unsigned tgt = ((unsigned)slot << 7) | (unsigned)(((rr & 63) / 6) * kPosesPerTile + (rc & 63) / 6);
            key[t] = ((unsigned long long)tgt << 32) | (unsigned)t;
            val[t] = (unsigned long long)(unsigned)er | ((unsigned long long)(unsigned)ec << 32);
            t++;
        }
    }
}

// segment heads of the sorted pair list: head[t] = 1 where the target block changes (valid pairs only)
__global__ void __launch_bounds__(256)
k_ba_seg_heads(const unsigned long long *__restrict__ key, const int *__restrict__ n_pairs, int *__restrict__ head, int n_cap)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cap) return;
    const int n = *n_pairs;
    head[t] = (t < n && (t == 0 || (key[t] >> 32) != (key[t - 1] >> 32))) ? 1 : 0;
}

// seg_start[id] = t for heads (id = inclusive scan - 1), seg_start[nseg] = n_pairs, *n_seg = nseg
__global__ void __launch_bounds__(256)
k_ba_seg_starts(const int *__restrict__ head, const int *__restrict__ scan, const int *__restrict__ n_pairs, int *__restrict__ seg_start, int *__restrict__ n_seg, int n_cap)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cap) return;
    if (head[t]) seg_start[scan[t] - 1] = t;
    if (t == n_cap - 1) { const int ns = scan[t]; seg_start[ns] = *n_pairs; *n_seg = ns; }
}

// float32 inputs of the caller widened on the device (the host stages 4-byte values: its copy loop is bound by one core's memory bandwidth)
__global__ void __launch_bounds__(256)
k_ba_widen(size_t n_obs, const float *__restrict__ obs, double *__restrict__ obs_d, size_t n_w, const float *__restrict__ w, double *__restrict__ w_d,
           size_t n_pt, const float *__restrict__ pt, double *__restrict__ pt_d)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_obs) obs_d[i] = (double)obs[i];
    if (i < n_w) w_d[i] = (double)w[i];
    if (i < n_pt) pt_d[i] = (double)pt[i];
}

__global__ void k_ba_iota(int n, int *__restrict__ v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

__global__ void k_ba_import_poses(int K, const float *__restrict__ T, Se3 *__restrict__ pose)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) { Se3 s; se3_from_Tcw(T + 16 * i, s); pose[i] = s; }
}

__global__ void k_ba_export_poses(int K, const Se3 *__restrict__ pose, const uint8_t *__restrict__ fixed, float *__restrict__ T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K && fixed[i] != 2) se3_to_Tcw(pose[i], T + 16 * i);
}

// points back to float (MapPoint::SetWorldPos takes a float cv::Mat, Converter.cc:101-108)
__global__ void k_ba_export_points(int P, const double *__restrict__ pt, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * P) out[i] = (float)pt[i];
}

}  // namespace orbs
