// pose_graph.cuh -- numeric core of Optimizer::OptimizeEssentialGraph / MMOptimizeEssentialGraph (S/src/Optimizer.cc:804-1067, 1069-1346): a pose
// graph of VertexSim3Expmap vertices and EdgeSim3 edges, error = log(Sji * Siw * Sjw^-1) (types_seven_dof_expmap.h:64-84, sim3.h:110-181), identity
// information, numeric Jacobians for both vertices (base_binary_edge.hpp:131-205, delta = 1e-9), g2o's Levenberg from lambda = 1e-16.
// The 7K x 7K normal equations are as sparse as the reduced BA system, so they go through the same tiled sparse Cholesky (ba_kernels.cuh:
// k_chol_panel / k_chol_update / k_chol_solve, nested-dissection tile order): 9 vertices per 64-row tile (63 rows + 1 identity row).
#pragma once
#include "common.cuh"
#include "sim3_opt.cuh"

namespace orbs {

constexpr int kPgPerTile = 9;

__device__ inline void lu3_solve(const double *A, const double *b, double *x)          // Eigen PartialPivLU, 3x3
{
    double M[3][4] = {{A[0], A[1], A[2], b[0]}, {A[3], A[4], A[5], b[1]}, {A[6], A[7], A[8], b[2]}};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int p = c;
        for (int r = c + 1; r < 3; r++) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
        if (p != c) for (int k = 0; k < 4; k++) { const double t = M[c][k]; M[c][k] = M[p][k]; M[p][k] = t; }
        for (int r = c + 1; r < 3; r++) { const double f = M[r][c] / M[c][c]; for (int k = c; k < 4; k++) M[r][k] -= f * M[c][k]; }
    }
    for (int r = 2; r >= 0; r--) { double s = M[r][3]; for (int k = r + 1; k < 3; k++) s -= M[r][k] * x[k]; x[r] = s / M[r][r]; }
}

__device__ inline void sim3_log(const Sim3 &S, double res[7])                          // Sim3::log, sim3.h:110-181
{
    const double sigma = log(S.s);
    double R[9], omega[3], O[9], O2[9], W[9];
    quat_to_R(S.q, R);
    const double d = 0.5 * (R[0] + R[4] + R[8] - 1), eps = 0.00001;
    const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    double A, B, C;
    if (fabs(sigma) < eps) {
        C = 1;
        if (d > 1 - eps) { for (int k = 0; k < 3; k++) omega[k] = 0.5 * dR[k]; A = 1. / 2.; B = 1. / 6.; }
        else {
            const double theta = acos(d), theta2 = theta * theta;
            for (int k = 0; k < 3; k++) omega[k] = theta / (2 * sqrt(1 - d * d)) * dR[k];
            A = (1 - cos(theta)) / (theta2); B = (theta - sin(theta)) / (theta2 * theta);
        }
    } else {
        C = (S.s - 1) / sigma;
        if (d > 1 - eps) {
            const double sigma2 = sigma * sigma;
            for (int k = 0; k < 3; k++) omega[k] = 0.5 * dR[k];
            A = ((sigma - 1) * S.s + 1) / (sigma2); B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
        } else {
            const double theta = acos(d);
            for (int k = 0; k < 3; k++) omega[k] = theta / (2 * sqrt(1 - d * d)) * dR[k];
            const double theta2 = theta * theta, a = S.s * sin(theta), b = S.s * cos(theta), c = theta2 + sigma * sigma;
            A = (a * sigma + (1 - b) * theta) / (theta * c); B = (C - ((b - 1) * sigma + a * theta) / (c)) * 1. / (theta2);
        }
    }
    O[0] = 0; O[1] = -omega[2]; O[2] = omega[1]; O[3] = omega[2]; O[4] = 0; O[5] = -omega[0]; O[6] = -omega[1]; O[7] = omega[0]; O[8] = 0;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
    for (int i = 0; i < 9; i++) W[i] = A * O[i] + B * O2[i] + C * (i % 4 == 0 ? 1.0 : 0.0);
    double ups[3];
    lu3_solve(W, S.t, ups);
    for (int k = 0; k < 3; k++) { res[k] = omega[k]; res[k + 3] = ups[k]; }
    res[6] = sigma;
}

__device__ inline void pg_edge_error(const Sim3 &meas, const Sim3 &Si, const Sim3 &Sj, double e[7])   // EdgeSim3::computeError
{
    Sim3 Sjinv, a, b;
    sim3_inverse(Sj, Sjinv);
    sim3_mul(meas, Si, a);
    sim3_mul(a, Sjinv, b);
    sim3_log(b, e);
}

struct PgDev {
    int K, E, nA, ld, fix_scale;
    Sim3 *verts; Sim3 *backup;        // [K]
    const int *hidx;                  // [K] hessian index or -1 (fixed)
    const int *rowbase;               // [nA] first row of the vertex's 7x7 block in the permuted tile layout
    const uint8_t *row_pad;           // [ld] 1 = identity padding row
    const int *e_i, *e_j; const Sim3 *meas;
    double *err;                      // [E,7] stored _error
    double *S;                        // [ld, ld] lower triangle, tiles
    double *b;                        // [ld] right-hand side (kept), v = working copy solved in place
    double *v;
    double *partial;                  // per-block chi2 partials
    double *scalars;                  // [4]: chi2
};

// computeActiveErrors + activeChi2 (no robust kernel, identity information)
__global__ void __launch_bounds__(128)
k_pg_errors(const PgDev P)
{
    __shared__ double s_w[4];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0;
    if (e < P.E) {
        const int i = P.e_i[e], j = P.e_j[e];
        if (P.hidx[i] >= 0 || P.hidx[j] >= 0) {
            double er[7];
            pg_edge_error(P.meas[e], P.verts[i], P.verts[j], er);
#pragma unroll
            for (int k = 0; k < 7; k++) { P.err[7 * (size_t)e + k] = er[k]; c = __dadd_rn(c, __dmul_rn(er[k], er[k])); }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) P.partial[blockIdx.x] = s_w[0] + s_w[1] + s_w[2] + s_w[3];
}

__global__ void k_pg_sum(const double *__restrict__ partial, int n, double *__restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { double s = 0; for (int i = 0; i < n; i++) s += partial[i]; out[0] = s; }
}

// buildSystem: numeric linearizeOplus for both vertices + constructQuadraticForm, one thread per edge; 7x7 blocks are added into the lower
// triangle of S (fp64 atomics: several edges share a vertex) and into b
__global__ void __launch_bounds__(64)
k_pg_build(const PgDev P)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.E) return;
    const int vid[2] = {P.e_i[e], P.e_j[e]};
    const int hs[2] = {P.hidx[vid[0]], P.hidx[vid[1]]};
    if (hs[0] < 0 && hs[1] < 0) return;
    const Sim3 Sv[2] = {P.verts[vid[0]], P.verts[vid[1]]};
    const Sim3 M = P.meas[e];
    double J[2][7][7];
    for (int s = 0; s < 2; s++) {
        if (hs[s] < 0) continue;
        for (int d = 0; d < 7; d++) {
            double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[7], em[7];
            Sim3 Pp;
            add[d] = 1e-9; sim3_oplus(Sv[s], add, P.fix_scale != 0, Pp);
            pg_edge_error(M, s == 0 ? Pp : Sv[0], s == 1 ? Pp : Sv[1], ep);
            add[d] = -1e-9; sim3_oplus(Sv[s], add, P.fix_scale != 0, Pp);
            pg_edge_error(M, s == 0 ? Pp : Sv[0], s == 1 ? Pp : Sv[1], em);
            for (int r = 0; r < 7; r++) J[s][r][d] = (1.0 / (2 * 1e-9)) * (ep[r] - em[r]);
        }
    }
    const double *er = P.err + 7 * (size_t)e;
    for (int s = 0; s < 2; s++) {
        if (hs[s] < 0) continue;
        const int rs = P.rowbase[hs[s]];
        for (int a = 0; a < 7; a++) { double g = 0; for (int r = 0; r < 7; r++) g += J[s][r][a] * er[r]; atomicAdd(&P.b[rs + a], -g); }
        for (int t = 0; t < 2; t++) {
            if (hs[t] < 0) continue;
            const int rt = P.rowbase[hs[t]];
            if (rt > rs) continue;                                     // lower triangle only: block row >= block column
            for (int a = 0; a < 7; a++)
                for (int c = 0; c < 7; c++) {
                    if (s == t && c > a) continue;                     // diagonal block: its lower part
                    double g = 0;
                    for (int r = 0; r < 7; r++) g += J[s][r][a] * J[t][r][c];
                    atomicAdd(&P.S[(size_t)(rs + a) * P.ld + rt + c], g);
                }
        }
    }
}

// + lambda on the diagonal of the real rows, 1 on the padding rows; v <- b
__global__ void k_pg_prepare(const PgDev P, double lambda)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.ld) return;
    if (P.row_pad[t]) { P.S[(size_t)t * P.ld + t] = 1.0; P.v[t] = 0.0; }
    else { P.S[(size_t)t * P.ld + t] += lambda; P.v[t] = P.b[t]; }
}

// push + oplus: estimate <- exp(x) * estimate for the free vertices
__global__ void k_pg_update(const PgDev P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.K) return;
    const Sim3 cur = P.verts[k];
    P.backup[k] = cur;
    const int h = P.hidx[k];
    if (h < 0) return;
    double x[7];
    for (int a = 0; a < 7; a++) x[a] = P.v[P.rowbase[h] + a];
    Sim3 nv;
    sim3_oplus(cur, x, P.fix_scale != 0, nv);
    P.verts[k] = nv;
}

__global__ void k_pg_restore(const PgDev P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < P.K) P.verts[k] = P.backup[k];
}

}  // namespace orbs
