// pose_graph.cuh -- numeric core of Optimizer::OptimizeEssentialGraph / MMOptimizeEssentialGraph (S/src/Optimizer.cc:804-1067, 1069-1346): a pose
// graph of VertexSim3Expmap vertices and EdgeSim3 edges, error = log(Sji * Siw * Sjw^-1) (types_seven_dof_expmap.h:64-84, sim3.h:110-181), identity
// information, numeric Jacobians for both vertices (base_binary_edge.hpp:131-205, delta = 1e-9), g2o's Levenberg from lambda = 1e-16.
// The 7K x 7K normal equations are as sparse as the reduced BA system, so they go through the same tiled sparse solver (tile_solver.cuh:
// packed nonzero 64x64 tiles, nested-dissection tile order, one persistent dataflow kernel): 9 vertices per 64-row tile (63 rows + 1
// identity row).  The normal equations are assembled per TARGET 7x7 block in a fixed order (no atomics): results are bit-reproducible.
#pragma once
#include "common.cuh"
#include "sim3_opt.cuh"
#include "tile_solver.cuh"

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
    // W = A*Omega + B*Omega*Omega + C*I evaluates left to right, (B*Omega)*Omega (pinned against the reference's g2o object code)
    double BO[9];
    for (int i = 0; i < 9; i++) BO[i] = B * O[i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) O2[3 * r + c] = BO[3 * r] * O[c] + BO[3 * r + 1] * O[3 + c] + BO[3 * r + 2] * O[6 + c];
    for (int i = 0; i < 9; i++) W[i] = A * O[i] + O2[i] + C * (i % 4 == 0 ? 1.0 : 0.0);
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

struct PgBlock { int slot, r0, c0, diag, start, n, rs, pad; };   // target block: tile slot, offsets inside the tile, entries [start, start + n), rs = global row (diag: b)
struct PgEntry { int e, s_row, s_col, pad; };                      // contribution J[e][s_row]^T J[e][s_col]

struct PgDev {
    int K, E, nA, nt, ns, fix_scale, n_blocks;
    Sim3 *verts; Sim3 *backup;        // [K]
    const int *hidx;                  // [K] hessian index or -1 (fixed)
    const int *rowbase;               // [nA] first row of the vertex's 7x7 block in the permuted tile layout
    const uint8_t *row_pad;           // [nt*64] 1 = identity padding row
    const int *e_i, *e_j; const Sim3 *meas;
    double *err;                      // [E,7] stored _error
    double *eJ;                       // [E,2,49] numeric Jacobians of both vertices (row-major 7x7: J[r][d])
    const PgBlock *blocks; const PgEntry *entries;
    double *H;                        // [ns][4096] normal equations, packed tiles (kept: every LM trial factors H + lambda I)
    double *A;                        // [ns][4096] H + lambda I
    double *b;                        // [nt*64] right-hand side
    double *x;                        // [nt*64] solution
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

// numeric linearizeOplus for both vertices (base_binary_edge.hpp:131-205, central differences, delta = 1e-9): thread per (edge, side, column)
__global__ void __launch_bounds__(128)
k_pg_jac(const PgDev P)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.E * 14) return;
    const int e = t / 14, s = (t % 14) / 7, d = t % 7;
    const int vi = P.e_i[e], vj = P.e_j[e];
    if (P.hidx[s == 0 ? vi : vj] < 0) return;
    const Sim3 Si = P.verts[vi], Sj = P.verts[vj], M = P.meas[e];
    double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[7], em[7];
    Sim3 Pp;
    add[d] = 1e-9; sim3_oplus(s == 0 ? Si : Sj, add, P.fix_scale != 0, Pp);
    pg_edge_error(M, s == 0 ? Pp : Si, s == 1 ? Pp : Sj, ep);
    add[d] = -1e-9; sim3_oplus(s == 0 ? Si : Sj, add, P.fix_scale != 0, Pp);
    pg_edge_error(M, s == 0 ? Pp : Si, s == 1 ? Pp : Sj, em);
    double *J = P.eJ + ((size_t)e * 2 + s) * 49;
#pragma unroll
    for (int r = 0; r < 7; r++) J[7 * r + d] = (1.0 / (2 * 1e-9)) * (ep[r] - em[r]);
}

// constructQuadraticForm per target block (base_binary_edge.hpp:55-120): one warp sums the block's contributions J_row^T J_col in list order
// (fixed = deterministic) and writes it once into the packed tile; diagonal blocks also give b = - sum J^T e
__global__ void __launch_bounds__(256)
k_pg_accum(const PgDev P)
{
    const int blk = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blk >= P.n_blocks) return;
    const PgBlock Bk = P.blocks[blk];
    double g0 = 0, g1 = 0, gb = 0;                       // entries lane and lane + 32 of the 7x7 block; b entry lane (< 7)
    const int q0 = lane, q1 = lane + 32;
    const int a0 = q0 / 7, c0 = q0 % 7, a1 = q1 / 7, c1 = q1 % 7;
    for (int t = 0; t < Bk.n; t++) {
        const PgEntry en = P.entries[Bk.start + t];
        const double *Jr = P.eJ + ((size_t)en.e * 2 + en.s_row) * 49, *Jc = P.eJ + ((size_t)en.e * 2 + en.s_col) * 49;
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int r = 0; r < 7; r++) { s0 += Jr[7 * r + a0] * Jc[7 * r + c0]; if (q1 < 49) s1 += Jr[7 * r + a1] * Jc[7 * r + c1]; }
        g0 += s0; g1 += s1;
        if (Bk.diag && lane < 7) {
            const double *er = P.err + 7 * (size_t)en.e;
            double s = 0;
#pragma unroll
            for (int r = 0; r < 7; r++) s += Jr[7 * r + lane] * er[r];
            gb += s;
        }
    }
    double *T = P.H + (size_t)Bk.slot * TS2 + Bk.r0 * TS + Bk.c0;
    T[a0 * TS + c0] = g0;
    if (q1 < 49) T[a1 * TS + c1] = g1;
    if (Bk.diag && lane < 7) P.b[Bk.rs + lane] = -gb;
}

// A <- H + lambda on the diagonal of the real rows, 1 on the padding rows (A was copied from H)
__global__ void k_pg_prepare(const PgDev P, double lambda)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nt * TS) return;
    double *d = P.A + (size_t)(t >> 6) * TS2 + (t & 63) * TS + (t & 63);          // diagonal tile i is slot i
    if (P.row_pad[t]) *d = 1.0; else *d += lambda;
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
    for (int a = 0; a < 7; a++) x[a] = P.x[P.rowbase[h] + a];
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
