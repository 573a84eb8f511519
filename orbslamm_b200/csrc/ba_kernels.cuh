// ba_kernels.cuh -- bundle-adjustment kernels (Optimizer::LocalBundleAdjustment / BundleAdjustment on g2o's
// BlockSolver_6_3 + Levenberg, restated).  All fp64.  Edges are stored grouped by map point (CSR), which makes
// the per-point work (Hll, b_l, Schur complement, back-substitution) a warp-level job without atomics; pose
// blocks are built by a second CSR (by keyframe).  Only the scatter of the Schur complement into the dense
// reduced system uses fp64 atomics.
//
//   k_ba_errors        computeActiveErrors + activeRobustChi2          sparse_optimizer.cpp:61-114
//   k_ba_build_points  linearizeOplus + constructQuadraticForm (Hll, b_l, Hpl)   block_solver.hpp:506-564,
//   k_ba_build_poses   ... (Hpp, b_p)                                  base_binary_edge.hpp:55-120
//   k_ba_schur_init/k_ba_schur   setLambda + Schur complement           block_solver.hpp:371-436,568-593
//   k_chol_* / k_trs_*  dense Cholesky + triangular solves of the reduced system (replaces LinearSolverEigen)
//   k_ba_backsub       x_l = Dinv (b_l - Hpl^T x_p) + computeScale     block_solver.hpp:463-487, levenberg.cpp:182-189
//   k_ba_update/restore  oplus (exp(dx) * T, X += dx) with push/pop     sparse_optimizer.cpp:422-435
#pragma once
#include "common.cuh"
#include "se3.cuh"

namespace orbs {

struct BaDev {
    int K, P, E;
    int nA;                       // free active poses (reduced system has n = 6 nA unknowns)
    int n;                        // 6 * nA
    int ld;                       // leading dimension of S (n rounded up to the tile size)
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
    // edges grouped by pose
    const int *pose_start;        // [K+1]
    const int *pose_edges;        // [E] -> edge index (point-grouped order)
    const int *e_point;           // [E] point of an edge (point-grouped order)
    // index mapping
    const int *pose_idx;          // [K] hessian index or -1 (fixed / inactive)
    const int *rowbase;           // [nA] first row of a free keyframe's 6x6 block in the (tile-permuted) reduced system
    const uint8_t *row_pad;       // [ld] 1 = identity padding row of the reduced system
    const uint8_t *pt_active;     // [P]
    // system
    double *Hpp, *bp;             // [nA*36], [nA*6]
    double *Hll, *bl;             // [P*9], [P*3]
    double *x;                    // [n + 3P]  (poses by hessian index, points by point id)
    double *S, *bs;               // [ld*ld] lower triangle used, [ld]
    double *partial;              // [blocks] reduction scratch
    double *scalars;              // [8]: 0 chi2, 1 scale (points), 2 scale (poses), 3 max diagonal
    int *flags;                   // [4]: 0 cholesky failure
    double delta, dsqr;
    int robust;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ double edge_chi2(double e0, double e1, double w) { return e0 * (w * e0 + 0.0 * e1) + e1 * (0.0 * e0 + w * e1); }

// errors of all active edges at the current estimate + robust chi2 (per-block partial sums, fixed order)
__global__ void __launch_bounds__(256)
k_ba_errors(const BaDev B)
{
    __shared__ double s_w[8];
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
    chi = warp_sum(chi);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = chi;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; w++) s += s_w[w]; B.partial[blockIdx.x] = s; }
}

// sum partial[0..n) in fixed order into scalars[slot]
__global__ void __launch_bounds__(256)
k_reduce_partials(const double *__restrict__ partial, int n, double *__restrict__ scalars, int slot, int use_max)
{
    __shared__ double s[256];
    double a = 0;
    for (int i = threadIdx.x; i < n; i += 256) a = use_max ? fmax(a, partial[i]) : a + partial[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) s[threadIdx.x] = use_max ? fmax(s[threadIdx.x], s[threadIdx.x + d]) : s[threadIdx.x] + s[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) scalars[slot] = s[0];
}

// warp per point: Hll, b_l and the per-edge Hpl blocks
__global__ void __launch_bounds__(256)
k_ba_build_points(const BaDev B)
{
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= B.P || !B.pt_active[p]) return;
    double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    const double X[3] = {B.pt[3 * p], B.pt[3 * p + 1], B.pt[3 * p + 2]};
    for (int e = B.pt_start[p] + lane; e < B.pt_start[p + 1]; e += 32) {
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
            double *W = &B.e_W[18 * (size_t)e];
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int c = 0; c < 3; c++) W[3 * a + c] = Jp[a] * wo * Jl[c] + Jp[6 + a] * wo * Jl[3 + c];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) H[i] = warp_sum(H[i]);
#pragma unroll
    for (int i = 0; i < 3; i++) b[i] = warp_sum(b[i]);
    if (lane == 0) {
        double *Ho = &B.Hll[9 * (size_t)p];
        Ho[0] = H[0]; Ho[1] = H[1]; Ho[2] = H[2]; Ho[3] = H[1]; Ho[4] = H[3]; Ho[5] = H[4]; Ho[6] = H[2]; Ho[7] = H[4]; Ho[8] = H[5];
        B.bl[3 * p] = b[0]; B.bl[3 * p + 1] = b[1]; B.bl[3 * p + 2] = b[2];
    }
}

// warp per free pose: Hpp (full 6x6) and b_p; also per-pose max |diag| candidates are taken later on the host side kernel
__global__ void __launch_bounds__(256)
k_ba_build_poses(const BaDev B)
{
    const int kf = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (kf >= B.K) return;
    const int ip = B.pose_idx[kf];
    if (ip < 0) return;
    double acc[27];
#pragma unroll
    for (int i = 0; i < 27; i++) acc[i] = 0;
    const Se3 T = B.pose[kf];
    const double fx = B.intr[4 * kf], fy = B.intr[4 * kf + 1];
    for (int j = B.pose_start[kf] + lane; j < B.pose_start[kf + 1]; j += 32) {
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
    if (lane == 0) {
        double *H = &B.Hpp[36 * (size_t)ip];
        int q = 0;
        for (int a = 0; a < 6; a++) for (int c = a; c < 6; c++) { H[6 * a + c] = acc[q]; H[6 * c + a] = acc[q]; q++; }
        for (int a = 0; a < 6; a++) B.bp[6 * ip + a] = acc[21 + a];
    }
}

// max |diagonal| of the free blocks (computeLambdaInit, levenberg.cpp:166-180): per-block partial maxima
__global__ void __launch_bounds__(256)
k_ba_max_diag(const BaDev B)
{
    __shared__ double s_w[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0;
    if (i < B.nA) for (int j = 0; j < 6; j++) m = fmax(m, fabs(B.Hpp[36 * (size_t)i + 7 * j]));
    const int p = i - B.nA;
    if (p >= 0 && p < B.P && B.pt_active[p]) for (int j = 0; j < 3; j++) m = fmax(m, fabs(B.Hll[9 * (size_t)p + 4 * j]));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; w++) s = fmax(s, s_w[w]); B.partial[blockIdx.x] = s; }
}

// S (lower triangle, ld x ld, zeroed by a memset before) <- diag blocks Hpp + lambda I ; bs <- bp ; padding diag = 1
__global__ void __launch_bounds__(256)
k_ba_schur_init(const BaDev B, double lambda, int lead)
{
    // sharded solve: Hpp / bp are already summed over the ranks; they (and lambda, and the padding diagonal) enter through the
    // lead rank only, the other ranks start from zero and contribute their points' Schur terms
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B.nA * 36) {
        const int i = t / 36, a = (t % 36) / 6, c = t % 6;
        const int rb = B.rowbase[i];
        if (c <= a && lead) B.S[(size_t)(rb + a) * B.ld + rb + c] = B.Hpp[t] + (a == c ? lambda : 0.0);
    }
    if (t < B.n) { const int i = t / 6, a = t % 6; B.bs[B.rowbase[i] + a] = lead ? B.bp[t] : 0.0; }
    if (t < B.ld && B.row_pad[t]) { B.bs[t] = 0.0; if (lead) B.S[(size_t)t * B.ld + t] = 1.0; }
}

// warp per point: Dinv = (Hll + lambda I)^-1, then for every pair of its free-pose edges the 6x6 block
// W_row Dinv W_col^T is subtracted from the lower triangle of S, and W Dinv b_l from bs.
__global__ void __launch_bounds__(256)
k_ba_schur(const BaDev B, double lambda)
{
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= B.P || !B.pt_active[p]) return;
    double D[9], Di[9];
#pragma unroll
    for (int i = 0; i < 9; i++) D[i] = B.Hll[9 * (size_t)p + i];
    D[0] += lambda; D[4] += lambda; D[8] += lambda;
    inv3(D, Di);
    const double b0 = B.bl[3 * p], b1 = B.bl[3 * p + 1], b2 = B.bl[3 * p + 2];
    const double db[3] = {Di[0] * b0 + Di[1] * b1 + Di[2] * b2, Di[3] * b0 + Di[4] * b1 + Di[5] * b2, Di[6] * b0 + Di[7] * b1 + Di[8] * b2};
    const int e0 = B.pt_start[p], m = B.pt_start[p + 1] - e0;
    // bs part: one edge per lane
    for (int u = lane; u < m; u += 32) {
        const int e = e0 + u;
        if (B.e_level[e]) continue;
        const int i1 = B.pose_idx[B.e_kf[e]];
        if (i1 < 0) continue;
        const double *W = &B.e_W[18 * (size_t)e];
        const int rb = B.rowbase[i1];
#pragma unroll
        for (int a = 0; a < 6; a++) atomicAdd(&B.bs[rb + a], -(W[3 * a] * db[0] + W[3 * a + 1] * db[1] + W[3 * a + 2] * db[2]));
    }
    // pair part: pairs (u, v), u <= v, linearised
    const int npairs = m * (m + 1) / 2;
    for (int t = lane; t < npairs; t += 32) {
        // invert t = v (v + 1) / 2 + u
        int v = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while (v * (v + 1) / 2 > t) v--;
        while ((v + 1) * (v + 2) / 2 <= t) v++;
        const int u = t - v * (v + 1) / 2;
        const int eu = e0 + u, ev = e0 + v;
        if (B.e_level[eu] || B.e_level[ev]) continue;
        const int iu = B.pose_idx[B.e_kf[eu]], iv = B.pose_idx[B.e_kf[ev]];
        if (iu < 0 || iv < 0) continue;
        // block (row = lower in the permuted system, col = the other) = W_row Dinv W_col^T
        const int ru = B.rowbase[iu], rv = B.rowbase[iv];
        const bool swap = ru > rv;
        const double *Wr = &B.e_W[18 * (size_t)(swap ? eu : ev)], *Wc = &B.e_W[18 * (size_t)(swap ? ev : eu)];
        const int ir = swap ? ru : rv, ic = swap ? rv : ru;
        double BD[18];
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = 0; c < 3; c++) BD[3 * a + c] = Wr[3 * a] * Di[c] + Wr[3 * a + 1] * Di[3 + c] + Wr[3 * a + 2] * Di[6 + c];
        double *Sb = &B.S[(size_t)ir * B.ld + ic];
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                if (ir == ic && c > a) continue;                   // diagonal block: lower part only
                const double val = BD[3 * a] * Wc[3 * c] + BD[3 * a + 1] * Wc[3 * c + 1] + BD[3 * a + 2] * Wc[3 * c + 2];
                atomicAdd(&Sb[(size_t)a * B.ld + c], -val);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Sparse tiled Cholesky S = L L^T of the reduced pose system (lower triangle, 64x64 tiles).  Replaces
// LinearSolverEigen::solve (SimplicialLDLT with a fill-reducing ordering, linear_solver_eigen.h:94-124); a non-positive pivot
// raises flags[0] (solve() == false).
//
// Layout: free keyframes are packed kPosesPerTile = 10 to a tile (60 rows + 4 identity padding rows), and the TILES are
// permuted by a nested-dissection ordering of the keyframe covisibility graph (host, optimizer.cu: build_schedule): the
// elimination DAG of a trajectory-like graph then has depth O(log) instead of one dependent step per tile column.
// rowbase[hessian index] is the first row of a keyframe's 6x6 block.  The symbolic factorisation (tile level) gives, per
// tile column k, rows(k) = { i > k : L_ik != 0 } and the elimination level of k; columns of one level are independent.
//
// Two launches per level:
//   k_chol_panel   one CTA per (k, i), i = k or i in rows(k): every CTA factors the diagonal tile A_kk redundantly in shared
//                  memory (cheaper than a dependent launch); CTA (k, k) stores inv(L_kk) for the triangular solves, CTA (k, i)
//                  writes L_ik = A_ik L_kk^-T.  The diagonal tiles of L are not written back (never read again).
//   k_chol_update  one CTA per (k, i, j), i >= j in rows(k): A_ij -= L_ik L_jk^T (fp64 atomics only where two columns of the
//                  same level update the same tile).
constexpr int NB = 64;
constexpr int kPosesPerTile = 10;
constexpr int kTilePitch = NB + 1;                      // doubles; conflict-free for row- and column-wise access
constexpr int kTileElems = NB * kTilePitch;
constexpr int kPanelSmem = 2 * kTileElems * (int)sizeof(double);
constexpr int kUpdateSmem = 2 * kTileElems * (int)sizeof(double);

struct CholPlan {
    const int *rows_start;  // [nt + 1]  CSR by column: rows(k), ascending
    const int *rows;
    const int *cols_start;  // [nt + 1]  CSR by row: cols(i) = { k < i : L_ik != 0 }, ascending
    const int *cols;
    const int4 *panel;      // tasks {k, i, 0, 0}, grouped by level
    const int4 *update;     // tasks {k, i, j, shared}, grouped by level
};

// Sharded solve: only the structurally nonzero tiles of the reduced system travel over NVLink.  The panel task list names exactly the tiles
// (i, k) of the lower triangle that S or its factor can touch; pack them into one contiguous buffer for the all-reduce and scatter them back.
__global__ void __launch_bounds__(256)
k_tiles_pack(const double *__restrict__ S, int ld, const int4 *__restrict__ tiles, double *__restrict__ packed, int to_packed)
{
    const int4 t = tiles[blockIdx.x];                 // {k, i, 0, 0}: tile row i, tile column k
    double *tile = const_cast<double *>(S) + (size_t)t.y * 64 * ld + (size_t)t.x * 64;
    double *buf = packed + (size_t)blockIdx.x * 4096;
    for (int q = threadIdx.x; q < 4096; q += 256) {
        const int r = q >> 6, c = q & 63;
        if (to_packed) buf[q] = tile[(size_t)r * ld + c]; else tile[(size_t)r * ld + c] = buf[q];
    }
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}

typedef double (*TilePtr)[kTilePitch];

__device__ __forceinline__ void tile_load_async(TilePtr dst, const double *src, int ld, int tid)
{
#pragma unroll 4
    for (int q = tid; q < NB * NB; q += 256) { const int r = q >> 6, c = q & 63; cp_async8(&dst[r][c], &src[(size_t)r * ld + c]); }
}

// 256 threads: 4 threads per matrix row, each holding 16 consecutive columns of the row in registers.
__global__ void __launch_bounds__(256)
k_chol_panel(double *__restrict__ S, int ld, const int4 *__restrict__ tasks, double *__restrict__ Linv, int *__restrict__ flags)
{
    extern __shared__ double smem_d[];
    TilePtr T = reinterpret_cast<TilePtr>(smem_d);                      // A_kk, then L_kk (lower)
    TilePtr X = reinterpret_cast<TilePtr>(smem_d + kTileElems);         // the tile being solved
    __shared__ double colbuf[2][NB];
    __shared__ double invd[NB];
    const int tid = threadIdx.x;
    const int4 task = tasks[blockIdx.x];
    const int k = task.x, i = task.y;
    const bool diag = i == k;
    tile_load_async(T, S + (size_t)(k * NB) * ld + k * NB, ld, tid);
    if (!diag) tile_load_async(X, S + (size_t)(i * NB) * ld + k * NB, ld, tid);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();

    const int r = tid >> 2, sub = tid & 3, lane = tid & 31;
    double a[16];
#pragma unroll
    for (int u = 0; u < 16; u++) a[u] = T[r][16 * sub + u];
    // right-looking Cholesky of the diagonal tile (lower part only), one barrier per column
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int js = j >> 4, ju = j & 15;
        if (sub == js && r >= j) colbuf[j & 1][r] = a[ju];
        __syncthreads();
        double d = colbuf[j & 1][j];
        if (!(d > 0.0)) { if (diag && tid == 0) flags[0] = 1; d = 1.0; }
        const double rinv = rsqrt(d), sd = d * rinv;
        if (r >= j) {
            const double l = (r == j) ? sd : colbuf[j & 1][r] * rinv;
            if (sub == js) a[ju] = l;
            const double lr = l * rinv;
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int c = 16 * sub + u;
                if (c > j && c <= r) a[u] = fma(-lr, colbuf[j & 1][c], a[u]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 16; u++) { const int c = 16 * sub + u; T[r][c] = c <= r ? a[u] : 0.0; }
    __syncthreads();
    if (tid < NB) invd[tid] = 1.0 / T[tid][tid];
    __syncthreads();
    // Solve X L^T = A row by row (left-looking): the 4 threads of row r own the entries x_q with q % 4 == sub; entry c needs
    // sum_{q<c} x_q L[c][q] (L row c is a broadcast read), reduced over the 4 threads with two shuffles.  The diagonal CTA
    // solves for A = I, i.e. X = L^-T, and stores its transpose L^-1 for the triangular solves.
    const unsigned full = 0xffffffffu;
    double xo[16];
#pragma unroll
    for (int m = 0; m < 16; m++) { const int q = 4 * m + sub; xo[m] = diag ? (q == r ? 1.0 : 0.0) : X[r][q]; }
#pragma unroll
    for (int c = 0; c < NB; c++) {
        double p0 = 0.0, p1 = 0.0;
#pragma unroll
        for (int m = 0; m < c / 4; m++) {
            if (m & 1) p1 = fma(xo[m], T[c][4 * m + sub], p1); else p0 = fma(xo[m], T[c][4 * m + sub], p0);
        }
        if (sub < (c & 3)) p0 = fma(xo[c / 4], T[c][4 * (c / 4) + sub], p0);
        double p = p0 + p1;
        p += __shfl_xor_sync(full, p, 1);
        p += __shfl_xor_sync(full, p, 2);
        if (sub == (c & 3)) xo[c / 4] = (xo[c / 4] - p) * invd[c];
    }
    __syncthreads();
    if (diag) {
        double *Li = Linv + (size_t)k * NB * NB;
#pragma unroll
        for (int m = 0; m < 16; m++) Li[(size_t)(4 * m + sub) * NB + r] = xo[m];          // L^-1[c][r] = X[r][c]
    } else {
#pragma unroll
        for (int m = 0; m < 16; m++) X[r][4 * m + sub] = xo[m];
        __syncthreads();
        double *P = S + (size_t)(i * NB) * ld + k * NB;
        for (int q = tid; q < NB * NB; q += 256) { const int rr = q >> 6, cc = q & 63; P[(size_t)rr * ld + cc] = X[rr][cc]; }
    }
}

// A[i][j] -= A[i][k] * A[j][k]^T.  Panels staged row-major with cp.async (pitch 65 doubles: conflict-free for both operands),
// thread (ty, tx) owns the interleaved 4x4 micro-tile C[ty + 16u][tx + 16v].
__global__ void __launch_bounds__(256)
k_chol_update(double *__restrict__ S, int ld, const int4 *__restrict__ tasks)
{
    extern __shared__ double smem_d[];
    TilePtr Ai = reinterpret_cast<TilePtr>(smem_d);
    TilePtr Aj = reinterpret_cast<TilePtr>(smem_d + kTileElems);
    const int4 task = tasks[blockIdx.x];
    const int k = task.x, i = task.y, j = task.z, tid = threadIdx.x;
    const bool shared_target = task.w != 0;
    tile_load_async(Ai, S + (size_t)(i * NB) * ld + k * NB, ld, tid);
    if (i != j) tile_load_async(Aj, S + (size_t)(j * NB) * ld + k * NB, ld, tid);
    asm volatile("cp.async.commit_group;\n" ::);
    const int ty = tid >> 4, tx = tid & 15;
    double *C = S + (size_t)(i * NB) * ld + j * NB;
    double cval[4][4];
    if (!shared_target) {
        // prefetch the C micro-tile while the panels land
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) cval[u][v] = C[(size_t)(ty + 16 * u) * ld + tx + 16 * v];
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    TilePtr Bj = i != j ? Aj : Ai;
    double acc[4][4] = {};
#pragma unroll 8
    for (int q = 0; q < NB; q++) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { a[u] = Ai[ty + 16 * u][q]; b[u] = Bj[tx + 16 * u][q]; }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++)
            if (i != j || tx + 16 * v <= ty + 16 * u) {
                double *dst = &C[(size_t)(ty + 16 * u) * ld + tx + 16 * v];
                if (shared_target) atomicAdd(dst, -acc[u][v]); else *dst = cval[u][v] - acc[u][v];
            }
}

// Triangular solves as two dataflow kernels.  CTA b handles the tile rows b, b + G, b + 2G, ... (G = gridDim.x <= the number of
// co-resident CTAs) in solve order; a row only waits for rows that come earlier in that order, and every CTA walks its rows in
// that order, so the wait graph is acyclic for any number of tile rows.
//   forward  (dir = 0): row i waits for y_k, k in cols(i), accumulates L_ik y_k, then y_i = Linv_ii (b_i - acc)
//   backward (dir = 1): row j (visited in descending order) waits for x_i, i in rows(j), accumulates L_ij^T x_i, then
//                       x_j = Linv_jj^T (y_j - acc)
// v is solved in place; ready[] must be zero on entry.
__global__ void __launch_bounds__(256)
k_chol_solve(const double *__restrict__ S, int ld, int nt, const CholPlan plan, const double *__restrict__ Linv, double *v, int *ready, int dir)
{
    __shared__ double s_vec[NB];
    __shared__ double s_part[4][NB];
    __shared__ double s_acc[NB];
    const int tid = threadIdx.x, lane64 = tid & 63, part = tid >> 6;
    for (int ord = blockIdx.x; ord < nt; ord += gridDim.x) {
        const int me = dir == 0 ? ord : nt - 1 - ord;
        if (tid < NB) s_acc[tid] = 0.0;
        __syncthreads();
        const int p0 = dir == 0 ? plan.cols_start[me] : plan.rows_start[me], p1 = dir == 0 ? plan.cols_start[me + 1] : plan.rows_start[me + 1];
        for (int s = 0; s < p1 - p0; s++) {
            const int o = dir == 0 ? plan.cols[p0 + s] : plan.rows[p1 - 1 - s];     // the tile whose solution we consume
            if (tid == 0) { while (*((volatile int *)&ready[o]) == 0) { } __threadfence(); }
            __syncthreads();
            if (tid < NB) s_vec[tid] = *((volatile double *)&v[o * NB + tid]);
            __syncthreads();
            double sum = 0;
            if (dir == 0) {
                // acc[r] += sum_c L[me*64 + r][o*64 + c] * y_o[c]; thread (r = lane64, quarter = part)
                const double *row = S + (size_t)(me * NB + lane64) * ld + o * NB + part * 16;
#pragma unroll
                for (int c = 0; c < 16; c++) sum = fma(row[c], s_vec[part * 16 + c], sum);
            } else {
                // acc[c] += sum_r L[o*64 + r][me*64 + c] * x_o[r]; thread (c = lane64, quarter = part)
                const double *col = S + (size_t)(o * NB + part * 16) * ld + me * NB + lane64;
#pragma unroll
                for (int r = 0; r < 16; r++) sum = fma(col[(size_t)r * ld], s_vec[part * 16 + r], sum);
            }
            s_part[part][lane64] = sum;
            __syncthreads();
            if (tid < NB) s_acc[tid] += s_part[0][tid] + s_part[1][tid] + s_part[2][tid] + s_part[3][tid];
            __syncthreads();
        }
        // diagonal tile through its explicit inverse
        if (tid < NB) s_vec[tid] = v[me * NB + tid] - s_acc[tid];
        __syncthreads();
        const double *Li = Linv + (size_t)me * NB * NB;
        double sum = 0;
        if (dir == 0) { for (int c = part * 16; c < part * 16 + 16; c++) sum = fma(Li[lane64 * NB + c], s_vec[c], sum); }     // y = Linv r
        else { for (int r = part * 16; r < part * 16 + 16; r++) sum = fma(Li[r * NB + lane64], s_vec[r], sum); }               // x = Linv^T r
        s_part[part][lane64] = sum;
        __syncthreads();
        if (tid < NB) v[me * NB + tid] = s_part[0][tid] + s_part[1][tid] + s_part[2][tid] + s_part[3][tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) { *((volatile int *)&ready[me]) = 1; }
    }
}

// copy the pose part of the solution; pose part of computeScale
__global__ void __launch_bounds__(256)
k_ba_take_xp(const BaDev B, double lambda, int lead)
{
    __shared__ double s_w[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double sc = 0;
    if (i < B.n) { const double x = B.bs[B.rowbase[i / 6] + i % 6]; B.x[i] = x; sc = lead ? x * (lambda * x + B.bp[i]) : 0.0; }   // pose part counted once (lead rank)
    sc = warp_sum(sc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; w++) s += s_w[w]; B.partial[blockIdx.x] = s; }
}

// warp per point: x_l = Dinv (b_l - sum_i Hpl_i^T x_i); point part of computeScale (per-block partials)
__global__ void __launch_bounds__(256)
k_ba_backsub(const BaDev B, double lambda)
{
    __shared__ double s_w[8];
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    double sc = 0;
    if (p < B.P && B.pt_active[p]) {
        double c[3] = {0, 0, 0};
        for (int e = B.pt_start[p] + lane; e < B.pt_start[p + 1]; e += 32) {
            if (B.e_level[e]) continue;
            const int i1 = B.pose_idx[B.e_kf[e]];
            if (i1 < 0) continue;
            const double *W = &B.e_W[18 * (size_t)e];
            const double *xp = &B.x[6 * i1];
#pragma unroll
            for (int a = 0; a < 6; a++) { c[0] -= W[3 * a] * xp[a]; c[1] -= W[3 * a + 1] * xp[a]; c[2] -= W[3 * a + 2] * xp[a]; }
        }
        c[0] = warp_sum(c[0]); c[1] = warp_sum(c[1]); c[2] = warp_sum(c[2]);
        if (lane == 0) {
            double D[9], Di[9];
            for (int i = 0; i < 9; i++) D[i] = B.Hll[9 * (size_t)p + i];
            D[0] += lambda; D[4] += lambda; D[8] += lambda;
            inv3(D, Di);
            const double bl[3] = {B.bl[3 * p], B.bl[3 * p + 1], B.bl[3 * p + 2]};
            c[0] += bl[0]; c[1] += bl[1]; c[2] += bl[2];
            double *xl = &B.x[B.n + 3 * (size_t)p];
            for (int a = 0; a < 3; a++) {
                xl[a] = Di[3 * a] * c[0] + Di[3 * a + 1] * c[1] + Di[3 * a + 2] * c[2];
                sc += xl[a] * (lambda * xl[a] + bl[a]);
            }
        }
    }
    if (lane == 0) s_w[threadIdx.x >> 5] = sc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; w++) s += s_w[w]; B.partial[blockIdx.x] = s; }
}

// push + oplus
__global__ void __launch_bounds__(256)
k_ba_update(const BaDev B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.K) {
        const int ip = B.pose_idx[i];
        if (ip >= 0) {
            const Se3 T = B.pose[i];
            B.pose_bak[i] = T;
            Se3 d, r;
            se3_exp(&B.x[6 * ip], d);
            se3_mul(d, T, r);
            B.pose[i] = r;
        }
    }
    const int p = i - B.K;
    if (p >= 0 && p < B.P && B.pt_active[p]) {
        for (int a = 0; a < 3; a++) { const double v = B.pt[3 * p + a]; B.pt_bak[3 * p + a] = v; B.pt[3 * p + a] = v + B.x[B.n + 3 * (size_t)p + a]; }
    }
}

// pop
__global__ void __launch_bounds__(256)
k_ba_restore(const BaDev B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.K && B.pose_idx[i] >= 0) B.pose[i] = B.pose_bak[i];
    const int p = i - B.K;
    if (p >= 0 && p < B.P && B.pt_active[p]) for (int a = 0; a < 3; a++) B.pt[3 * p + a] = B.pt_bak[3 * p + a];
}

// stored chi2 and current depth sign of every edge (Optimizer.cc:691-705, 734-766)
__global__ void __launch_bounds__(256)
k_ba_edge_check(const BaDev B, double *__restrict__ chi2, uint8_t *__restrict__ depth_ok)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B.E) return;
    chi2[e] = edge_chi2(B.e_err[2 * e], B.e_err[2 * e + 1], B.e_w[e]);
    double Xc[3];
    se3_map(B.pose[B.e_kf[e]], &B.pt[3 * B.e_point[e]], Xc);
    depth_ok[e] = Xc[2] > 0.0;
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

}  // namespace orbs
