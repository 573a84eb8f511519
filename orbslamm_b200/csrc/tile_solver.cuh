// tile_solver.cuh -- sparse tiled Cholesky + triangular solves of a reduced (pose) system as ONE persistent dataflow kernel.
//
// Replaces LinearSolverEigen::solve (SimplicialLDLT under a fill-reducing ordering, linear_solver_eigen.h:94-124) for the
// reduced camera system of BlockSolver_6_3 (block_solver.hpp:371-460) and for the Sim3 pose graph (BlockSolver_7_3).
//
// Storage: only the structurally nonzero 64x64 tiles of the lower triangle exist, packed ("slots"): the system A and its
// factor L are two arrays [ns][64][64] fp64, row-major.  The tile order is a nested-dissection elimination order computed on the
// host from the tile adjacency (TilePlanHost::build), with a symbolic factorisation that yields the slot list.
//
// k_rs_solve: LEFT-LOOKING tile tasks in one launch.  Task (i, j) keeps the 64x64 accumulator A_ij - sum_k L_ik L_jk^T in
// DMMA fragments (mma.sync m8n8k4 f64: the one dense contraction of the path), in ascending k = fixed order = deterministic, no
// atomics, every tile read once and written once; diagonal tasks factor the tile and store inv(L_jj); off-diagonal tasks
// multiply by inv(L_jj)^T (a fourth GEMM instead of a 64-step substitution).  Forward / backward substitution are further
// tasks of the same kernel.  Tasks are fetched from an atomic ticket counter in topological order and wait on per-tile epoch
// flags, so any set of resident CTAs makes progress (no co-residency assumption) and no flag ever needs a reset.
#pragma once
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace orbs {

constexpr int TS = 64;                       // tile size
constexpr int TS2 = TS * TS;
constexpr int kOpPitch = 68;                 // operand tiles in smem: (68 r + c) -> DMMA fragment loads are conflict-free per half-warp
constexpr int kRsSmemBytes = 4 * TS * kOpPitch * (int)sizeof(double);   // two operand double-buffers

enum { LM_BUILD = 0, LM_RETRY = 1, LM_DONE = 2 };

// Device-resident Levenberg-Marquardt control block (OptimizationAlgorithmLevenberg::solve, levenberg.cpp:61-164, and the
// iteration loop of SparseOptimizer::optimize, sparse_optimizer.cpp:354-419): every kernel of an LM "slot" reads it, the
// one-thread decide kernel advances it; the host only enqueues slots and reads a pinned copy a slot or two later.
struct LmCtl {
    double lambda, ni, currentChi, iniChi, tempChi, rho, scale;
    int state;                 // LM_BUILD: next slot linearises first; LM_RETRY: the last trial was rejected, same system, larger lambda; LM_DONE
    int iteration, max_iterations, qmax, nbad, first;
    int lm_iterations, lm_trials, chol_failures;
    int last_rejected;         // the trial just decided was rejected: k_*_restore pops the estimate
    int fresh_first;           // the current slot linearised before lambda was known (first iteration)
    int pad[1];
};

struct RsTask { int type, i, j, slot, dep0, ndep, diag, part0, npart, pad0, pad1, pad2; };
// type 0: factor tile (i, j); 1: forward row i; 2: backward row j; 3: partial update sum of a chunk of a tile's dependencies (slot = partial tile)

struct RsPlan {
    const RsTask *tasks; int ntasks;
    const int2 *deps;          // factor: (slot(i,k), slot(j,k)); forward: (slot(i,k), k); backward: (slot(i,j), i)
    int nt, ns;
};

struct RsBuf {
    const double *A;           // [ns][4096] system tiles (input, untouched)
    double *L;                 // [ns][4096] factor tiles (off-diagonal slots)
    double *Linv;              // [nt][4096] inverses of the diagonal factor tiles
    const double *b;           // [nt*64] right-hand side
    double *y;                 // [nt*64] forward solution
    double *x;                 // [nt*64] solution
    double *part;              // [npart][4096] partial update sums of tiles with many dependencies (helper tasks)
    int *done_slot, *done_y, *done_x, *done_part;   // epoch flags
    int *counters;             // [0] ticket, [1] exited
    int *flags;                // [0] non-positive pivot (solve() == false)
    long long *trace;          // optional [ntasks][4]: globaltimer at task start / dependencies done / end, SM id (profiling builds of the caller)
};

__device__ __forceinline__ void rs_cp16(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void rs_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void rs_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }
__device__ __forceinline__ int rs_ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rs_st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }

// D(8x8) += A(8x4) * B(4x8), fp64 tensor-core MMA (SASS: DMMA)
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// stage a packed 64x64 tile into an operand buffer (pitch 68), 256 threads, L2-only loads (the tile may have been written by another CTA)
__device__ __forceinline__ void rs_stage_tile(double *dst, const double *src, int tid)
{
#pragma unroll
    for (int q = tid; q < TS * TS / 2; q += 256) { const int r = q >> 5, c2 = (q & 31) * 2; rs_cp16(&dst[r * kOpPitch + c2], &src[r * TS + c2]); }
}

// acc[nb][0..1] (rows 8w + lane/4, cols 8nb + 2(lane%4) + {0,1}) += sign * Aop[r][q] * Bop[c][q] over q = 0..63
template <bool NEG>
__device__ __forceinline__ void rs_gemm_nt(double (&acc)[8][2], const double *Aop, const double *Bop, int warp, int lane)
{
    const int fr = lane >> 2, fk = lane & 3;
    const double *ap = Aop + (8 * warp + fr) * kOpPitch + fk;
    const double *bp = Bop + fr * kOpPitch + fk;
#pragma unroll 4
    for (int k0 = 0; k0 < TS; k0 += 4) {
        double a = ap[k0];
        if (NEG) a = -a;
#pragma unroll
        for (int nb = 0; nb < 8; nb++) dmma884(acc[nb][0], acc[nb][1], a, bp[nb * 8 * kOpPitch + k0]);
    }
}

// Cholesky of the 64x64 tile in T (operand pitch 68, lower part valid) and the inverse of its factor, blocked 8 x 8 on DMMA.
//   for block column b = 0..7:
//     panel    : warps 0..6-b   P_i = T(i,b) inv(D_b)^T                                  (one 8x8x8 DMMA product each)
//     trailing : T(i,j) -= P_i P_j^T for b < j <= i; warp 0 takes block (b+1, b+1) FIRST and then factors + inverts it (8 pivots in registers, row r of the
//                block in lanes r, r+8, r+16, r+24; column broadcasts by shuffle) while warps 1..7 do the rest of the trailing update (look-ahead: the
//                serial 8-pivot chain of the next diagonal block runs beside the update instead of behind it)
//   two barriers per block column (16 in all; the former right-looking scalar version needed one per pivot, 64, with every thread touching 16 columns each time:
//   21 us per tile on the critical path of the elimination tree, measured with the task trace).
//   inverse: the diagonal 8x8 inverses come out of the factor steps; off-diagonal blocks by block distance d = 1..7,
//            X[i][j] = -inv(D_i) sum_{k=j..i-1} L[i][k] X[k][j], one warp per block.
// T / Xi: operand-pitch (68) buffers, on return T = L (lower block triangle; diagonal blocks with zero upper part), Xi = inv(L);  Sw: [8][64] per-warp
// scratch.  Linv_out (global, row-major 64x64) = inv(L).
__device__ __forceinline__ void rs_factor8_invert(double *Tb, double *Xb, bool &bad, int lane)
{
    // Tb / Xb: the diagonal block in T and in Xi.  All 32 lanes run (lane l mirrors row l & 7) so that every shuffle is warp-uniform.
    constexpr unsigned full = 0xffffffffu;
    const int r = lane & 7;
    // The inverse rides along: lane c also solves column c of inv(L) by forward substitution, x[q] = (delta_qc - sum_{p<q} L[q][p] x[p]) / L[q][q].  Its partial
    // sums s[q] are advanced with the SAME column broadcasts the factor update needs (L[q][j] = lane q's `l`), one extra FMA per shuffle, off the critical path.
    // Critical path per pivot: diagonal -> rsqrt -> scale -> FMA into the next diagonal.  Two things keep the shuffles OFF that path: (1) the column entries are
    // broadcast UNSCALED (A[q][j], final since the previous pivot) while the rsqrt is still running, every lane scales them itself (same expression as the
    // owning lane: bit-identical); (2) every lane tracks all eight running diagonals dd[q] = A[q][q] - sum_p L[q][p]^2 with the same FMAs as the owner.
    double a[8], x[8], sx[8], dd[8];
    const int c = r;
#pragma unroll
    for (int q = 0; q < 8; q++) { a[q] = Tb[r * kOpPitch + q]; sx[q] = 0.0; }
#pragma unroll
    for (int q = 0; q < 8; q++) dd[q] = __shfl_sync(full, a[q], q);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double u[8];
#pragma unroll
        for (int q = j + 1; q < 8; q++) u[q] = __shfl_sync(full, a[j], q);   // A[q][j] before scaling
        double d = dd[j];
        if (!(d > 0.0)) { bad = true; d = 1.0; }
        const double rinv = rsqrt(d);
        const double l = (r == j) ? d * rinv : a[j] * rinv;                // column j of the factor (rows >= j)
        a[j] = l;
        x[j] = (j >= c) ? ((j == c ? 1.0 : 0.0) - sx[j]) * rinv : 0.0;
#pragma unroll
        for (int q = j + 1; q < 8; q++) {
            const double lq = u[q] * rinv;                                   // L[q][j]
            if (q <= r) a[q] = fma(-l, lq, a[q]);
            dd[q] = fma(-lq, lq, dd[q]);
            sx[q] = fma(lq, x[j], sx[q]);
        }
    }
    __syncwarp();                                                        // the mirror lanes read the block above; lanes 0..7 overwrite it
    if (lane < 8) {
#pragma unroll
        for (int q = 0; q < 8; q++) { Tb[r * kOpPitch + q] = q <= r ? a[q] : 0.0; Xb[q * kOpPitch + c] = x[q]; }
    }
}

__device__ __forceinline__ void rs_factor_invert(double *T, double *Xi, double *Sw, double *Linv_out, int *flags, int tid)
{
    const int warp = tid >> 5, lane = tid & 31, fr = lane >> 2, fk = lane & 3;
    bool bad = false;
    // warp 0 factors the first diagonal block while the others clear Xi (everything but that block, which warp 0 writes in full)
    if (warp == 0) rs_factor8_invert(T, Xi, bad, lane);
    else for (int q = tid - 32; q < TS * kOpPitch; q += 224) { const int rr = q / kOpPitch, cc = q - rr * kOpPitch; if (rr >= 8 || cc >= 8) Xi[q] = 0.0; }
    __syncthreads();
    for (int b = 0; b < 8; b++) {
        const int m = 7 - b;                                                 // block rows below the diagonal block
        // ---- panel: P_i = T(i,b) inv(D_b)^T, i = b + 1 + warp
        if (warp < m) {
            const int i = b + 1 + warp;
            double *Ab = T + (8 * i + fr) * kOpPitch + 8 * b;
            const double *Db = Xi + (8 * b + fr) * kOpPitch + 8 * b;
            const double a0 = Ab[fk], a1 = Ab[4 + fk];
            double c0 = 0.0, c1 = 0.0;
            dmma884(c0, c1, a0, Db[fk]);
            dmma884(c0, c1, a1, Db[4 + fk]);
            __syncwarp();
            Ab[2 * fk] = c0; Ab[2 * fk + 1] = c1;
        }
        __syncthreads();
        if (b == 7) break;
        // ---- trailing update T(i,j) -= P_i P_j^T, b < j <= i <= 7; block (b+1, b+1) goes to warp 0, which then factors it (look-ahead)
        if (warp == 0) {
            const int i = b + 1;
            double *Cb = T + (8 * i + fr) * kOpPitch + 8 * i + 2 * fk;
            const double *Pi = T + (8 * i + fr) * kOpPitch + 8 * b;
            double c0 = Cb[0], c1 = Cb[1];
            dmma884(c0, c1, -Pi[fk], Pi[fk]);
            dmma884(c0, c1, -Pi[4 + fk], Pi[4 + fk]);
            Cb[0] = c0; Cb[1] = c1;
            __syncwarp();
            rs_factor8_invert(T + (8 * i) * kOpPitch + 8 * i, Xi + (8 * i) * kOpPitch + 8 * i, bad, lane);
        } else {
            int cnt = -1;                                                    // blocks after the first, dealt round-robin to warps 1..7
            for (int i = b + 1; i < 8; i++)
                for (int j = b + 1; j <= i; j++, cnt++) {
                    if (cnt < 0 || cnt % 7 != warp - 1) continue;
                    double *Cb = T + (8 * i + fr) * kOpPitch + 8 * j + 2 * fk;
                    const double *Pi = T + (8 * i + fr) * kOpPitch + 8 * b, *Pj = T + (8 * j + fr) * kOpPitch + 8 * b;
                    double c0 = Cb[0], c1 = Cb[1];
                    dmma884(c0, c1, -Pi[fk], Pj[fk]);
                    dmma884(c0, c1, -Pi[4 + fk], Pj[4 + fk]);
                    Cb[0] = c0; Cb[1] = c1;
                }
        }
        __syncthreads();
    }
    if (__syncthreads_or(bad) && tid == 0) flags[0] = 1;
    // ---- inverse, off-diagonal blocks by distance (the diagonal blocks of Xi are in place)
    double *S = Sw + warp * 64;
    const double *Lb = T;
    for (int dist = 1; dist < 8; dist++) {
        const int j = warp, i = warp + dist;
        if (i < 8) {
            double c0 = 0.0, c1 = 0.0;
            for (int k = j; k < i; k++) {
#pragma unroll
                for (int k0 = 0; k0 < 8; k0 += 4)
                    dmma884(c0, c1, Lb[(8 * i + fr) * kOpPitch + 8 * k + k0 + fk], Xi[(8 * k + k0 + fk) * kOpPitch + 8 * j + fr]);
            }
            S[fr * 8 + 2 * fk] = c0; S[fr * 8 + 2 * fk + 1] = c1;
            __syncwarp();
            double e0 = 0.0, e1 = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < 8; k0 += 4) dmma884(e0, e1, Xi[(8 * i + fr) * kOpPitch + 8 * i + k0 + fk], S[(k0 + fk) * 8 + fr]);
            Xi[(8 * i + fr) * kOpPitch + 8 * j + 2 * fk] = -e0; Xi[(8 * i + fr) * kOpPitch + 8 * j + 2 * fk + 1] = -e1;
        }
        __syncthreads();
    }
    for (int q = tid; q < TS * TS; q += 256) Linv_out[q] = Xi[(q >> 6) * kOpPitch + (q & 63)];
}

__global__ void __launch_bounds__(256, 1)
k_rs_solve(const RsPlan P, const RsBuf B, const LmCtl *__restrict__ ctl, int epoch)
{
    if (ctl && ctl->state == LM_DONE) return;
    extern __shared__ __align__(16) double rs_smem[];
    double *opA[2] = {rs_smem, rs_smem + 2 * TS * kOpPitch};
    double *opB[2] = {rs_smem + TS * kOpPitch, rs_smem + 3 * TS * kOpPitch};
    __shared__ int s_task, s_pref;
    __shared__ double s_vec[TS], s_red[4 * TS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(&B.counters[0], 1);
        __syncthreads();
        const int t = s_task;
        if (t >= P.ntasks) break;
        const RsTask task = P.tasks[t];
        if (B.trace && tid == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); B.trace[4 * t] = g; unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); B.trace[4 * t + 3] = sm; }
        if (task.type == 0 || task.type == 3) {
            // ---- factor tile (i, j): acc = A_ij - sum_k L_ik L_jk^T; tiles with many dependencies get their sum from helper tasks
            // (type 3: partial sums of a chunk of the dependencies into a scratch tile), subtracted here in fixed order
            const bool helper = task.type == 3;
            double acc[8][2];
            if (helper) {
#pragma unroll
                for (int nb = 0; nb < 8; nb++) acc[nb][0] = acc[nb][1] = 0.0;
            } else {
                const double *At = B.A + (size_t)task.slot * TS2 + (8 * warp + fr) * TS + 2 * fk;
#pragma unroll
                for (int nb = 0; nb < 8; nb++) { const double2 v = __ldcg(reinterpret_cast<const double2 *>(At + 8 * nb)); acc[nb][0] = v.x; acc[nb][1] = v.y; }
                for (int pp = 0; pp < task.npart; pp++) {
                    if (tid == 0) { while (rs_ld_acquire(&B.done_part[task.part0 + pp]) != epoch) { } }
                    __syncthreads();
                    const double *Pt = B.part + (size_t)(task.part0 + pp) * TS2 + (8 * warp + fr) * TS + 2 * fk;
#pragma unroll
                    for (int nb = 0; nb < 8; nb++) { const double2 v = __ldcg(reinterpret_cast<const double2 *>(Pt + 8 * nb)); acc[nb][0] -= v.x; acc[nb][1] -= v.y; }
                }
            }
            const bool diag = task.i == task.j;
            int staged = -1;                                   // dependency already in flight into buffer (staged & 1)
            for (int d = 0; d < task.ndep; d++) {
                const int2 dep = P.deps[task.dep0 + d];
                const int buf = d & 1;
                if (staged != d) {
                    if (tid == 0) { while (rs_ld_acquire(&B.done_slot[dep.x]) != epoch) { } while (rs_ld_acquire(&B.done_slot[dep.y]) != epoch) { } }
                    __syncthreads();
                    rs_stage_tile(opA[buf], B.L + (size_t)dep.x * TS2, tid);
                    if (!diag) rs_stage_tile(opB[buf], B.L + (size_t)dep.y * TS2, tid);
                    rs_commit();
                }
                rs_wait_all();
                __syncthreads();
                // prefetch the next dependency if it is already there (a non-blocking look at its flags)
                if (d + 1 < task.ndep) {
                    const int2 nx = P.deps[task.dep0 + d + 1];
                    if (tid == 0) s_pref = (rs_ld_acquire(&B.done_slot[nx.x]) == epoch && rs_ld_acquire(&B.done_slot[nx.y]) == epoch) ? 1 : 0;
                    __syncthreads();
                    if (s_pref) {
                        rs_stage_tile(opA[buf ^ 1], B.L + (size_t)nx.x * TS2, tid);
                        if (!diag) rs_stage_tile(opB[buf ^ 1], B.L + (size_t)nx.y * TS2, tid);
                        rs_commit();
                        staged = d + 1;
                    }
                }
                if (helper) rs_gemm_nt<false>(acc, opA[buf], diag ? opA[buf] : opB[buf], warp, lane);
                else rs_gemm_nt<true>(acc, opA[buf], diag ? opA[buf] : opB[buf], warp, lane);
            }
            __syncthreads();
            if (helper) {
                double *Pt = B.part + (size_t)task.slot * TS2 + (8 * warp + fr) * TS + 2 * fk;
#pragma unroll
                for (int nb = 0; nb < 8; nb++) *reinterpret_cast<double2 *>(Pt + 8 * nb) = make_double2(acc[nb][0], acc[nb][1]);
                __threadfence();
                __syncthreads();
                if (tid == 0) rs_st_release(&B.done_part[task.slot], epoch);
                if (B.trace && tid == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); B.trace[4 * t + 1] = g; B.trace[4 * t + 2] = g; }
                continue;
            }
            if (B.trace && tid == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); B.trace[4 * t + 1] = g; }
            if (diag) {
                double *T = rs_smem;                                           // operand-pitch view of the first buffer
#pragma unroll
                for (int nb = 0; nb < 8; nb++) { double *p = &T[(8 * warp + fr) * kOpPitch + 8 * nb + 2 * fk]; p[0] = acc[nb][0]; p[1] = acc[nb][1]; }
                __syncthreads();
                rs_factor_invert(T, rs_smem + TS * kOpPitch, rs_smem + 2 * TS * kOpPitch, B.Linv + (size_t)task.i * TS2, B.flags, tid);
            } else {
                // L_ij = acc * inv(L_jj)^T
#pragma unroll
                for (int nb = 0; nb < 8; nb++) { double *p = &opA[0][(8 * warp + fr) * kOpPitch + 8 * nb + 2 * fk]; p[0] = acc[nb][0]; p[1] = acc[nb][1]; }
                if (tid == 0) { while (rs_ld_acquire(&B.done_slot[task.diag]) != epoch) { } }
                __syncthreads();
                rs_stage_tile(opB[0], B.Linv + (size_t)task.j * TS2, tid);
                rs_commit(); rs_wait_all();
                __syncthreads();
                double out[8][2];
#pragma unroll
                for (int nb = 0; nb < 8; nb++) out[nb][0] = out[nb][1] = 0.0;
                rs_gemm_nt<false>(out, opA[0], opB[0], warp, lane);
                double *Lt = B.L + (size_t)task.slot * TS2 + (8 * warp + fr) * TS + 2 * fk;
#pragma unroll
                for (int nb = 0; nb < 8; nb++) *reinterpret_cast<double2 *>(Lt + 8 * nb) = make_double2(out[nb][0], out[nb][1]);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) rs_st_release(&B.done_slot[task.slot], epoch);
        } else if (task.type == 1) {
            // ---- forward row i: y_i = inv(L_ii) (b_i - sum_k L_ik y_k); thread (r = tid / 4, quarter = tid % 4)
            const int r = tid >> 2, qt = tid & 3;
            double sum = 0.0;
            for (int d = 0; d < task.ndep; d++) {
                const int2 dep = P.deps[task.dep0 + d];
                if (tid == 0) { while (rs_ld_acquire(&B.done_y[dep.y]) != epoch) { } while (rs_ld_acquire(&B.done_slot[dep.x]) != epoch) { } }
                __syncthreads();
                if (tid < TS) s_vec[tid] = __ldcg(&B.y[dep.y * TS + tid]);
                __syncthreads();
                const double *row = B.L + (size_t)dep.x * TS2 + r * TS + 16 * qt;
#pragma unroll
                for (int c = 0; c < 16; c += 2) { const double2 v = __ldcg(reinterpret_cast<const double2 *>(row + c)); sum = fma(v.x, s_vec[16 * qt + c], sum); sum = fma(v.y, s_vec[16 * qt + c + 1], sum); }
                __syncthreads();
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (tid == 0) { while (rs_ld_acquire(&B.done_slot[task.diag]) != epoch) { } }
            if (qt == 0) s_vec[r] = B.b[task.i * TS + r] - sum;
            __syncthreads();
            const double *Li = B.Linv + (size_t)task.i * TS2 + r * TS + 16 * qt;
            double s2 = 0.0;
#pragma unroll
            for (int c = 0; c < 16; c += 2) { const double2 v = __ldcg(reinterpret_cast<const double2 *>(Li + c)); s2 = fma(v.x, s_vec[16 * qt + c], s2); s2 = fma(v.y, s_vec[16 * qt + c + 1], s2); }
            s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            if (qt == 0) B.y[task.i * TS + r] = s2;
            __threadfence();
            __syncthreads();
            if (tid == 0) rs_st_release(&B.done_y[task.i], epoch);
        } else {
            // ---- backward row j: x_j = inv(L_jj)^T (y_j - sum_i L_ij^T x_i); thread (c = tid % 64, part = tid / 64)
            const int c = tid & 63, part = tid >> 6;
            double sum = 0.0;
            for (int d = 0; d < task.ndep; d++) {
                const int2 dep = P.deps[task.dep0 + d];
                if (tid == 0) { while (rs_ld_acquire(&B.done_x[dep.y]) != epoch) { } }
                __syncthreads();
                if (tid < TS) s_vec[tid] = __ldcg(&B.x[dep.y * TS + tid]);
                __syncthreads();
                const double *col = B.L + (size_t)dep.x * TS2 + (16 * part) * TS + c;
#pragma unroll
                for (int r = 0; r < 16; r++) sum = fma(__ldcg(col + r * TS), s_vec[16 * part + r], sum);
                __syncthreads();
            }
            if (tid == 0) { while (rs_ld_acquire(&B.done_y[task.j]) != epoch) { } }
            s_red[part * TS + c] = sum;
            __syncthreads();
            if (tid < TS) s_vec[tid] = __ldcg(&B.y[task.j * TS + tid]) - (s_red[tid] + s_red[TS + tid] + s_red[2 * TS + tid] + s_red[3 * TS + tid]);
            __syncthreads();
            const double *Li = B.Linv + (size_t)task.j * TS2 + (16 * part) * TS + c;
            double s2 = 0.0;
#pragma unroll
            for (int r = 0; r < 16; r++) s2 = fma(__ldcg(Li + r * TS), s_vec[16 * part + r], s2);
            __syncthreads();
            s_red[part * TS + c] = s2;
            __syncthreads();
            if (tid < TS) B.x[task.j * TS + tid] = s_red[tid] + s_red[TS + tid] + s_red[2 * TS + tid] + s_red[3 * TS + tid];
            __threadfence();
            __syncthreads();
            if (tid == 0) rs_st_release(&B.done_x[task.j], epoch);
        }
        if (B.trace && tid == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); B.trace[4 * t + 2] = g; }
    }
    if (tid == 0) {
        __threadfence();
        const int prev = atomicAdd(&B.counters[1], 1);
        if (prev == (int)gridDim.x - 1) { B.counters[0] = 0; B.counters[1] = 0; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Host side: elimination order, symbolic factorisation, slots and the task list.
struct TilePlanHost {
    static constexpr int kChunk = 2, kMaxPart = 4096;      // helper tasks: dependencies per chunk, cap on scratch tiles (128 MB)
    int nt = 0, ns = 0, nlevels = 0, n_part = 0;
    std::vector<int> pos;            // group -> tile row in elimination order
    std::vector<int> slot_of;        // [nt*nt] slot of tile (i, j), i >= j, or -1
    std::vector<int2> slot_tile;     // [ns] (i, j)
    std::vector<RsTask> tasks;
    std::vector<int2> deps;

    // Nested-dissection order of the tile groups by recursive bisection of the natural (temporal) order: the separator of [lo, hi)
    // is the run of groups right of the middle that the left half reaches; the halves are then independent.
    static void nd_order(const std::vector<uint8_t> &adj, int ng, int lo, int hi, std::vector<int> &out)
    {
        const int n = hi - lo;
        if (n <= 2) { for (int g = lo; g < hi; g++) out.push_back(g); return; }     // (three in natural order are a chain of three; middle-last is two levels)
        const int mid = lo + n / 2;
        int reach = mid - 1;
        for (int g = lo; g < mid; g++)
            for (int h2 = hi - 1; h2 > reach; h2--) if (adj[(size_t)g * ng + h2]) { reach = h2; break; }
        const int w = reach - mid + 1;
        if (w == 0) { nd_order(adj, ng, lo, mid, out); nd_order(adj, ng, mid, hi, out); return; }
        if (2 * w >= n || mid + w >= hi) { for (int g = lo; g < hi; g++) out.push_back(g); return; }   // no useful separator
        nd_order(adj, ng, lo, mid, out);
        nd_order(adj, ng, mid + w, hi, out);
        for (int g = mid; g < mid + w; g++) out.push_back(g);
    }

    // adj[ng x ng]: symmetric coupling of the tile groups
    void build(const std::vector<uint8_t> &adj, int ng)
    {
        nt = ng;
        std::vector<int> order; order.reserve(nt);
        nd_order(adj, ng, 0, ng, order);
        pos.assign(ng, 0);
        for (int t = 0; t < nt; t++) pos[order[t]] = t;
        std::vector<uint8_t> pat((size_t)nt * nt, 0);
        for (int a = 0; a < ng; a++)
            for (int b = 0; b < ng; b++) if (a != b && adj[(size_t)a * ng + b]) { const int i = std::max(pos[a], pos[b]), j = std::min(pos[a], pos[b]); pat[(size_t)i * nt + j] = 1; }
        // symbolic factorisation (fill-in), column by column
        std::vector<std::vector<int>> rows(nt), cols(nt);
        for (int k = 0; k < nt; k++) {
            for (int i = k + 1; i < nt; i++) if (pat[(size_t)i * nt + k]) rows[k].push_back(i);
            for (size_t x = 0; x < rows[k].size(); x++) for (size_t y = 0; y < x; y++) pat[(size_t)rows[k][x] * nt + rows[k][y]] = 1;
        }
        std::vector<int> level(nt, 0);
        nlevels = 0;
        for (int i = 0; i < nt; i++) {
            int lv = 0;
            for (int k = 0; k < i; k++) if (pat[(size_t)i * nt + k]) { cols[i].push_back(k); lv = std::max(lv, level[k] + 1); }
            level[i] = lv; nlevels = std::max(nlevels, lv + 1);
        }
        // slots: diagonal tiles first (slot i = tile (i, i)), then the off-diagonal tiles column by column
        slot_of.assign((size_t)nt * nt, -1); slot_tile.clear();
        for (int i = 0; i < nt; i++) { slot_of[(size_t)i * nt + i] = i; slot_tile.push_back(make_int2(i, i)); }
        for (int k = 0; k < nt; k++) for (int i : rows[k]) { slot_of[(size_t)i * nt + k] = (int)slot_tile.size(); slot_tile.push_back(make_int2(i, k)); }
        ns = (int)slot_tile.size();
        // tasks: factor tasks by (level of column, diagonal first), then forward rows ascending, then backward rows descending
        tasks.clear(); deps.clear();
        std::vector<int> by_level(nt);
        for (int k = 0; k < nt; k++) by_level[k] = k;
        std::stable_sort(by_level.begin(), by_level.end(), [&](int a, int b) { return level[a] < level[b]; });
        n_part = 0;
        auto factor_task = [&](int i, int j) {
            RsTask t = {}; t.type = 0; t.i = i; t.j = j; t.slot = slot_of[(size_t)i * nt + j]; t.diag = j; t.dep0 = (int)deps.size();
            for (int k : cols[j]) if (pat[(size_t)i * nt + k] || i == j) deps.push_back(make_int2(slot_of[(size_t)i * nt + k], slot_of[(size_t)j * nt + k]));
            t.ndep = (int)deps.size() - t.dep0;
            // a tile with many dependencies would serialise ~2 us GEMMs in one CTA (the root separator of the elimination tree): hand
            // chunks of kChunk dependencies to helper tasks on other SMs, the owner subtracts their partial sums in chunk order
            if (t.ndep > kChunk && n_part + (t.ndep + kChunk - 1) / kChunk <= kMaxPart) {
                t.part0 = n_part;
                for (int d0 = 0; d0 < t.ndep; d0 += kChunk) {
                    RsTask hlp = {}; hlp.type = 3; hlp.i = i; hlp.j = j; hlp.slot = n_part++; hlp.diag = j; hlp.dep0 = t.dep0 + d0; hlp.ndep = std::min(kChunk, t.ndep - d0);
                    tasks.push_back(hlp);
                }
                t.npart = n_part - t.part0; t.ndep = 0;
            }
            tasks.push_back(t);
        };
        {   // per level: the diagonal tasks of its columns, then their panels
            int a = 0;
            while (a < nt) {
                int b = a;
                while (b < nt && level[by_level[b]] == level[by_level[a]]) b++;
                for (int q = a; q < b; q++) factor_task(by_level[q], by_level[q]);
                for (int q = a; q < b; q++) for (int i : rows[by_level[q]]) factor_task(i, by_level[q]);
                a = b;
            }
        }
        for (int i = 0; i < nt; i++) {
            RsTask t = {}; t.type = 1; t.i = i; t.j = i; t.slot = i; t.diag = i; t.dep0 = (int)deps.size();
            for (int k : cols[i]) deps.push_back(make_int2(slot_of[(size_t)i * nt + k], k));
            t.ndep = (int)deps.size() - t.dep0;
            tasks.push_back(t);
        }
        for (int j = nt - 1; j >= 0; j--) {
            RsTask t = {}; t.type = 2; t.i = j; t.j = j; t.slot = j; t.diag = j; t.dep0 = (int)deps.size();
            for (size_t x = rows[j].size(); x-- > 0;) deps.push_back(make_int2(slot_of[(size_t)rows[j][x] * nt + j], rows[j][x]));
            t.ndep = (int)deps.size() - t.dep0;
            tasks.push_back(t);
        }
    }
};

}  // namespace orbs
