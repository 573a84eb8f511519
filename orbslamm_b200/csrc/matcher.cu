// matcher.cu -- ORBmatcher hot path on sm_100a (orbm_*): Hamming distance, last-frame projection, and the
// grid-cell projection search with the reference's sequential claim semantics.
//
// Replaces S/src/ORBmatcher.cc:45-137, 1330-1472, 1603-1665 and S/src/Frame.cc:230-245, 327-392.
// Pure integer work (XOR + POPC over 8 x u32) plus fp32 window arithmetic with explicit rounding.
#include <algorithm>
#include <vector>
#include <mutex>
#include "common.cuh"

namespace orbs {

constexpr int kGridCols = ORBS_FRAME_GRID_COLS, kGridRows = ORBS_FRAME_GRID_ROWS, kGridCells = kGridCols * kGridRows;
constexpr int kHisto = ORBM_HISTO_LENGTH;

// win_x / win_y: origin GetFeaturesInArea subtracts (= min_x / min_y for a Frame; a KeyFrame's integer mnMinX / mnMinY, KeyFrame.cc:623-635)
struct GridParams { float min_x, min_y, max_x, max_y, w_inv, h_inv, win_x, win_y; };

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// all-pairs DescriptorDistance: one thread per (i, j)
__global__ void __launch_bounds__(256)
k_hamming_pairs(const uint4 *__restrict__ a, int n, const uint4 *__restrict__ b, int m, int *__restrict__ out)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * m) return;
    const int i = (int)(t / m), j = (int)(t - (size_t)i * m);
    out[t] = hamming256(__ldg(&a[2 * i]), __ldg(&a[2 * i + 1]), __ldg(&b[2 * j]), __ldg(&b[2 * j + 1]));
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:242-307) for a batch of map points: the descriptor with the least MEDIAN distance to the point's other
// observations (`vDists[0.5 * (N - 1)]` of the sorted row, the zero self-distance included; the first row wins ties).  One warp per point, lanes over the rows;
// a row's median is found by bisection on the distance value (9 counting passes over the row: the smallest v with |{j : d_ij <= v}| >= k + 1), which needs no
// per-row storage however many observations a point has.  The reference does this on the host with an N x N float array on the stack per call.
__global__ void __launch_bounds__(256)
k_distinctive(int n_points, const uint4 *__restrict__ desc, const int *__restrict__ start, int *__restrict__ best_idx)
{
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= n_points) return;
    const int b = start[p], N = start[p + 1] - b;
    if (N <= 0) { if (lane == 0) best_idx[p] = -1; return; }
    const int kth = (N - 1) >> 1;                                          // (int)(0.5 * (N - 1))
    unsigned best = 0xffffffffu;                                           // median << 16 | row: minimum = least median, first row
    for (int i = lane; i < N; i += 32) {
        const uint4 a0 = __ldg(&desc[2 * (size_t)(b + i)]), a1 = __ldg(&desc[2 * (size_t)(b + i) + 1]);
        int lo = 0, hi = 256;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int c = 0;
            for (int j = 0; j < N; j++) c += hamming256(a0, a1, __ldg(&desc[2 * (size_t)(b + j)]), __ldg(&desc[2 * (size_t)(b + j) + 1])) <= mid;
            if (c >= kth + 1) hi = mid; else lo = mid + 1;
        }
        best = min(best, ((unsigned)lo << 16) | (unsigned)min(i, 0xffff));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (lane == 0) best_idx[p] = (int)(best & 0xffffu);
}

// ORBmatcher.cc:1354-1391, one thread per (frame, last-frame slot)
__global__ void __launch_bounds__(256)
k_project_last(int q_slab, const float *__restrict__ Tcw, float fx, float fy, float cx, float cy, GridParams g,
               const float *__restrict__ scale_factors, int nlevels, const float *__restrict__ Xw,
               const int *__restrict__ last_octave, const int *__restrict__ q_counts, float th,
               uint8_t *__restrict__ q_valid, float2 *__restrict__ q_uv, float *__restrict__ q_radius,
               int *__restrict__ q_minl, int *__restrict__ q_maxl)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q_counts[f]) return;
    const size_t q = (size_t)f * q_slab + i;
    if (!q_valid[q]) return;
    const float *T = Tcw + 16 * f;
    const float X = Xw[3 * q], Y = Xw[3 * q + 1], Z = Xw[3 * q + 2];
    float xc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float s = __fmul_rn(T[4 * r], X);
        s = __fadd_rn(s, __fmul_rn(T[4 * r + 1], Y));
        s = __fadd_rn(s, __fmul_rn(T[4 * r + 2], Z));
        xc[r] = __fadd_rn(s, T[4 * r + 3]);
    }
    const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)xc[2]));
    bool ok = !(invzc < 0);
    const float u = __fadd_rn(__fmul_rn(__fmul_rn(fx, xc[0]), invzc), cx);
    const float v = __fadd_rn(__fmul_rn(__fmul_rn(fy, xc[1]), invzc), cy);
    if (u < g.min_x || u > g.max_x) ok = false;
    if (v < g.min_y || v > g.max_y) ok = false;
    int oct = last_octave[q];
    if (ok) {
        oct = min(max(oct, 0), nlevels - 1);
        q_uv[q] = make_float2(u, v);
        q_radius[q] = __fmul_rn(th, scale_factors[oct]);
        q_minl[q] = oct - 1;
        q_maxl[q] = oct + 1;
    }
    q_valid[q] = ok ? 1 : 0;
}

// Frame::UndistortKeyPoints (Frame.cc:404-434) = cv::undistortPoints(pts, K, dist, R = I, P = K): 5 fixed-point iterations in
// fp64 (cvUndistortPointsInternal, default criteria), result rounded to fp32.  One thread per keypoint.
__global__ void __launch_bounds__(256)
k_undistort(int slab, const int *__restrict__ counts, const float2 *__restrict__ xy, float2 *__restrict__ xy_un,
            double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2, double k3)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[f]) return;
    const size_t q = (size_t)f * slab + i;
    const float2 p = xy[q];
    if (k1 == 0.0) { xy_un[q] = p; return; }                     // Frame.cc:406-410
    const double ifx = 1. / fx, ify = 1. / fy;
    double x = p.x, y = p.y;
    const double u = x, v = y;
    x = (x - cx) * ifx; y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = 1. / (1. + ((k3 * r2 + k2) * r2 + k1) * r2);
        if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
        const double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
        const double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - dX) * icdist; y = (y0 - dY) * icdist;
    }
    xy_un[q] = make_float2((float)((fx * x + cx) * 1.0), (float)((fy * y + cy) * 1.0));
}

// Frame::isInFrustum (Frame.cc:269-325) + MapPoint::PredictScale (MapPoint.cc:385-394), one thread per map point.
// cv::Mat algebra restated as in k_project_last (small gemm: fp32, left to right); cv::norm / Mat::dot accumulate in double.
__global__ void __launch_bounds__(256)
k_in_frustum(int slab, const int *__restrict__ counts, const float *__restrict__ Tcw, const float *__restrict__ Ow,
             float fx, float fy, float cx, float cy, GridParams g, float log_sf, float cos_limit,
             const float *__restrict__ Xw, const float *__restrict__ normal, const float *__restrict__ mf_min, const float *__restrict__ mf_max,
             uint8_t *__restrict__ in_view, float2 *__restrict__ proj, int *__restrict__ level, float *__restrict__ view_cos)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[f]) return;
    const size_t q = (size_t)f * slab + i;
    in_view[q] = 0;
    const float *T = Tcw + 16 * f;
    const float X = Xw[3 * q], Y = Xw[3 * q + 1], Z = Xw[3 * q + 2];
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float s = __fmul_rn(T[4 * r], X);
        s = __fadd_rn(s, __fmul_rn(T[4 * r + 1], Y));
        s = __fadd_rn(s, __fmul_rn(T[4 * r + 2], Z));
        pc[r] = __fadd_rn(s, T[4 * r + 3]);
    }
    if (pc[2] < 0.0f) return;
    const float invz = __fdiv_rn(1.0f, pc[2]);
    const float u = __fadd_rn(__fmul_rn(__fmul_rn(fx, pc[0]), invz), cx);
    const float v = __fadd_rn(__fmul_rn(__fmul_rn(fy, pc[1]), invz), cy);
    if (u < g.min_x || u > g.max_x) return;
    if (v < g.min_y || v > g.max_y) return;
    const float maxd = __fmul_rn(1.2f, mf_max[q]), mind = __fmul_rn(0.8f, mf_min[q]);
    const float po[3] = {__fsub_rn(X, Ow[3 * f]), __fsub_rn(Y, Ow[3 * f + 1]), __fsub_rn(Z, Ow[3 * f + 2])};
    double s2 = 0, dot = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) { s2 = __dadd_rn(s2, __dmul_rn((double)po[k], (double)po[k])); dot = __dadd_rn(dot, __dmul_rn((double)po[k], (double)normal[3 * q + k])); }
    const float dist = (float)sqrt(s2);
    if (dist < mind || dist > maxd) return;
    const float vc = (float)(dot / (double)dist);
    if (vc < cos_limit) return;
    const float ratio = __fdiv_rn(mf_max[q], dist);
    in_view[q] = 1;
    proj[q] = make_float2(u, v);
    // std::log(float) / float, std::ceil(float): log evaluated in double and rounded to float (= a correctly rounded logf)
    level[q] = (int)ceilf(__fdiv_rn((float)log((double)ratio), log_sf));
    view_cos[q] = vc;
}

// Projection half of the KeyFrame / Sim3 search family (ORBmatcher.cc:292-405, 827-977, 979-1102, 1104-1328, 1474-1601), one thread per
// (view, map point).  cv::Mat algebra as in k_project_last (small gemm: fp32, left to right, no FMA); cv::norm / Mat::dot in double.
struct ProjView {            // = orbm_projection (include/orbslamm_b200.h)
    float R[9], t[3], R2[9], t2[3], Ow[3];
    float fx, fy, cx, cy, min_x, min_y, max_x, max_y, log_sf, th;
    int flags;
};
static_assert(sizeof(ProjView) == sizeof(orbm_projection), "ProjView must mirror orbm_projection");

__device__ __forceinline__ void rt_apply(const float *R, const float *t, const float *x, float *out)
{
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float s = __fmul_rn(R[3 * r], x[0]);
        s = __fadd_rn(s, __fmul_rn(R[3 * r + 1], x[1]));
        s = __fadd_rn(s, __fmul_rn(R[3 * r + 2], x[2]));
        out[r] = __fadd_rn(s, t[r]);
    }
}

__global__ void __launch_bounds__(256)
k_project_points(int q_slab, const ProjView *__restrict__ views, const float *__restrict__ scale_factors, int nlevels,
                 const float *__restrict__ Xw, const float *__restrict__ normal, const float *__restrict__ mf_min, const float *__restrict__ mf_max,
                 const int *__restrict__ q_counts, uint8_t *__restrict__ q_valid, float2 *__restrict__ q_uv, float *__restrict__ q_radius,
                 int *__restrict__ q_minl, int *__restrict__ q_maxl, int *__restrict__ q_level)
{
    __shared__ ProjView V;
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = threadIdx.x; k < (int)(sizeof(ProjView) / 4); k += blockDim.x) reinterpret_cast<int *>(&V)[k] = reinterpret_cast<const int *>(views + f)[k];
    __syncthreads();
    if (i >= q_counts[f]) return;
    const size_t q = (size_t)f * q_slab + i;
    if (!q_valid[q]) return;
    q_valid[q] = 0;
    const int flags = V.flags;
    const float X[3] = {Xw[3 * q], Xw[3 * q + 1], Xw[3 * q + 2]};
    float pc[3];
    rt_apply(V.R, V.t, X, pc);
    if (flags & ORBM_PROJ_TWO_STEP) { float p1[3] = {pc[0], pc[1], pc[2]}; rt_apply(V.R2, V.t2, p1, pc); }
    if (!(flags & ORBM_PROJ_NO_DEPTH) && pc[2] < 0.0f) return;
    const float invz = __fdiv_rn(1.0f, pc[2]);          // 1/z in float == (float)(1.0/z) in double: division is innocuous under double rounding
    float u, v;
    if (flags & ORBM_PROJ_FRAME_UV) {
        u = __fadd_rn(__fmul_rn(__fmul_rn(V.fx, pc[0]), invz), V.cx);
        v = __fadd_rn(__fmul_rn(__fmul_rn(V.fy, pc[1]), invz), V.cy);
    } else {
        u = __fadd_rn(__fmul_rn(V.fx, __fmul_rn(pc[0], invz)), V.cx);
        v = __fadd_rn(__fmul_rn(V.fy, __fmul_rn(pc[1], invz)), V.cy);
    }
    if (flags & ORBM_PROJ_FRAME_BOUNDS) {
        if (u < V.min_x || u > V.max_x) return;
        if (v < V.min_y || v > V.max_y) return;
    } else if (!(u >= V.min_x && u < V.max_x && v >= V.min_y && v < V.max_y)) return;
    const float maxd = __fmul_rn(1.2f, mf_max[q]), mind = __fmul_rn(0.8f, mf_min[q]);
    float po[3];
    if (flags & ORBM_PROJ_DIST_CAMERA) { po[0] = pc[0]; po[1] = pc[1]; po[2] = pc[2]; }
    else { po[0] = __fsub_rn(X[0], V.Ow[0]); po[1] = __fsub_rn(X[1], V.Ow[1]); po[2] = __fsub_rn(X[2], V.Ow[2]); }
    double s2 = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) s2 = __dadd_rn(s2, __dmul_rn((double)po[k], (double)po[k]));
    const float dist = (float)sqrt(s2);
    if (dist < mind || dist > maxd) return;
    if (flags & ORBM_PROJ_CHECK_NORMAL) {
        double dot = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) dot = __dadd_rn(dot, __dmul_rn((double)po[k], (double)normal[3 * q + k]));
        if (dot < __dmul_rn(0.5, (double)dist)) return;
    }
    const float ratio = __fdiv_rn(mf_max[q], dist);
    const int level = (int)ceilf(__fdiv_rn((float)log((double)ratio), V.log_sf));
    if (level < 0 || level >= nlevels) return;       // the reference indexes mvScaleFactors out of range here (undefined)
    q_uv[q] = make_float2(u, v);
    q_radius[q] = __fmul_rn(V.th, scale_factors[level]);
    q_minl[q] = level - 1;
    q_maxl[q] = (flags & ORBM_PROJ_LEVEL_PLUS1) ? level + 1 : level;
    if (q_level) q_level[q] = level;
    q_valid[q] = 1;
}

// Frame::AssignFeaturesToGrid / PosInGrid: CSR per frame, cell = ix * 48 + iy.  Order inside a cell is not
// preserved; the search re-creates the reference's visiting order through an explicit (ix, iy, index) key.
__global__ void __launch_bounds__(512)
k_grid_build(int f_slab, GridParams g, const float2 *__restrict__ f_xy, const int *__restrict__ f_counts,
             int *__restrict__ cell_start /*[n_frames, cells+1]*/, int *__restrict__ cell_items /*[n_frames, f_slab]*/)
{
    __shared__ int cnt[kGridCells];
    __shared__ int warp_tot[16];
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int N = f_counts[f];
    const float2 *xy = f_xy + (size_t)f * f_slab;
    for (int c = tid; c < kGridCells; c += nt) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += nt) {
        const float2 p = xy[i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(p.x, g.min_x), g.w_inv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(p.y, g.min_y), g.h_inv));
        if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) atomicAdd(&cnt[px * kGridRows + py], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counts: 6 per thread (512 threads)
    constexpr int PER = kGridCells / 512;
    int loc[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { loc[k] = cnt[tid * PER + k]; sum += loc[k]; }
    int inc = sum;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int v = lane < 16 ? warp_tot[lane] : 0, w = v;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += o; }
        if (lane < 16) warp_tot[lane] = w - v;
    }
    __syncthreads();
    int run = warp_tot[wid] + inc - sum;
    int *cs = cell_start + (size_t)f * (kGridCells + 1);
#pragma unroll
    for (int k = 0; k < PER; k++) { cs[tid * PER + k] = run; cnt[tid * PER + k] = run; run += loc[k]; }
    if (tid == nt - 1) cs[kGridCells] = run;
    __syncthreads();
    int *items = cell_items + (size_t)f * f_slab;
    for (int i = tid; i < N; i += nt) {
        const float2 p = xy[i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(p.x, g.min_x), g.w_inv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(p.y, g.min_y), g.h_inv));
        if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) items[atomicAdd(&cnt[px * kGridRows + py], 1)] = i;
    }
}

// Projection search in two kernels.
//
// The reference loop is sequential over queries: an accepted query claims its feature and later queries skip
// it.  Equivalent fixed point: query i may take feature k unless a query j < i took it.
//   k_search_candidates  (one warp per query, whole chip): the kTop best candidates of every query in the
//       reference's comparison order (distance, then GetFeaturesInArea visiting order = (ix, iy, index)).
//   k_search_resolve     (one CTA per frame): rounds -- every query proposes the first feature of its list that
//       is not owned by a lower-index query (ownership from the previous round); ownership = lowest proposing
//       query; repeat until no proposal changes.  After round r the first r queries are final, so the loop ends
//       with exactly the sequential result.  A query whose list is exhausted although it had more than kTop
//       candidates rescans its window (rare).  Then commit + rotation histogram.
constexpr int kTop = 8;
constexpr unsigned long long kNoKey = ~0ull;

struct SearchArgs {
    int f_slab, q_slab;
    GridParams g;
    const float2 *f_xy; const int *f_octave; const float *f_angle; const uint4 *f_desc; const int *f_counts;
    const uint8_t *q_valid; const float2 *q_uv; const float *q_radius; const int *q_minl; const int *q_maxl;
    const float *q_angle; const uint4 *q_desc; const int *q_counts;
    const int *cell_start; const int *cell_items;
    int th_dist; float ratio; int check_ori;
    int *feat_match; int *nmatches;
    unsigned *top;  // [n_frames, q_slab, kTop]  dist << 20 | feature index, ascending in comparison order; 0xffffffff = none
    int *ncand;     // [n_frames, q_slab] number of candidates in the window
    int *prop;      // [n_frames, q_slab] proposal of every query (feature index or -1)
    int *owner;     // [n_frames, 2, f_slab]
};

struct Window { int c0, c1, r0, r1; bool empty, check; };

__device__ __forceinline__ Window make_window(const GridParams &g, float2 uv, float r, int minl, int maxl)
{
    // Frame::GetFeaturesInArea, Frame.cc:327-380
    Window w;
    w.c0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(uv.x, g.win_x), r), g.w_inv));
    w.c1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(uv.x, g.win_x), r), g.w_inv));
    w.r0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(uv.y, g.win_y), r), g.h_inv));
    w.r1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(uv.y, g.win_y), r), g.h_inv));
    w.c0 = max(w.c0, 0); w.r0 = max(w.r0, 0); w.c1 = min(w.c1, kGridCols - 1); w.r1 = min(w.r1, kGridRows - 1);
    w.empty = w.c0 >= kGridCols || w.c1 < 0 || w.r0 >= kGridRows || w.r1 < 0;
    w.check = (minl > 0) || (maxl >= 0);
    return w;
}

__device__ __forceinline__ bool in_window(const Window &w, int minl, int maxl, float2 uv, float r, int oct, float2 p)
{
    if (w.check) {
        if (oct < minl) return false;
        if (maxl >= 0 && oct > maxl) return false;
    }
    return fabsf(__fsub_rn(p.x, uv.x)) < r && fabsf(__fsub_rn(p.y, uv.y)) < r;
}

// Frame::AssignFeaturesToGrid / GetFeaturesInArea as stand-alone entry points (Frame.cc:230-245, 327-380; KeyFrame.cc:618-657).
// The reference's cells hold their features in ascending index order (push_back in a loop over i); k_grid_build appends unordered, so the
// exported grid is sorted per cell.
__global__ void __launch_bounds__(256)
k_grid_sort_cells(int f_slab, const int *__restrict__ cell_start, int *__restrict__ cell_items)
{
    const int f = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kGridCells) return;
    const int *cs = cell_start + (size_t)f * (kGridCells + 1);
    int *it = cell_items + (size_t)f * f_slab;
    const int b = cs[c], e = cs[c + 1];
    for (int i = b + 1; i < e; i++) {
        const int v = it[i];
        int j = i - 1;
        while (j >= b && it[j] > v) { it[j + 1] = it[j]; j--; }
        it[j + 1] = v;
    }
}

// one warp per query; cells in the reference's order (ix outer, iy inner), the features of a cell in ascending index, appended in order
__global__ void __launch_bounds__(256)
k_features_in_area(int f_slab, int q_slab, int cap, GridParams g, const float2 *__restrict__ f_xy, const int *__restrict__ f_octave,
                   const int *__restrict__ cell_start, const int *__restrict__ cell_items, const float *__restrict__ q_xyr, const int *__restrict__ q_minl,
                   const int *__restrict__ q_maxl, const int *__restrict__ q_counts, int *__restrict__ out_idx, int *__restrict__ out_count)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= q_counts[f]) return;
    const size_t fo = (size_t)f * f_slab, q = (size_t)f * q_slab + i;
    const float2 uv = make_float2(q_xyr[3 * q], q_xyr[3 * q + 1]);
    const float r = q_xyr[3 * q + 2];
    const int minl = q_minl ? q_minl[q] : -1, maxl = q_maxl ? q_maxl[q] : -1;
    const Window w = make_window(g, uv, r, minl, maxl);
    const int *cs = cell_start + (size_t)f * (kGridCells + 1); const int *items = cell_items + fo;
    int *out = out_idx + q * cap;
    int n = 0;
    if (!w.empty)
        for (int ix = w.c0; ix <= w.c1; ix++)
            for (int iy = w.r0; iy <= w.r1; iy++) {
                const int cell = ix * kGridRows + iy, je = cs[cell + 1];
                for (int j0 = cs[cell]; j0 < je; j0 += 32) {
                    const int j = j0 + lane;
                    bool hit = false; int k = -1;
                    if (j < je) { k = items[j]; hit = in_window(w, minl, maxl, uv, r, f_octave ? f_octave[fo + k] : 0, f_xy[fo + k]); }
                    const unsigned bal = __ballot_sync(0xffffffffu, hit);
                    if (hit) { const int pos = n + __popc(bal & ((1u << lane) - 1)); if (pos < cap) out[pos] = k; }
                    n += __popc(bal);
                }
            }
    if (lane == 0) out_count[q] = n;
}

__device__ __forceinline__ unsigned long long make_key(int d, int ix, int iy, int k)
{
    return ((unsigned long long)d << 32) | ((unsigned long long)ix << 26) | ((unsigned long long)iy << 20) | (unsigned long long)k;
}

// Key of a candidate in the reference's comparison order (distance, then GetFeaturesInArea visiting order = (ix, iy, index)).
// K32: f_slab <= 2048 -> the whole key fits 32 bits (9 + 6 + 6 + 11) and the sorted insert / warp merge are single VIMNMX ops.
template <bool K32> struct CandKey;
template <> struct CandKey<true> {
    typedef unsigned T;
    static constexpr T none = 0xffffffffu;
    static __device__ __forceinline__ T make(int d, int ix, int iy, int k) { return ((unsigned)d << 23) | ((unsigned)ix << 17) | ((unsigned)iy << 11) | (unsigned)k; }
    static __device__ __forceinline__ unsigned pack(T m) { return ((m >> 23) << 20) | (m & 0x7ffu); }
};
template <> struct CandKey<false> {
    typedef unsigned long long T;
    static constexpr T none = ~0ull;
    static __device__ __forceinline__ T make(int d, int ix, int iy, int k) { return make_key(d, ix, iy, k); }
    static __device__ __forceinline__ unsigned pack(T m) { return ((unsigned)(m >> 32) << 20) | (unsigned)(m & 0xfffff); }
};

// ascending bitonic sort of one key per lane across the warp (15 compare-exchange steps; the keys are unique, `none` sorts last)
template <typename Key>
__device__ __forceinline__ Key warp_sort_ascending(Key v, int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const Key o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
            v = take_min ? (o < v ? o : v) : (o < v ? v : o);
        }
    }
    return v;
}

// One warp per query.  The features of the window's grid cells are dealt to the lanes one by one (prefix sum of the cell populations + a 5-step
// search for the owning cell), so a typical window -- 4 to 9 cells holding about a dozen features -- keeps a dozen lanes busy for ONE pass instead of
// 4 to 9 lanes walking their cells serially.  Every batch of <= 32 candidate keys is sorted across the warp and merged into the running top-kTop list
// (lanes 0 .. kTop-1) with a 16-lane bitonic merge.  Same result as a sequential scan: the kTop smallest keys in the reference's comparison order.
template <bool K32>
__global__ void __launch_bounds__(256)
k_search_candidates(const SearchArgs A)
{
    typedef CandKey<K32> CK;
    typedef typename CK::T Key;
    constexpr unsigned full = 0xffffffffu;
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= A.q_counts[f]) return;
    const size_t fo = (size_t)f * A.f_slab, qo = (size_t)f * A.q_slab;
    unsigned *top = A.top + (qo + i) * kTop;
    if (!A.q_valid[qo + i]) { if (lane == 0) A.ncand[qo + i] = 0; return; }
    const float2 *f_xy = A.f_xy + fo; const int *f_octave = A.f_octave + fo; const uint4 *f_desc = A.f_desc + 2 * fo;
    const int *cs = A.cell_start + (size_t)f * (kGridCells + 1); const int *items = A.cell_items + fo;
    const float2 uv = A.q_uv[qo + i];
    const float r = A.q_radius[qo + i];
    const int minl = A.q_minl[qo + i], maxl = A.q_maxl[qo + i];
    const Window w = make_window(A.g, uv, r, minl, maxl);
    Key best = CK::none;      // lane t < kTop: the t-th smallest key so far
    bool first = true;
    int n = 0;
    if (!w.empty) {
        const uint4 qa = __ldg(&A.q_desc[2 * (qo + i)]), qb = __ldg(&A.q_desc[2 * (qo + i) + 1]);
        const int ncy = w.r1 - w.r0 + 1, ncells = (w.c1 - w.c0 + 1) * ncy;
        for (int cb = 0; cb < ncells; cb += 32) {
            const int c = cb + lane;
            int start = 0, cnt = 0, cxy = 0;
            if (c < ncells) {
                const int ix = w.c0 + c / ncy, iy = w.r0 + c % ncy;
                const int cell = ix * kGridRows + iy;
                start = cs[cell]; cnt = cs[cell + 1] - start; cxy = (ix << 8) | iy;
            }
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(full, incl, d); if (lane >= d) incl += o; }
            const int tot = __shfl_sync(full, incl, 31);
            const int first_item = start - (incl - cnt);                   // items index of the cell's first feature minus its rank in the batch order
            for (int t0 = 0; t0 < tot; t0 += 32) {
                const int t = t0 + lane;
                int lo = 0;                                                 // the cell that holds item t: number of lanes with incl <= t
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) { const int v = __shfl_sync(full, incl, lo + sft - 1); if (v <= t) lo += sft; }
                const int o_first = __shfl_sync(full, first_item, lo), o_cxy = __shfl_sync(full, cxy, lo);
                Key key = CK::none;
                if (t < tot) {
                    const int k = items[o_first + t];
                    if (in_window(w, minl, maxl, uv, r, f_octave[k], f_xy[k])) {
                        n++;
                        key = CK::make(hamming256(qa, qb, __ldg(&f_desc[2 * k]), __ldg(&f_desc[2 * k + 1])), o_cxy >> 8, o_cxy & 0xff, k);
                    }
                }
                key = warp_sort_ascending(key, lane);
                if (first) { best = lane < kTop ? key : CK::none; first = false; }
                else {
                    // lanes 0..7: the list so far (ascending), lanes 8..15: the 8 smallest new keys, descending -> one bitonic sequence of 16
                    const Key rev = __shfl_sync(full, key, (15 - lane) & 31);
                    Key v = lane < kTop ? best : (lane < 2 * kTop ? rev : CK::none);
#pragma unroll
                    for (int j = kTop; j > 0; j >>= 1) {
                        const Key o = __shfl_xor_sync(full, v, j);
                        v = ((lane & j) == 0) ? (o < v ? o : v) : (o < v ? v : o);
                    }
                    best = lane < kTop ? v : CK::none;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(full, n, d);
    if (lane < kTop) top[lane] = best == CK::none ? 0xffffffffu : CK::pack(best);
    if (lane == 0) A.ncand[qo + i] = n;
}


// Search half without claims (Fuse, SearchBySim3): one warp per query, lanes over the grid cells of the window; the minimum of
// (distance, visiting order) is the feature the reference's "dist < bestDist" loop ends with.
struct BestArgs {
    int f_slab, q_slab;
    GridParams g;
    const float2 *f_xy; const int *f_octave; const uint4 *f_desc; const int *f_counts;
    const uint8_t *q_valid; const float2 *q_uv; const float *q_radius; const int *q_minl; const int *q_maxl; const uint4 *q_desc; const int *q_counts;
    const int *cell_start; const int *cell_items;
    int th_dist, use_gate, nlevels; float chi2_gate; float inv_sigma2[32];
    int *best_idx; int *best_dist;
};

__global__ void __launch_bounds__(256)
k_search_best(const BestArgs A)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= A.q_counts[f]) return;
    const size_t fo = (size_t)f * A.f_slab, qo = (size_t)f * A.q_slab;
    unsigned long long best = kNoKey;
    if (A.q_valid[qo + i]) {
        const float2 *f_xy = A.f_xy + fo; const int *f_octave = A.f_octave + fo; const uint4 *f_desc = A.f_desc + 2 * fo;
        const int *cs = A.cell_start + (size_t)f * (kGridCells + 1); const int *items = A.cell_items + fo;
        const float2 uv = A.q_uv[qo + i];
        const float r = A.q_radius[qo + i];
        const int minl = A.q_minl[qo + i], maxl = A.q_maxl[qo + i];
        const Window w = make_window(A.g, uv, r, minl, maxl);
        if (!w.empty) {
            const uint4 qa = __ldg(&A.q_desc[2 * (qo + i)]), qb = __ldg(&A.q_desc[2 * (qo + i) + 1]);
            const int ncy = w.r1 - w.r0 + 1, ncells = (w.c1 - w.c0 + 1) * ncy;
            for (int c = lane; c < ncells; c += 32) {
                const int ix = w.c0 + c / ncy, iy = w.r0 + c % ncy;
                const int cell = ix * kGridRows + iy;
                const int je = cs[cell + 1];
                for (int j = cs[cell]; j < je; j++) {
                    const int k = items[j];
                    const int oct = f_octave[k];
                    const float2 p = f_xy[k];
                    if (!in_window(w, minl, maxl, uv, r, oct, p)) continue;
                    if (A.use_gate) {                              // ORBmatcher.cc:918-928
                        const float ex = __fsub_rn(uv.x, p.x), ey = __fsub_rn(uv.y, p.y);
                        const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                        if (__fmul_rn(e2, A.inv_sigma2[min(max(oct, 0), A.nlevels - 1)]) > A.chi2_gate) continue;
                    }
                    const unsigned long long key = make_key(hamming256(qa, qb, __ldg(&f_desc[2 * k]), __ldg(&f_desc[2 * k + 1])), ix, iy, k);
                    best = key < best ? key : best;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o < best ? o : best; }
    if (lane == 0) {
        const int dist = best == kNoKey ? -1 : (int)(best >> 32);
        A.best_dist[qo + i] = dist;
        A.best_idx[qo + i] = (dist >= 0 && dist <= A.th_dist) ? (int)(best & 0xfffff) : -1;
    }
}

// slow path: full rescan of one query's window by one WARP (lanes over grid cells, like k_search_candidates); best and
// second-best free candidates in the reference's comparison order, valid in every lane
__device__ void rescan_window_warp(const SearchArgs &A, int f, int i, const int *own_prev, int lane, unsigned long long &b1, unsigned long long &b2)
{
    const size_t fo = (size_t)f * A.f_slab, qo = (size_t)f * A.q_slab;
    const float2 *f_xy = A.f_xy + fo; const int *f_octave = A.f_octave + fo; const uint4 *f_desc = A.f_desc + 2 * fo;
    const int *cs = A.cell_start + (size_t)f * (kGridCells + 1); const int *items = A.cell_items + fo;
    const float2 uv = A.q_uv[qo + i];
    const float r = A.q_radius[qo + i];
    const int minl = A.q_minl[qo + i], maxl = A.q_maxl[qo + i];
    const Window w = make_window(A.g, uv, r, minl, maxl);
    b1 = kNoKey; b2 = kNoKey;
    if (w.empty) return;
    const uint4 qa = __ldg(&A.q_desc[2 * (qo + i)]), qb = __ldg(&A.q_desc[2 * (qo + i) + 1]);
    const int ncy = w.r1 - w.r0 + 1, ncells = (w.c1 - w.c0 + 1) * ncy;
    for (int c = lane; c < ncells; c += 32) {
        const int ix = w.c0 + c / ncy, iy = w.r0 + c % ncy;
        const int cell = ix * kGridRows + iy;
        const int je = cs[cell + 1];
        for (int j = cs[cell]; j < je; j++) {
            const int k = items[j];
            if (!in_window(w, minl, maxl, uv, r, f_octave[k], f_xy[k])) continue;
            if (own_prev[k] < i) continue;
            const unsigned long long key = make_key(hamming256(qa, qb, __ldg(&f_desc[2 * k]), __ldg(&f_desc[2 * k + 1])), ix, iy, k);
            if (key < b1) { b2 = b1; b1 = key; } else if (key < b2) b2 = key;
        }
    }
    // warp minimum, then the minimum of what is left (keys are unique: they contain the feature index)
    unsigned long long m1 = b1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m1, d); m1 = o < m1 ? o : m1; }
    unsigned long long m2 = (b1 == m1) ? b2 : b1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m2, d); m2 = o < m2 ? o : m2; }
    b1 = m1; b2 = m2;
}

// use_smem: the working set of the fixed-point rounds (candidate lists, ownership, proposals) is staged in shared memory
// (88 KB at 2000 features); the rounds then cost barriers + smem traffic instead of L2 round trips.
__global__ void __launch_bounds__(1024)
k_search_resolve(const SearchArgs A, const int use_smem)
{
    extern __shared__ __align__(16) unsigned char resolve_smem[];
    __shared__ int s_changed, s_nqueue;
    __shared__ int s_hist[kHisto];
    __shared__ int s_keep[3];
    __shared__ int s_removed, s_accepted;
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int N = A.f_counts[f], M = A.q_counts[f];
    const size_t fo = (size_t)f * A.f_slab, qo = (size_t)f * A.q_slab;
    const int *f_octave = A.f_octave + fo;
    int *fm = A.feat_match + fo;
    int *prop = A.prop + qo;
    int *owner[2] = {A.owner + 2 * fo, A.owner + 2 * fo + A.f_slab};
    const unsigned *top_all = A.top + qo * kTop;
    const int *ncand = A.ncand + qo;
    int *s_queue = reinterpret_cast<int *>(resolve_smem);                 // [q_slab] queries waiting for a window rescan
    const bool need2 = A.ratio > 0.f;
    if (use_smem) {
        unsigned *s_top = reinterpret_cast<unsigned *>(resolve_smem) + A.q_slab;
        int *s_nc = reinterpret_cast<int *>(s_top + (size_t)A.q_slab * kTop);
        int *s_prop = s_nc + A.q_slab;
        int *s_own = s_prop + A.q_slab;
        for (int t = tid; t < M * kTop; t += nt) s_top[t] = top_all[t];
        for (int t = tid; t < M; t += nt) s_nc[t] = ncand[t];
        top_all = s_top; ncand = s_nc; prop = s_prop; owner[0] = s_own; owner[1] = s_own + A.f_slab;
    }

    // features that already hold a map point are owned by "query -1": every query skips them
    for (int k = tid; k < N; k += nt) { const int o = fm[k] >= 0 ? -1 : 0x7fffffff; owner[0][k] = o; owner[1][k] = o; }
    for (int i = tid; i < M; i += nt) prop[i] = -2;
    __syncthreads();

    int cur = 0;
    for (int round = 0; round <= M; round++) {
        if (tid == 0) { s_changed = 0; s_nqueue = 0; }
        __syncthreads();
        const int *own_prev = owner[cur];
        int *own_next = owner[cur ^ 1];
        for (int i = tid; i < M; i += nt) {
            int choice = -1;
            const int nc = ncand[i];
            if (nc > 0) {
                const unsigned *top = top_all + (size_t)i * kTop;
                // first and second free entries of the sorted list
                unsigned e1 = 0xffffffffu, e2 = 0xffffffffu;
#pragma unroll
                for (int t = 0; t < kTop; t++) {
                    const unsigned e = top[t];
                    if (e == 0xffffffffu) continue;
                    if (own_prev[e & 0xfffff] < i) continue;
                    if (e1 == 0xffffffffu) e1 = e; else if (e2 == 0xffffffffu) e2 = e;
                }
                int best = -1, bidx = -1, best2 = 256, lvl2 = -1;
                const bool list_complete = nc <= kTop;
                if ((e1 == 0xffffffffu || (need2 && e2 == 0xffffffffu)) && !list_complete) {
                    s_queue[atomicAdd(&s_nqueue, 1)] = i;          // rare: resolved by a whole warp below
                    continue;
                }
                if (e1 != 0xffffffffu) { best = (int)(e1 >> 20); bidx = (int)(e1 & 0xfffff); }
                if (e2 != 0xffffffffu) { best2 = (int)(e2 >> 20); lvl2 = f_octave[e2 & 0xfffff]; }
                if (bidx >= 0 && best <= A.th_dist) {
                    bool acc = true;
                    if (need2 && f_octave[bidx] == lvl2 && (float)best > __fmul_rn(A.ratio, (float)best2)) acc = false;
                    if (acc) choice = bidx;
                }
            }
            if (prop[i] != choice) { prop[i] = choice; s_changed = 1; }
        }
        __syncthreads();
        // queries whose list is exhausted although their window holds more than kTop candidates: one warp each
        for (int q = tid >> 5; q < s_nqueue; q += nt >> 5) {
            const int i = s_queue[q];
            unsigned long long b1, b2;
            rescan_window_warp(A, f, i, own_prev, tid & 31, b1, b2);
            if ((tid & 31) == 0) {
                int choice = -1, best = -1, bidx = -1, best2 = 256, lvl2 = -1;
                if (b1 != kNoKey) { best = (int)(b1 >> 32); bidx = (int)(b1 & 0xfffff); }
                if (b2 != kNoKey) { best2 = (int)(b2 >> 32); lvl2 = f_octave[(int)(b2 & 0xfffff)]; }
                if (bidx >= 0 && best <= A.th_dist) {
                    bool acc = true;
                    if (need2 && f_octave[bidx] == lvl2 && (float)best > __fmul_rn(A.ratio, (float)best2)) acc = false;
                    if (acc) choice = bidx;
                }
                if (prop[i] != choice) { prop[i] = choice; s_changed = 1; }
            }
        }
        __syncthreads();
        if (!s_changed) break;
        for (int k = tid; k < N; k += nt) own_next[k] = fm[k] >= 0 ? -1 : 0x7fffffff;
        __syncthreads();
        for (int i = tid; i < M; i += nt) { const int k = prop[i]; if (k >= 0) atomicMin(&own_next[k], i); }
        __syncthreads();
        cur ^= 1;
    }

    // commit + rotation consistency (ORBmatcher.cc:1421-1468)
    if (tid < kHisto) s_hist[tid] = 0;
    if (tid == 0) { s_removed = 0; s_accepted = 0; }
    __syncthreads();
    const float factor = 1.0f / kHisto;
    int my_acc = 0;
    for (int i = tid; i < M; i += nt) {
        const int k = prop[i];
        if (k < 0) continue;
        fm[k] = i;
        my_acc++;
        if (A.check_ori) {
            float rot = __fsub_rn(A.q_angle[qo + i], A.f_angle[fo + k]);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == kHisto) bin = 0;
            bin = min(max(bin, 0), kHisto - 1);
            atomicAdd(&s_hist[bin], 1);
            prop[i] = k | (bin << 24);
        }
    }
    if (my_acc) atomicAdd(&s_accepted, my_acc);
    __syncthreads();
    if (A.check_ori) {
        if (tid == 0) {                       // ComputeThreeMaxima, ORBmatcher.cc:1603-1644
            int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < kHisto; i++) {
                const int s = s_hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
                else if (s > max3) { max3 = s; i3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
            s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3;
        }
        __syncthreads();
        int my_rm = 0;
        for (int i = tid; i < M; i += nt) {
            const int v = prop[i];
            if (v < 0) continue;
            const int bin = v >> 24, k = v & 0xffffff;
            if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { fm[k] = -1; my_rm++; }
        }
        if (my_rm) atomicAdd(&s_removed, my_rm);
        __syncthreads();
    }
    if (tid == 0) A.nmatches[f] = s_accepted - s_removed;
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// Vocabulary-bucket matchers: SearchByBoW(KeyFrame*, Frame&) :159-290, SearchByBoW(KeyFrame*, KeyFrame*) :524-657 (mode 0) and
// SearchForTriangulation :659-825 (mode 1).  Features are only compared inside a vocabulary node, and a claim never leaves its node, so
// the nodes are independent: ONE WARP PER COMMON NODE walks the side-1 features of the node in the reference's order (the claim chain),
// the 32 lanes scan the node's side-2 features (XOR + POPC) and a warp reduction yields best / second best in the reference's comparison
// order.  The merge join over the two std::map<NodeId, ...> is a binary search of the side-1 node id in side 2's sorted node list.
struct BowArgs {
    int mode, slab1, slab2, nslab1, nslab2;
    const uint4 *desc1, *desc2; const float *angle1, *angle2; const uint8_t *elig1, *elig2;
    const int *counts1;
    const int *nodes1, *start1, *items1, *ncount1;      // [n_pairs, nslab1], [n_pairs, nslab1 + 1], [n_pairs, slab1], [n_pairs]
    const int *nodes2, *start2, *items2, *ncount2;
    float ratio; int check_ori;
    const float2 *xy1, *xy2; const int *octave2; const float *F12, *epipole; float scale_factors2[32], level_sigma2_2[32]; int nlevels;   // mode 1
    uint8_t *matched2;                                  // scratch [n_pairs, slab2], zeroed
    int *match12; int *rot_bin;                         // [n_pairs, slab1]
    int *nmatches;                                      // [n_pairs]
};

__device__ __forceinline__ int rot_hist_bin(float a1, float a2)
{
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / kHisto));
    if (bin == kHisto) bin = 0;
    return min(max(bin, 0), kHisto - 1);
}

__global__ void __launch_bounds__(256)
k_bow_match(const BowArgs A)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= A.ncount1[f]) return;
    const int *nodes2 = A.nodes2 + (size_t)f * A.nslab2;
    const int node = A.nodes1[(size_t)f * A.nslab1 + a];
    int lo = 0, hi = A.ncount2[f];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (nodes2[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= A.ncount2[f] || nodes2[lo] != node) return;
    const size_t o1 = (size_t)f * A.slab1, o2 = (size_t)f * A.slab2;
    const int *st1 = A.start1 + (size_t)f * (A.nslab1 + 1), *st2 = A.start2 + (size_t)f * (A.nslab2 + 1);
    const int *items1 = A.items1 + o1, *items2 = A.items2 + o2;
    const int u0 = st1[a], u1 = st1[a + 1], v0 = st2[lo], v1 = st2[lo + 1];
    uint8_t *matched2 = A.matched2 + o2;
    float F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ex = 0, ey = 0;
    if (A.mode == 1) { for (int k = 0; k < 9; k++) F[k] = A.F12[9 * f + k]; ex = A.epipole[2 * f]; ey = A.epipole[2 * f + 1]; }
    for (int u = u0; u < u1; u++) {
        const int idx1 = items1[u];
        if (!A.elig1[o1 + idx1]) continue;
        const uint4 qa = __ldg(&A.desc1[2 * (o1 + idx1)]), qb = __ldg(&A.desc1[2 * (o1 + idx1) + 1]);
        // key = distance << 20 | position in the node (mode 0: first minimum wins; mode 1: last minimum wins -> position inverted)
        unsigned b1 = 0xffffffffu, b2 = 0xffffffffu;
        float la = 0, lb = 0, lc = 0;
        if (A.mode == 1) {                                  // epipolar line of kp1 in image 2 (CheckDistEpipolarLine, :140-146)
            const float2 p1 = A.xy1[o1 + idx1];
            la = __fadd_rn(__fadd_rn(__fmul_rn(p1.x, F[0]), __fmul_rn(p1.y, F[3])), F[6]);
            lb = __fadd_rn(__fadd_rn(__fmul_rn(p1.x, F[1]), __fmul_rn(p1.y, F[4])), F[7]);
            lc = __fadd_rn(__fadd_rn(__fmul_rn(p1.x, F[2]), __fmul_rn(p1.y, F[5])), F[8]);
        }
        for (int v = v0 + lane; v < v1; v += 32) {
            const int idx2 = items2[v];
            if (matched2[idx2] || !A.elig2[o2 + idx2]) continue;
            const int d = hamming256(qa, qb, __ldg(&A.desc2[2 * (o2 + idx2)]), __ldg(&A.desc2[2 * (o2 + idx2) + 1]));
            unsigned key;
            if (A.mode == 0) key = ((unsigned)d << 20) | (unsigned)(v - v0);
            else {
                if (d > ORBM_TH_LOW) continue;
                const float2 p2 = A.xy2[o2 + idx2];
                const int oct = min(max(A.octave2[o2 + idx2], 0), A.nlevels - 1);
                const float dx = __fsub_rn(ex, p2.x), dy = __fsub_rn(ey, p2.y);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, A.scale_factors2[oct])) continue;      // :747-753
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, p2.x), __fmul_rn(lb, p2.y)), lc);
                const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
                if (den == 0) continue;
                const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                if (!((double)dsqr < 3.84 * (double)A.level_sigma2_2[oct])) continue;
                key = ((unsigned)d << 20) | (unsigned)(0xfffff - (v - v0));
            }
            if (key < b1) { b2 = b1; b1 = key; } else if (key < b2) b2 = key;
        }
        unsigned m1 = b1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m1 = min(m1, __shfl_xor_sync(0xffffffffu, m1, d));
        unsigned m2 = (b1 == m1) ? b2 : b1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, d));
        if (m1 == 0xffffffffu) continue;
        const int best1 = (int)(m1 >> 20), best2 = m2 == 0xffffffffu ? 256 : (int)(m2 >> 20);
        bool accept;
        int idx2;
        if (A.mode == 0) { accept = best1 <= ORBM_TH_LOW && (float)best1 < __fmul_rn(A.ratio, (float)best2); idx2 = items2[v0 + (int)(m1 & 0xfffff)]; }
        else { accept = true; idx2 = items2[v0 + (0xfffff - (int)(m1 & 0xfffff))]; }
        if (!accept) continue;
        if (lane == 0) {
            A.match12[o1 + idx1] = idx2;
            if (A.mode == 0) matched2[idx2] = 1;
            if (A.check_ori) A.rot_bin[o1 + idx1] = rot_hist_bin(A.angle1[o1 + idx1], A.angle2[o2 + idx2]);
        }
        __syncwarp();
    }
}

__device__ inline void three_maxima_dev(const int *hist, int *keep)        // ComputeThreeMaxima, ORBmatcher.cc:1603-1644
{
    int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
    for (int i = 0; i < kHisto; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
        else if (s > max3) { max3 = s; i3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
    keep[0] = i1; keep[1] = i2; keep[2] = i3;
}

// rotation-consistency filter + match count of one pair (one CTA per pair)
__global__ void __launch_bounds__(256)
k_bow_finish(int slab1, const int *__restrict__ counts1, int check_ori, int *__restrict__ match12, const int *__restrict__ rot_bin, int *__restrict__ nmatches)
{
    __shared__ int s_hist[kHisto], s_keep[3], s_n;
    const int f = blockIdx.x, tid = threadIdx.x, N = counts1[f];
    int *m = match12 + (size_t)f * slab1;
    const int *rb = rot_bin + (size_t)f * slab1;
    if (tid < kHisto) s_hist[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();
    if (check_ori) {
        for (int i = tid; i < N; i += blockDim.x) if (m[i] >= 0) atomicAdd(&s_hist[rb[i]], 1);
        __syncthreads();
        if (tid == 0) three_maxima_dev(s_hist, s_keep);
        __syncthreads();
    }
    int cnt = 0;
    for (int i = tid; i < N; i += blockDim.x) {
        if (m[i] < 0) continue;
        if (check_ori) { const int b = rb[i]; if (b != s_keep[0] && b != s_keep[1] && b != s_keep[2]) { m[i] = -1; continue; } }
        cnt++;
    }
    if (cnt) atomicAdd(&s_n, cnt);
    __syncthreads();
    if (tid == 0) nmatches[f] = s_n;
}

// ORBmatcher::SearchForInitialization (:407-522): one CTA per frame pair walks the level-0 keypoints of F1 in order (a later keypoint may
// steal a feature of F2 from an earlier one, so the loop is a chain); the threads scan the grid cells of the window.
struct InitArgs {
    int slab1, slab2;
    GridParams g;
    const int *octave1; const float *angle1; const uint4 *desc1; const int *counts1;
    const float2 *xy2; const int *octave2; const float *angle2; const uint4 *desc2; const int *counts2;
    const int *cell_start, *cell_items;
    float2 *prev_matched;                 // [n_pairs, slab1] in/out
    float window, ratio; int check_ori;
    int *matched_dist, *matches21;        // scratch [n_pairs, slab2]
    int *matches12, *rot_bin;             // [n_pairs, slab1]
    int *nmatches;
};

constexpr int kInitThreads = 256;

__global__ void __launch_bounds__(kInitThreads)
k_search_init(const InitArgs A)
{
    __shared__ int s_hist[kHisto], s_keep[3];
    __shared__ unsigned long long s_b1[kInitThreads / 32], s_b2[kInitThreads / 32];
    __shared__ int s_cnt;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nt = blockDim.x;
    const int N1 = A.counts1[f], N2 = A.counts2[f];
    const size_t o1 = (size_t)f * A.slab1, o2 = (size_t)f * A.slab2;
    int *matched_dist = A.matched_dist + o2, *matches21 = A.matches21 + o2, *matches12 = A.matches12 + o1, *rot_bin = A.rot_bin + o1;
    const float2 *xy2 = A.xy2 + o2; const int *octave2 = A.octave2 + o2; const uint4 *desc2 = A.desc2 + 2 * o2;
    const int *cs = A.cell_start + (size_t)f * (kGridCells + 1); const int *items = A.cell_items + o2;
    for (int i = tid; i < N2; i += nt) { matched_dist[i] = 0x7fffffff; matches21[i] = -1; }
    for (int i = tid; i < N1; i += nt) { matches12[i] = -1; rot_bin[i] = -1; }
    if (tid < kHisto) s_hist[tid] = 0;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int i1 = 0; i1 < N1; i1++) {
        if (A.octave1[o1 + i1] > 0) continue;
        const float2 uv = A.prev_matched[o1 + i1];
        const Window w = make_window(A.g, uv, A.window, 0, 0);
        if (w.empty) continue;
        const uint4 qa = __ldg(&A.desc1[2 * (o1 + i1)]), qb = __ldg(&A.desc1[2 * (o1 + i1) + 1]);
        unsigned long long b1 = kNoKey, b2 = kNoKey;
        const int ncy = w.r1 - w.r0 + 1, ncells = (w.c1 - w.c0 + 1) * ncy;
        for (int c = tid; c < ncells; c += nt) {
            const int ix = w.c0 + c / ncy, iy = w.r0 + c % ncy;
            const int cell = ix * kGridRows + iy;
            const int je = cs[cell + 1];
            for (int j = cs[cell]; j < je; j++) {
                const int k = items[j];
                if (!in_window(w, 0, 0, uv, A.window, octave2[k], xy2[k])) continue;
                const int d = hamming256(qa, qb, __ldg(&desc2[2 * k]), __ldg(&desc2[2 * k + 1]));
                if (matched_dist[k] <= d) continue;
                const unsigned long long key = make_key(d, ix, iy, k);
                if (key < b1) { b2 = b1; b1 = key; } else if (key < b2) b2 = key;
            }
        }
        // block-wide smallest and second smallest key (keys are unique: they contain the feature index)
        unsigned long long m1 = b1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m1, d); m1 = o < m1 ? o : m1; }
        unsigned long long m2 = (b1 == m1) ? b2 : b1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m2, d); m2 = o < m2 ? o : m2; }
        if (lane == 0) { s_b1[wid] = m1; s_b2[wid] = m2; }
        __syncthreads();
        m1 = kNoKey; m2 = kNoKey;
#pragma unroll
        for (int q = 0; q < kInitThreads / 32; q++) {
            const unsigned long long x1 = s_b1[q], x2 = s_b2[q];
            if (x1 < m1) { m2 = m1 < x2 ? m1 : x2; m1 = x1; } else { m2 = x1 < m2 ? x1 : m2; }
        }
        if (m1 != kNoKey) {
            const int best = (int)(m1 >> 32), bidx = (int)(m1 & 0xfffff);
            const float best2f = m2 == kNoKey ? 2147483648.0f : (float)(int)(m2 >> 32);          // (float)INT_MAX
            if (best <= ORBM_TH_LOW && (float)best < __fmul_rn(best2f, A.ratio) && tid == 0) {
                const int prev = matches21[bidx];
                if (prev >= 0) matches12[prev] = -1;
                matches12[i1] = bidx; matches21[bidx] = i1; matched_dist[bidx] = best;
                if (A.check_ori) { const int bin = rot_hist_bin(A.angle1[o1 + i1], A.angle2[o2 + bidx]); rot_bin[i1] = bin; s_hist[bin]++; }
            }
        }
        __syncthreads();          // the update is visible to every thread of the next step; s_b1 / s_b2 may be rewritten
    }
    if (A.check_ori) {
        if (tid == 0) three_maxima_dev(s_hist, s_keep);
        __syncthreads();
        for (int i = tid; i < N1; i += nt) { const int b = rot_bin[i]; if (b >= 0 && b != s_keep[0] && b != s_keep[1] && b != s_keep[2]) matches12[i] = -1; }
        __syncthreads();
    }
    int nmatches = 0;
    for (int i = tid; i < N1; i += nt) if (matches12[i] >= 0) { nmatches++; A.prev_matched[o1 + i] = xy2[matches12[i]]; }
    if (nmatches) atomicAdd(&s_cnt, nmatches);
    __syncthreads();
    if (tid == 0) A.nmatches[f] = s_cnt;
}

}  // namespace orbs

using namespace orbs;

struct orbm_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    long long launches = 0;
    std::mutex mu;
    DevBuf cell_start, cell_items, prop, owner, top, ncand;
    StagePool pool;
};

namespace {
GridParams make_grid(const float *b4)
{
    GridParams g;
    g.min_x = b4[0]; g.min_y = b4[1]; g.max_x = b4[2]; g.max_y = b4[3];
    g.w_inv = (float)kGridCols / (g.max_x - g.min_x);          // Frame.cc:101-102
    g.h_inv = (float)kGridRows / (g.max_y - g.min_y);
    g.win_x = g.min_x; g.win_y = g.min_y;
    return g;
}
}  // namespace

extern "C" {

int orbm_create(orbm_handle **out, int device)
{
    ORBS_REQUIRE(out, ORBS_E_INVALID, "orbm_create: null out pointer");
    *out = nullptr;
    ORBS_CUDA(cudaSetDevice(device));
    orbm_handle *h = new orbm_handle();
    h->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    // per-device function attribute, shared by all handles (the drop-in creates one per calling thread): always the static bound
    e = cudaFuncSetAttribute(k_search_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { cudaStreamDestroy(h->stream); delete h; return cuda_fail(e, "cudaFuncSetAttribute(k_search_resolve)", __FILE__, __LINE__); }
    *out = h;
    return ORBS_OK;
}

int orbm_destroy(orbm_handle *h)
{
    if (!h) return ORBS_OK;
    cudaSetDevice(h->device);
    if (h->stream && h->own_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    else cudaDeviceSynchronize();
    h->cell_start.release(); h->cell_items.release(); h->prop.release(); h->owner.release(); h->top.release(); h->ncand.release();
    h->pool.release();
    delete h;
    return ORBS_OK;
}

int orbm_set_stream(orbm_handle *h, void *stream)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)stream; h->own_stream = false;
    return ORBS_OK;
}

void *orbm_stream(orbm_handle *h) { return h ? (void *)h->stream : nullptr; }
int orbm_synchronize(orbm_handle *h)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    return ORBS_OK;
}
long long orbm_kernel_launches(const orbm_handle *h) { return h ? h->launches : 0; }

int orbm_descriptor_distance(orbm_handle *h, const uint8_t *a, int n, const uint8_t *b, int m, int32_t *out, int memspace)
{
    ORBS_REQUIRE(h && a && b && out, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n >= 0 && m >= 0, ORBS_E_INVALID, "negative size");
    if (n == 0 || m == 0) return ORBS_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const uint8_t *da = S.in(a, (size_t)n * 32), *db = S.in(b, (size_t)m * 32);
    int32_t *dout = S.inout(out, (size_t)n * m, false);
    if (S.rc) return S.rc;
    ORBS_REQUIRE(((uintptr_t)da % 16 == 0) && ((uintptr_t)db % 16 == 0), ORBS_E_INVALID, "descriptor arrays must be 16-byte aligned");
    const size_t total = (size_t)n * m;
    k_hamming_pairs<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>((const uint4 *)da, n, (const uint4 *)db, m, dout);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_distinctive_descriptors(orbm_handle *h, int n_points, const uint8_t *desc, const int32_t *start, int32_t *best_idx, int memspace)
{
    ORBS_REQUIRE(h && desc && start && best_idx, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_points >= 0, ORBS_E_INVALID, "negative size");
    if (n_points == 0) return ORBS_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    int32_t total = 0;
    if (memspace == ORBS_MEM_HOST) total = start[n_points];
    else ORBS_CUDA(cudaMemcpy(&total, start + n_points, sizeof(int32_t), cudaMemcpyDeviceToHost));
    ORBS_REQUIRE(total >= 0, ORBS_E_INVALID, "start[] must be non-decreasing from 0");
    if (memspace == ORBS_MEM_HOST)
        for (int p = 0; p < n_points; p++)
            ORBS_REQUIRE(start[p + 1] >= start[p] && start[p + 1] - start[p] <= 65535, ORBS_E_INVALID, "a map point has a negative number or more than 65535 observations");
    Stager S(&h->pool, h->stream, memspace);
    const uint8_t *dd = S.in(desc, (size_t)std::max(total, 1) * 32);
    const int32_t *ds = S.in(start, (size_t)n_points + 1);
    int32_t *dout = S.inout(best_idx, (size_t)n_points, false);
    if (S.rc) return S.rc;
    ORBS_REQUIRE((uintptr_t)dd % 16 == 0, ORBS_E_INVALID, "descriptor array must be 16-byte aligned");
    k_distinctive<<<(n_points + 7) / 8, 256, 0, h->stream>>>(n_points, (const uint4 *)dd, ds, dout);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_project_last_frame(orbm_handle *h, int n_frames, const float *Tcw, const float *K4, const float *bounds4,
                            const float *scale_factors, int nlevels, const float *Xw, const int32_t *last_octave,
                            const int32_t *q_counts, int q_slab, float th, uint8_t *q_valid, float *q_uv,
                            float *q_radius, int32_t *q_minl, int32_t *q_maxl, int memspace)
{
    ORBS_REQUIRE(h && Tcw && K4 && bounds4 && scale_factors && Xw && last_octave && q_counts && q_valid && q_uv && q_radius && q_minl && q_maxl,
                 ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && q_slab > 0 && nlevels > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t nq = (size_t)n_frames * q_slab;
    // K4 / bounds4 are tiny host-side parameter blocks in both memory spaces
    const float *dT = S.in(Tcw, (size_t)n_frames * 16), *dsf = S.in(scale_factors, nlevels), *dX = S.in(Xw, nq * 3);
    const int32_t *doct = S.in(last_octave, nq), *dqc = S.in(q_counts, n_frames);
    uint8_t *dval = S.inout(q_valid, nq);
    float *duv = S.inout(q_uv, nq * 2, false), *drad = S.inout(q_radius, nq, false);
    int32_t *dmn = S.inout(q_minl, nq, false), *dmx = S.inout(q_maxl, nq, false);
    if (S.rc) return S.rc;
    k_project_last<<<dim3((q_slab + 255) / 256, n_frames), 256, 0, h->stream>>>(q_slab, dT, K4[0], K4[1], K4[2], K4[3], make_grid(bounds4), dsf,
                                                                               nlevels, dX, doct, dqc, th, dval, (float2 *)duv, drad, dmn, dmx);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_undistort_keypoints(orbm_handle *h, int n_frames, const float *kp_xy, const int32_t *counts, int slab, const float *K4,
                             const float *dist5, float *kp_xy_un, int memspace)
{
    ORBS_REQUIRE(h && kp_xy && counts && K4 && dist5 && kp_xy_un, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && slab > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t n = (size_t)n_frames * slab;
    const float *dxy = S.in(kp_xy, n * 2);
    const int32_t *dc = S.in(counts, n_frames);
    float *dun = S.inout(kp_xy_un, n * 2, false);
    if (S.rc) return S.rc;
    k_undistort<<<dim3((slab + 255) / 256, n_frames), 256, 0, h->stream>>>(slab, dc, (const float2 *)dxy, (float2 *)dun, (double)K4[0], (double)K4[1],
                                                                           (double)K4[2], (double)K4[3], (double)dist5[0], (double)dist5[1],
                                                                           (double)dist5[2], (double)dist5[3], (double)dist5[4]);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_is_in_frustum(orbm_handle *h, int n_frames, const float *Tcw, const float *Ow, const float *K4, const float *bounds4,
                       float log_scale_factor, float viewing_cos_limit, const float *Xw, const float *normal,
                       const float *mf_min_distance, const float *mf_max_distance, const int32_t *counts, int slab,
                       uint8_t *in_view, float *proj_xy, int32_t *pred_level, float *view_cos, int memspace)
{
    ORBS_REQUIRE(h && Tcw && Ow && K4 && bounds4 && Xw && normal && mf_min_distance && mf_max_distance && counts && in_view && proj_xy && pred_level &&
                 view_cos, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && slab > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t n = (size_t)n_frames * slab;
    const float *dT = S.in(Tcw, (size_t)n_frames * 16), *dO = S.in(Ow, (size_t)n_frames * 3), *dX = S.in(Xw, n * 3), *dN = S.in(normal, n * 3);
    const float *dmn = S.in(mf_min_distance, n), *dmx = S.in(mf_max_distance, n);
    const int32_t *dc = S.in(counts, n_frames);
    uint8_t *div = S.inout(in_view, n, false);
    float *duv = S.inout(proj_xy, n * 2, false), *dvc = S.inout(view_cos, n, false);
    int32_t *dlv = S.inout(pred_level, n, false);
    if (S.rc) return S.rc;
    if (memspace == ORBS_MEM_HOST) {       // slots beyond counts[f] and out-of-view points keep defined values on the host
        ORBS_CUDA(cudaMemsetAsync(div, 0, n, h->stream)); ORBS_CUDA(cudaMemsetAsync(duv, 0, n * 8, h->stream));
        ORBS_CUDA(cudaMemsetAsync(dvc, 0, n * 4, h->stream)); ORBS_CUDA(cudaMemsetAsync(dlv, 0, n * 4, h->stream));
    }
    k_in_frustum<<<dim3((slab + 255) / 256, n_frames), 256, 0, h->stream>>>(slab, dc, dT, dO, K4[0], K4[1], K4[2], K4[3], make_grid(bounds4), log_scale_factor,
                                                                            viewing_cos_limit, dX, dN, dmn, dmx, div, (float2 *)duv, dlv, dvc);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_search_by_projection(orbm_handle *h, int n_frames, const float *bounds4,
                              const float *f_xy, const int32_t *f_octave, const float *f_angle, const uint8_t *f_desc,
                              const int32_t *f_counts, int f_slab,
                              const uint8_t *q_valid, const float *q_uv, const float *q_radius, const int32_t *q_minl,
                              const int32_t *q_maxl, const float *q_angle, const uint8_t *q_desc, const int32_t *q_counts,
                              int q_slab, int th_dist, float ratio, int check_ori,
                              int32_t *feat_match, int32_t *nmatches, int memspace)
{
    return orbm_search_by_projection_kf(h, n_frames, bounds4, nullptr, f_xy, f_octave, f_angle, f_desc, f_counts, f_slab, q_valid, q_uv, q_radius,
                                        q_minl, q_maxl, q_angle, q_desc, q_counts, q_slab, th_dist, ratio, check_ori, feat_match, nmatches, memspace);
}

int orbm_search_by_projection_kf(orbm_handle *h, int n_frames, const float *bounds4, const float *win_origin2,
                                 const float *f_xy, const int32_t *f_octave, const float *f_angle, const uint8_t *f_desc,
                                 const int32_t *f_counts, int f_slab,
                                 const uint8_t *q_valid, const float *q_uv, const float *q_radius, const int32_t *q_minl,
                                 const int32_t *q_maxl, const float *q_angle, const uint8_t *q_desc, const int32_t *q_counts,
                                 int q_slab, int th_dist, float ratio, int check_ori,
                                 int32_t *feat_match, int32_t *nmatches, int memspace)
{
    ORBS_REQUIRE(h && bounds4 && f_xy && f_octave && f_desc && f_counts && q_valid && q_uv && q_radius && q_minl && q_maxl && q_desc && q_counts &&
                 feat_match && nmatches, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(!check_ori || (f_angle && q_angle), ORBS_E_INVALID, "orientation check needs the angles");
    ORBS_REQUIRE(n_frames > 0 && f_slab > 0 && q_slab > 0, ORBS_E_INVALID, "non-positive size");
    ORBS_REQUIRE(f_slab < (1 << 20) && q_slab < (1 << 24), ORBS_E_INVALID, "slab too large (features < 2^20, queries < 2^24 per frame)");
    ORBS_REQUIRE(bounds4[2] > bounds4[0] && bounds4[3] > bounds4[1], ORBS_E_INVALID, "empty image bounds");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t nf = (size_t)n_frames * f_slab, nq = (size_t)n_frames * q_slab;
    SearchArgs A;
    A.f_slab = f_slab; A.q_slab = q_slab; A.g = make_grid(bounds4);
    if (win_origin2) { A.g.win_x = win_origin2[0]; A.g.win_y = win_origin2[1]; }
    A.f_xy = (const float2 *)S.in(f_xy, nf * 2); A.f_octave = S.in(f_octave, nf);
    A.f_angle = f_angle ? S.in(f_angle, nf) : nullptr; A.f_desc = (const uint4 *)S.in(f_desc, nf * 32);
    A.f_counts = S.in(f_counts, n_frames);
    A.q_valid = S.in(q_valid, nq); A.q_uv = (const float2 *)S.in(q_uv, nq * 2); A.q_radius = S.in(q_radius, nq);
    A.q_minl = S.in(q_minl, nq); A.q_maxl = S.in(q_maxl, nq); A.q_angle = q_angle ? S.in(q_angle, nq) : nullptr;
    A.q_desc = (const uint4 *)S.in(q_desc, nq * 32); A.q_counts = S.in(q_counts, n_frames);
    A.feat_match = S.inout(feat_match, nf); A.nmatches = S.inout(nmatches, n_frames, false);
    if (S.rc) return S.rc;
    ORBS_REQUIRE(((uintptr_t)A.f_desc % 16 == 0) && ((uintptr_t)A.q_desc % 16 == 0) && ((uintptr_t)A.f_xy % 8 == 0) && ((uintptr_t)A.q_uv % 8 == 0),
                 ORBS_E_INVALID, "descriptor arrays must be 16-byte aligned, coordinate arrays 8-byte aligned");
    int rc;
    if ((rc = h->cell_start.reserve((size_t)n_frames * (kGridCells + 1) * sizeof(int)))) return rc;
    if ((rc = h->cell_items.reserve(nf * sizeof(int)))) return rc;
    if ((rc = h->prop.reserve(nq * sizeof(int)))) return rc;
    if ((rc = h->owner.reserve(2 * nf * sizeof(int)))) return rc;
    if ((rc = h->top.reserve(nq * kTop * sizeof(unsigned)))) return rc;
    if ((rc = h->ncand.reserve(nq * sizeof(int)))) return rc;
    A.cell_start = h->cell_start.as<int>(); A.cell_items = h->cell_items.as<int>();
    A.prop = h->prop.as<int>(); A.owner = h->owner.as<int>(); A.top = h->top.as<unsigned>(); A.ncand = h->ncand.as<int>();
    A.th_dist = th_dist; A.ratio = ratio; A.check_ori = check_ori ? 1 : 0;
    k_grid_build<<<n_frames, 512, 0, h->stream>>>(f_slab, A.g, A.f_xy, A.f_counts, h->cell_start.as<int>(), h->cell_items.as<int>());
    if (f_slab <= 2048) k_search_candidates<true><<<dim3((q_slab + 7) / 8, n_frames), 256, 0, h->stream>>>(A);
    else k_search_candidates<false><<<dim3((q_slab + 7) / 8, n_frames), 256, 0, h->stream>>>(A);
    {
        size_t smem = ((size_t)q_slab * (kTop + 3) + 2 * (size_t)f_slab) * sizeof(int);
        const int use_smem = smem <= 200 * 1024;
        if (!use_smem) smem = (size_t)q_slab * sizeof(int);                    // rescan queue only
        ORBS_REQUIRE(smem <= 200 * 1024, ORBS_E_INVALID, "too many map points per frame for the projection search");
        // the opt-in limit is a per-device attribute of the FUNCTION, not of a handle: it is raised once, to the static upper
        // bound, in orbm_create and never lowered (handles of different sizes share it)
        k_search_resolve<<<n_frames, 1024, smem, h->stream>>>(A, use_smem);
    }
    h->launches += 3;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_project_points(orbm_handle *h, int n_views, const orbm_projection *proj, const float *scale_factors, int nlevels,
                        const float *Xw, const float *normal, const float *mf_min_distance, const float *mf_max_distance,
                        const int32_t *q_counts, int q_slab, uint8_t *q_valid, float *q_uv, float *q_radius, int32_t *q_minl,
                        int32_t *q_maxl, int32_t *q_level, int memspace)
{
    ORBS_REQUIRE(h && proj && scale_factors && Xw && mf_min_distance && mf_max_distance && q_counts && q_valid && q_uv && q_radius && q_minl && q_maxl,
                 ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_views > 0 && q_slab > 0 && nlevels > 0, ORBS_E_INVALID, "non-positive size");
    for (int f = 0; f < n_views; f++)
        ORBS_REQUIRE(normal || !(proj[f].flags & ORBM_PROJ_CHECK_NORMAL), ORBS_E_INVALID, "ORBM_PROJ_CHECK_NORMAL needs the normals");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t nq = (size_t)n_views * q_slab;
    ProjView *dviews = S.scratch<ProjView>(n_views);                       // the parameter blocks are host arrays in both memory spaces
    if (S.rc) return S.rc;
    ORBS_CUDA(cudaMemcpyAsync(dviews, proj, sizeof(ProjView) * (size_t)n_views, cudaMemcpyHostToDevice, h->stream));
    const float *dsf = S.in(scale_factors, nlevels), *dX = S.in(Xw, nq * 3), *dN = normal ? S.in(normal, nq * 3) : nullptr;
    const float *dmn = S.in(mf_min_distance, nq), *dmx = S.in(mf_max_distance, nq);
    const int32_t *dqc = S.in(q_counts, n_views);
    uint8_t *dval = S.inout(q_valid, nq);
    float *duv = S.inout(q_uv, nq * 2, false), *drad = S.inout(q_radius, nq, false);
    int32_t *dl0 = S.inout(q_minl, nq, false), *dl1 = S.inout(q_maxl, nq, false), *dlv = q_level ? S.inout(q_level, nq, false) : nullptr;
    if (S.rc) return S.rc;
    if (memspace == ORBS_MEM_HOST) {       // rejected points keep defined values on the host
        ORBS_CUDA(cudaMemsetAsync(duv, 0, nq * 8, h->stream)); ORBS_CUDA(cudaMemsetAsync(drad, 0, nq * 4, h->stream));
        ORBS_CUDA(cudaMemsetAsync(dl0, 0, nq * 4, h->stream)); ORBS_CUDA(cudaMemsetAsync(dl1, 0, nq * 4, h->stream));
        if (dlv) ORBS_CUDA(cudaMemsetAsync(dlv, 0, nq * 4, h->stream));
    }
    k_project_points<<<dim3((q_slab + 255) / 256, n_views), 256, 0, h->stream>>>(q_slab, dviews, dsf, nlevels, dX, dN, dmn, dmx, dqc, dval, (float2 *)duv,
                                                                                 drad, dl0, dl1, dlv);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    if (memspace == ORBS_MEM_DEVICE) ORBS_CUDA(cudaStreamSynchronize(h->stream));   // proj is a host array the caller may reuse
    return S.finish();
}

int orbm_search_best_in_window(orbm_handle *h, int n_frames, const float *grid_bounds4, const float *win_origin2,
                               const float *f_xy, const int32_t *f_octave, const uint8_t *f_desc, const int32_t *f_counts, int f_slab,
                               const uint8_t *q_valid, const float *q_uv, const float *q_radius, const int32_t *q_minl,
                               const int32_t *q_maxl, const uint8_t *q_desc, const int32_t *q_counts, int q_slab, int th_dist,
                               const float *inv_level_sigma2, int nlevels, float chi2_gate,
                               int32_t *q_best_idx, int32_t *q_best_dist, int memspace)
{
    ORBS_REQUIRE(h && grid_bounds4 && f_xy && f_octave && f_desc && f_counts && q_valid && q_uv && q_radius && q_minl && q_maxl && q_desc && q_counts &&
                 q_best_idx && q_best_dist, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && f_slab > 0 && q_slab > 0, ORBS_E_INVALID, "non-positive size");
    ORBS_REQUIRE(f_slab < (1 << 20), ORBS_E_INVALID, "slab too large (features < 2^20 per frame)");
    ORBS_REQUIRE(!inv_level_sigma2 || (nlevels > 0 && nlevels <= 32), ORBS_E_INVALID, "1..32 pyramid levels");
    ORBS_REQUIRE(grid_bounds4[2] > grid_bounds4[0] && grid_bounds4[3] > grid_bounds4[1], ORBS_E_INVALID, "empty image bounds");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t nf = (size_t)n_frames * f_slab, nq = (size_t)n_frames * q_slab;
    BestArgs A;
    A.f_slab = f_slab; A.q_slab = q_slab; A.g = make_grid(grid_bounds4);
    if (win_origin2) { A.g.win_x = win_origin2[0]; A.g.win_y = win_origin2[1]; }
    A.f_xy = (const float2 *)S.in(f_xy, nf * 2); A.f_octave = S.in(f_octave, nf); A.f_desc = (const uint4 *)S.in(f_desc, nf * 32);
    A.f_counts = S.in(f_counts, n_frames);
    A.q_valid = S.in(q_valid, nq); A.q_uv = (const float2 *)S.in(q_uv, nq * 2); A.q_radius = S.in(q_radius, nq);
    A.q_minl = S.in(q_minl, nq); A.q_maxl = S.in(q_maxl, nq); A.q_desc = (const uint4 *)S.in(q_desc, nq * 32); A.q_counts = S.in(q_counts, n_frames);
    A.best_idx = S.inout(q_best_idx, nq, false); A.best_dist = S.inout(q_best_dist, nq, false);
    if (S.rc) return S.rc;
    ORBS_REQUIRE(((uintptr_t)A.f_desc % 16 == 0) && ((uintptr_t)A.q_desc % 16 == 0) && ((uintptr_t)A.f_xy % 8 == 0) && ((uintptr_t)A.q_uv % 8 == 0),
                 ORBS_E_INVALID, "descriptor arrays must be 16-byte aligned, coordinate arrays 8-byte aligned");
    int rc;
    if ((rc = h->cell_start.reserve((size_t)n_frames * (kGridCells + 1) * sizeof(int)))) return rc;
    if ((rc = h->cell_items.reserve(nf * sizeof(int)))) return rc;
    A.cell_start = h->cell_start.as<int>(); A.cell_items = h->cell_items.as<int>();
    A.th_dist = th_dist; A.use_gate = inv_level_sigma2 ? 1 : 0; A.nlevels = inv_level_sigma2 ? nlevels : 1; A.chi2_gate = chi2_gate;
    for (int l = 0; l < 32; l++) A.inv_sigma2[l] = (inv_level_sigma2 && l < nlevels) ? inv_level_sigma2[l] : 0.f;
    if (memspace == ORBS_MEM_HOST) { ORBS_CUDA(cudaMemsetAsync(A.best_idx, 0xff, nq * 4, h->stream)); ORBS_CUDA(cudaMemsetAsync(A.best_dist, 0xff, nq * 4, h->stream)); }
    k_grid_build<<<n_frames, 512, 0, h->stream>>>(f_slab, A.g, A.f_xy, A.f_counts, h->cell_start.as<int>(), h->cell_items.as<int>());
    k_search_best<<<dim3((q_slab + 7) / 8, n_frames), 256, 0, h->stream>>>(A);
    h->launches += 2;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_search_by_bow(orbm_handle *h, int n_pairs, int mode,
                       const uint8_t *desc1, const float *angle1, const uint8_t *elig1, const int32_t *counts1, int slab1,
                       const int32_t *fv1_nodes, const int32_t *fv1_start, const int32_t *fv1_items, const int32_t *fv1_counts, int fv1_slab,
                       const uint8_t *desc2, const float *angle2, const uint8_t *elig2, const int32_t *counts2, int slab2,
                       const int32_t *fv2_nodes, const int32_t *fv2_start, const int32_t *fv2_items, const int32_t *fv2_counts, int fv2_slab,
                       float ratio, int check_ori, const orbm_epipolar *epi, int32_t *match12, int32_t *nmatches, int memspace)
{
    ORBS_REQUIRE(h && desc1 && elig1 && counts1 && fv1_nodes && fv1_start && fv1_items && fv1_counts && desc2 && elig2 && counts2 && fv2_nodes && fv2_start &&
                 fv2_items && fv2_counts && match12 && nmatches, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(mode == ORBM_BOW_MATCH || mode == ORBM_BOW_TRIANGULATION, ORBS_E_INVALID, "unknown mode");
    ORBS_REQUIRE(!check_ori || (angle1 && angle2), ORBS_E_INVALID, "orientation check needs the angles");
    ORBS_REQUIRE(mode != ORBM_BOW_TRIANGULATION || (epi && epi->xy1 && epi->xy2 && epi->octave2 && epi->F12 && epi->epipole && epi->scale_factors2 &&
                 epi->level_sigma2_2 && epi->nlevels > 0 && epi->nlevels <= 32), ORBS_E_INVALID, "triangulation mode needs the epipolar block (1..32 levels)");
    ORBS_REQUIRE(n_pairs > 0 && slab1 > 0 && slab2 > 0 && fv1_slab > 0 && fv2_slab > 0, ORBS_E_INVALID, "non-positive size");
    ORBS_REQUIRE(slab2 < (1 << 20), ORBS_E_INVALID, "slab too large (features < 2^20 per keyframe)");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t n1 = (size_t)n_pairs * slab1, n2 = (size_t)n_pairs * slab2;
    BowArgs A;
    A.mode = mode; A.slab1 = slab1; A.slab2 = slab2; A.nslab1 = fv1_slab; A.nslab2 = fv2_slab;
    A.desc1 = (const uint4 *)S.in(desc1, n1 * 32); A.desc2 = (const uint4 *)S.in(desc2, n2 * 32);
    A.angle1 = angle1 ? S.in(angle1, n1) : nullptr; A.angle2 = angle2 ? S.in(angle2, n2) : nullptr;
    A.elig1 = S.in(elig1, n1); A.elig2 = S.in(elig2, n2); A.counts1 = S.in(counts1, n_pairs);
    A.nodes1 = S.in(fv1_nodes, (size_t)n_pairs * fv1_slab); A.start1 = S.in(fv1_start, (size_t)n_pairs * (fv1_slab + 1)); A.items1 = S.in(fv1_items, n1);
    A.ncount1 = S.in(fv1_counts, n_pairs);
    A.nodes2 = S.in(fv2_nodes, (size_t)n_pairs * fv2_slab); A.start2 = S.in(fv2_start, (size_t)n_pairs * (fv2_slab + 1)); A.items2 = S.in(fv2_items, n2);
    A.ncount2 = S.in(fv2_counts, n_pairs);
    A.ratio = ratio; A.check_ori = check_ori ? 1 : 0;
    A.xy1 = A.xy2 = nullptr; A.octave2 = nullptr; A.F12 = A.epipole = nullptr; A.nlevels = 1;
    for (int l = 0; l < 32; l++) { A.scale_factors2[l] = 0.f; A.level_sigma2_2[l] = 0.f; }
    if (mode == ORBM_BOW_TRIANGULATION) {
        A.xy1 = (const float2 *)S.in(epi->xy1, n1 * 2); A.xy2 = (const float2 *)S.in(epi->xy2, n2 * 2); A.octave2 = S.in(epi->octave2, n2);
        A.F12 = S.in(epi->F12, (size_t)n_pairs * 9); A.epipole = S.in(epi->epipole, (size_t)n_pairs * 2);
        A.nlevels = epi->nlevels;
        for (int l = 0; l < epi->nlevels; l++) { A.scale_factors2[l] = epi->scale_factors2[l]; A.level_sigma2_2[l] = epi->level_sigma2_2[l]; }
    }
    A.matched2 = S.scratch<uint8_t>(n2); A.rot_bin = S.scratch<int>(n1);
    A.match12 = S.inout(match12, n1, false); A.nmatches = S.inout(nmatches, n_pairs, false);
    if (S.rc) return S.rc;
    ORBS_REQUIRE(((uintptr_t)A.desc1 % 16 == 0) && ((uintptr_t)A.desc2 % 16 == 0), ORBS_E_INVALID, "descriptor arrays must be 16-byte aligned");
    ORBS_CUDA(cudaMemsetAsync(A.matched2, 0, n2, h->stream));
    ORBS_CUDA(cudaMemsetAsync(A.match12, 0xff, n1 * sizeof(int), h->stream));
    k_bow_match<<<dim3((fv1_slab + 7) / 8, n_pairs), 256, 0, h->stream>>>(A);
    k_bow_finish<<<n_pairs, 256, 0, h->stream>>>(slab1, A.counts1, A.check_ori, A.match12, A.rot_bin, A.nmatches);
    h->launches += 2;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_search_for_initialization(orbm_handle *h, int n_pairs, const float *bounds4,
                                   const int32_t *octave1, const float *angle1, const uint8_t *desc1, const int32_t *counts1, int slab1,
                                   const float *xy2, const int32_t *octave2, const float *angle2, const uint8_t *desc2, const int32_t *counts2, int slab2,
                                   float *prev_matched, int window, float ratio, int check_ori, int32_t *matches12, int32_t *nmatches, int memspace)
{
    ORBS_REQUIRE(h && bounds4 && octave1 && desc1 && counts1 && xy2 && octave2 && desc2 && counts2 && prev_matched && matches12 && nmatches, ORBS_E_INVALID,
                 "null argument");
    ORBS_REQUIRE(!check_ori || (angle1 && angle2), ORBS_E_INVALID, "orientation check needs the angles");
    ORBS_REQUIRE(n_pairs > 0 && slab1 > 0 && slab2 > 0, ORBS_E_INVALID, "non-positive size");
    ORBS_REQUIRE(slab2 < (1 << 20), ORBS_E_INVALID, "slab too large (features < 2^20 per frame)");
    ORBS_REQUIRE(bounds4[2] > bounds4[0] && bounds4[3] > bounds4[1], ORBS_E_INVALID, "empty image bounds");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t n1 = (size_t)n_pairs * slab1, n2 = (size_t)n_pairs * slab2;
    InitArgs A;
    A.slab1 = slab1; A.slab2 = slab2; A.g = make_grid(bounds4);
    A.octave1 = S.in(octave1, n1); A.angle1 = angle1 ? S.in(angle1, n1) : nullptr; A.desc1 = (const uint4 *)S.in(desc1, n1 * 32); A.counts1 = S.in(counts1, n_pairs);
    A.xy2 = (const float2 *)S.in(xy2, n2 * 2); A.octave2 = S.in(octave2, n2); A.angle2 = angle2 ? S.in(angle2, n2) : nullptr;
    A.desc2 = (const uint4 *)S.in(desc2, n2 * 32); A.counts2 = S.in(counts2, n_pairs);
    A.prev_matched = (float2 *)S.inout(prev_matched, n1 * 2);
    A.window = (float)window; A.ratio = ratio; A.check_ori = check_ori ? 1 : 0;
    A.matched_dist = S.scratch<int>(n2); A.matches21 = S.scratch<int>(n2); A.rot_bin = S.scratch<int>(n1);
    A.matches12 = S.inout(matches12, n1, false); A.nmatches = S.inout(nmatches, n_pairs, false);
    if (S.rc) return S.rc;
    ORBS_REQUIRE(((uintptr_t)A.desc1 % 16 == 0) && ((uintptr_t)A.desc2 % 16 == 0) && ((uintptr_t)A.xy2 % 8 == 0) && ((uintptr_t)A.prev_matched % 8 == 0),
                 ORBS_E_INVALID, "descriptor arrays must be 16-byte aligned, coordinate arrays 8-byte aligned");
    int rc;
    if ((rc = h->cell_start.reserve((size_t)n_pairs * (kGridCells + 1) * sizeof(int)))) return rc;
    if ((rc = h->cell_items.reserve(n2 * sizeof(int)))) return rc;
    A.cell_start = h->cell_start.as<int>(); A.cell_items = h->cell_items.as<int>();
    if (memspace == ORBS_MEM_HOST) ORBS_CUDA(cudaMemsetAsync(A.matches12, 0xff, n1 * sizeof(int), h->stream));
    k_grid_build<<<n_pairs, 512, 0, h->stream>>>(slab2, A.g, A.xy2, A.counts2, h->cell_start.as<int>(), h->cell_items.as<int>());
    k_search_init<<<n_pairs, kInitThreads, 0, h->stream>>>(A);
    h->launches += 2;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_assign_features_to_grid(orbm_handle *h, int n_frames, const float *bounds4, const float *f_xy, const int32_t *f_counts, int f_slab,
                                 int32_t *cell_start, int32_t *cell_items, int memspace)
{
    ORBS_REQUIRE(h && bounds4 && f_xy && f_counts && cell_start && cell_items, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && f_slab > 0, ORBS_E_INVALID, "non-positive size");
    ORBS_REQUIRE(bounds4[2] > bounds4[0] && bounds4[3] > bounds4[1], ORBS_E_INVALID, "empty image bounds");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t nf = (size_t)n_frames * f_slab;
    const float2 *dxy = (const float2 *)S.in(f_xy, nf * 2);
    const int32_t *dc = S.in(f_counts, n_frames);
    int32_t *dcs = S.inout(cell_start, (size_t)n_frames * (kGridCells + 1), false), *dci = S.inout(cell_items, nf, false);
    if (S.rc) return S.rc;
    if (memspace == ORBS_MEM_HOST) ORBS_CUDA(cudaMemsetAsync(dci, 0xff, nf * sizeof(int), h->stream));
    k_grid_build<<<n_frames, 512, 0, h->stream>>>(f_slab, make_grid(bounds4), dxy, dc, dcs, dci);
    k_grid_sort_cells<<<dim3((kGridCells + 255) / 256, n_frames), 256, 0, h->stream>>>(f_slab, dcs, dci);
    h->launches += 2;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbm_get_features_in_area(orbm_handle *h, int n_frames, const float *bounds4, const float *win_origin2, const float *f_xy, const int32_t *f_octave,
                              const int32_t *f_counts, int f_slab, const float *q_xyr, const int32_t *q_minl, const int32_t *q_maxl,
                              const int32_t *q_counts, int q_slab, int cap, int32_t *out_idx, int32_t *out_count, int memspace)
{
    ORBS_REQUIRE(h && bounds4 && f_xy && f_counts && q_xyr && q_counts && out_idx && out_count, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(f_octave || (!q_minl && !q_maxl), ORBS_E_INVALID, "level limits need the feature octaves");
    ORBS_REQUIRE(n_frames > 0 && f_slab > 0 && q_slab > 0 && cap > 0, ORBS_E_INVALID, "non-positive size");
    ORBS_REQUIRE(bounds4[2] > bounds4[0] && bounds4[3] > bounds4[1], ORBS_E_INVALID, "empty image bounds");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t nf = (size_t)n_frames * f_slab, nq = (size_t)n_frames * q_slab;
    GridParams g = make_grid(bounds4);
    if (win_origin2) { g.win_x = win_origin2[0]; g.win_y = win_origin2[1]; }
    const float2 *dxy = (const float2 *)S.in(f_xy, nf * 2);
    const int32_t *doct = f_octave ? S.in(f_octave, nf) : nullptr, *dc = S.in(f_counts, n_frames);
    const float *dq = S.in(q_xyr, nq * 3);
    const int32_t *dmn = q_minl ? S.in(q_minl, nq) : nullptr, *dmx = q_maxl ? S.in(q_maxl, nq) : nullptr, *dqc = S.in(q_counts, n_frames);
    int32_t *dout = S.inout(out_idx, nq * cap, false), *dcnt = S.inout(out_count, nq, false);
    if (S.rc) return S.rc;
    int rc;
    if ((rc = h->cell_start.reserve((size_t)n_frames * (kGridCells + 1) * sizeof(int)))) return rc;
    if ((rc = h->cell_items.reserve(nf * sizeof(int)))) return rc;
    if (memspace == ORBS_MEM_HOST) { ORBS_CUDA(cudaMemsetAsync(dout, 0xff, nq * cap * sizeof(int), h->stream)); ORBS_CUDA(cudaMemsetAsync(dcnt, 0, nq * sizeof(int), h->stream)); }
    k_grid_build<<<n_frames, 512, 0, h->stream>>>(f_slab, g, dxy, dc, h->cell_start.as<int>(), h->cell_items.as<int>());
    k_grid_sort_cells<<<dim3((kGridCells + 255) / 256, n_frames), 256, 0, h->stream>>>(f_slab, h->cell_start.as<int>(), h->cell_items.as<int>());
    k_features_in_area<<<dim3((q_slab + 7) / 8, n_frames), 256, 0, h->stream>>>(f_slab, q_slab, cap, g, dxy, doct, h->cell_start.as<int>(), h->cell_items.as<int>(),
                                                                              dq, dmn, dmx, dqc, dout, dcnt);
    h->launches += 3;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

}  // extern "C"
