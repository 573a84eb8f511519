/*
 * oracle/slam_oracle.c -- CPU restatement of the reference matcher and optimizer arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into or called by the product.
 *
 * Matcher part follows (S/ = /root/reference/SingleRobotScenario/):
 *   S/src/ORBmatcher.cc:45-137     SearchByProjection(Frame&, vector<MapPoint*>&, th)
 *   S/src/ORBmatcher.cc:1330-1472  SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)
 *   S/src/ORBmatcher.cc:1603-1665  ComputeThreeMaxima, DescriptorDistance
 *   S/src/ORBmatcher.cc:292-405, 827-977, 979-1102, 1104-1328, 1474-1601  the KeyFrame / Sim3 projection family
 *       (SearchByProjection(KF, Scw), Fuse x2, SearchBySim3, SearchByProjection(Frame, KF)); S/src/KeyFrame.cc:618-662
 *   S/src/Frame.cc:230-245,327-392 AssignFeaturesToGrid, GetFeaturesInArea, PosInGrid
 * over flat arrays instead of Frame/MapPoint objects (monocular only: mvuRight < 0, every map point
 * has Observations() > 0).  cv::Mat float algebra (Rcw*x3Dw+tcw) is restated as OpenCV 4.13's small
 * gemm: fp32 products accumulated left to right in fp32, no FMA (pinned by tests against cv2.gemm).
 *
 * Optimizer part follows
 *   S/src/Optimizer.cc:262-474 (PoseOptimization), :476-801 (LocalBundleAdjustment), :68-260 (BundleAdjustment)
 * and the vendored g2o it drives (all fp64):
 *   types/types_six_dof_expmap.{h,cpp}, types/se3quat.h, types/types_sba.h, core/robust_kernel_impl.cpp,
 *   core/base_edge.h, core/base_{unary,binary}_edge.hpp, core/sparse_optimizer.cpp,
 *   core/block_solver.hpp, core/optimization_algorithm_levenberg.cpp, solvers/linear_solver_{dense,eigen}.h
 * Eigen (un-vendored, >=3.1, unpinned) supplies Quaterniond(R)/normalize/toRotationMatrix, Matrix3d::inverse
 * and the LDLT factorisations; they are direct methods restated here (reduced system: profile LDLT without
 * pivoting instead of SimplicialLDLT+AMD / pivoted dense LDLT -> equal up to fp64 rounding).  The reference
 * holds no golden vectors for this path: parity unpinned by the reference; this file is pinned against an
 * independent numpy twin (tests/test_oracle_ba.py).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================== */
/* Matcher                                                                                   */
#define GRID_COLS 64
#define GRID_ROWS 48
#define TH_HIGH 100
#define TH_LOW 50
#define HISTO_LENGTH 30

int oracle_descriptor_distance(const uint8_t *a, const uint8_t *b)     /* ORBmatcher.cc:1649-1665 */
{
    const int32_t *pa = (const int32_t *)a, *pb = (const int32_t *)b;
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

typedef struct {
    float min_x, min_y, max_x, max_y;      /* mnMinX.. (Frame.cc:436-463) */
    float w_inv, h_inv;                    /* mfGridElementWidthInv / HeightInv (Frame.cc:101-102) */
} oracle_grid_params;

void oracle_grid_params_init(oracle_grid_params *g, float min_x, float min_y, float max_x, float max_y)
{
    g->min_x = min_x; g->min_y = min_y; g->max_x = max_x; g->max_y = max_y;
    g->w_inv = (float)GRID_COLS / (max_x - min_x);
    g->h_inv = (float)GRID_ROWS / (max_y - min_y);
}

/* AssignFeaturesToGrid + PosInGrid: CSR over cells indexed ix*GRID_ROWS+iy, items in ascending feature index */
void oracle_grid_build(const oracle_grid_params *g, int N, const float *xy, int *cell_start /*[64*48+1]*/, int *cell_items /*[N]*/)
{
    const int nc = GRID_COLS * GRID_ROWS;
    int *cell_of = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    memset(cell_start, 0, sizeof(int) * (nc + 1));
    for (int i = 0; i < N; i++) {
        int px = (int)roundf((xy[2 * i] - g->min_x) * g->w_inv);
        int py = (int)roundf((xy[2 * i + 1] - g->min_y) * g->h_inv);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) { cell_of[i] = -1; continue; }
        cell_of[i] = px * GRID_ROWS + py;
        cell_start[cell_of[i] + 1]++;
    }
    for (int c = 0; c < nc; c++) cell_start[c + 1] += cell_start[c];
    int *fill = (int *)calloc(nc, sizeof(int));
    for (int i = 0; i < N; i++)
        if (cell_of[i] >= 0) cell_items[cell_start[cell_of[i]] + fill[cell_of[i]]++] = i;
    free(fill); free(cell_of);
}

/* GetFeaturesInArea (Frame.cc:327-380); returns count, indices in the reference's visiting order */
int oracle_features_in_area(const oracle_grid_params *g, const int *cell_start, const int *cell_items,
                            const float *xy, const int *octave, float x, float y, float r,
                            int min_level, int max_level, int *out)
{
    int n = 0;
    int c0 = (int)floorf((x - g->min_x - r) * g->w_inv); if (c0 < 0) c0 = 0;
    if (c0 >= GRID_COLS) return 0;
    int c1 = (int)ceilf((x - g->min_x + r) * g->w_inv); if (c1 > GRID_COLS - 1) c1 = GRID_COLS - 1;
    if (c1 < 0) return 0;
    int r0 = (int)floorf((y - g->min_y - r) * g->h_inv); if (r0 < 0) r0 = 0;
    if (r0 >= GRID_ROWS) return 0;
    int r1 = (int)ceilf((y - g->min_y + r) * g->h_inv); if (r1 > GRID_ROWS - 1) r1 = GRID_ROWS - 1;
    if (r1 < 0) return 0;
    const int check = (min_level > 0) || (max_level >= 0);
    for (int ix = c0; ix <= c1; ix++)
        for (int iy = r0; iy <= r1; iy++) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = cell_start[c]; j < cell_start[c + 1]; j++) {
                const int k = cell_items[j];
                if (check) {
                    if (octave[k] < min_level) continue;
                    if (max_level >= 0 && octave[k] > max_level) continue;
                }
                const float dx = xy[2 * k] - x, dy = xy[2 * k + 1] - y;
                if (fabsf(dx) < r && fabsf(dy) < r) out[n++] = k;
            }
        }
    return n;
}

/* Projection block of SearchByProjection(Cur, Last), ORBmatcher.cc:1354-1391 (mono: bForward=bBackward=false).
 * q_valid[i] in: last-frame slot holds a non-outlier map point; out: and it projects into the image with
 * positive depth.  Tcw row-major 4x4 float. */
void oracle_project_last_frame(const float *Tcw, const float *K4 /*fx,fy,cx,cy*/, const oracle_grid_params *g,
                               const float *scale_factors, int M, const float *Xw, const int *last_octave,
                               float th, uint8_t *q_valid, float *q_uv, float *q_radius, int *q_minl, int *q_maxl)
{
    for (int i = 0; i < M; i++) {
        if (!q_valid[i]) continue;
        q_valid[i] = 0;
        float xc[3];
        for (int r = 0; r < 3; r++) {       /* cv::gemm small-matrix path: fp32, sequential, no FMA */
            float s = Tcw[4 * r] * Xw[3 * i];
            s = s + Tcw[4 * r + 1] * Xw[3 * i + 1];
            s = s + Tcw[4 * r + 2] * Xw[3 * i + 2];
            xc[r] = s + Tcw[4 * r + 3];
        }
        const float invzc = (float)(1.0 / xc[2]);
        if (invzc < 0) continue;
        const float u = K4[0] * xc[0] * invzc + K4[2];
        const float v = K4[1] * xc[1] * invzc + K4[3];
        if (u < g->min_x || u > g->max_x) continue;
        if (v < g->min_y || v > g->max_y) continue;
        const int oct = last_octave[i];
        q_uv[2 * i] = u; q_uv[2 * i + 1] = v;
        q_radius[i] = th * scale_factors[oct];
        q_minl[i] = oct - 1; q_maxl[i] = oct + 1;
        q_valid[i] = 1;
    }
}

static void three_maxima(const int *hist, int L, int *i1, int *i2, int *i3)   /* ORBmatcher.cc:1603-1644 */
{
    int max1 = 0, max2 = 0, max3 = 0;
    *i1 = *i2 = *i3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *i3 = *i2; *i2 = *i1; *i1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *i3 = *i2; *i2 = i; }
        else if (s > max3) { max3 = s; *i3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { *i2 = -1; *i3 = -1; }
    else if (max3 < 0.1f * (float)max1) { *i3 = -1; }
}

/* Generic projection search over flat arrays; the sequential loop of ORBmatcher.cc:49-126 (ratio > 0: best and
 * second best with the same-level ratio test, no orientation check) and :1353-1467 (ratio <= 0, check_ori).
 * feat_match[N]: in = -1 or an id >= 0 for features that already hold a map point (they are skipped);
 *                out = index of the query assigned to the feature.
 * Returns nmatches. */
int oracle_search_by_projection_kf(const oracle_grid_params *g, const float *win_origin2, int N, const float *f_xy, const int *f_octave,
                                   const float *f_angle, const uint8_t *f_desc,
                                   int M, const uint8_t *q_valid, const float *q_uv, const float *q_radius,
                                   const int *q_minl, const int *q_maxl, const float *q_angle, const uint8_t *q_desc,
                                   int th_dist, float ratio, int check_ori, int *feat_match);
int oracle_search_by_projection(const oracle_grid_params *g, int N, const float *f_xy, const int *f_octave,
                                const float *f_angle, const uint8_t *f_desc,
                                int M, const uint8_t *q_valid, const float *q_uv, const float *q_radius,
                                const int *q_minl, const int *q_maxl, const float *q_angle, const uint8_t *q_desc,
                                int th_dist, float ratio, int check_ori, int *feat_match)
{
    return oracle_search_by_projection_kf(g, NULL, N, f_xy, f_octave, f_angle, f_desc, M, q_valid, q_uv, q_radius, q_minl, q_maxl, q_angle, q_desc,
                                          th_dist, ratio, check_ori, feat_match);
}

/* win_origin2 != NULL: the target is a KeyFrame, whose GetFeaturesInArea (KeyFrame.cc:618-657) offsets by its integer mnMinX / mnMinY
 * while the grid was filled by the Frame with its float bounds (the claim loops of :292-405 and :1474-1601 use this form). */
int oracle_search_by_projection_kf(const oracle_grid_params *g, const float *win_origin2, int N, const float *f_xy, const int *f_octave,
                                   const float *f_angle, const uint8_t *f_desc,
                                   int M, const uint8_t *q_valid, const float *q_uv, const float *q_radius,
                                   const int *q_minl, const int *q_maxl, const float *q_angle, const uint8_t *q_desc,
                                   int th_dist, float ratio, int check_ori, int *feat_match)
{
    oracle_grid_params gw = *g;
    if (win_origin2) { gw.min_x = win_origin2[0]; gw.min_y = win_origin2[1]; }
    const int nc = GRID_COLS * GRID_ROWS;
    int *cell_start = (int *)malloc(sizeof(int) * (nc + 1));
    int *cell_items = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    int *cands = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    int *rot_bin = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));     /* per feature: histogram bin of its match */
    int hist[HISTO_LENGTH];
    memset(hist, 0, sizeof hist);
    oracle_grid_build(g, N, f_xy, cell_start, cell_items);
    for (int i = 0; i < N; i++) rot_bin[i] = -1;
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    for (int i = 0; i < M; i++) {
        if (!q_valid[i]) continue;
        const int nc_i = oracle_features_in_area(&gw, cell_start, cell_items, f_xy, f_octave, q_uv[2 * i], q_uv[2 * i + 1],
                                                 q_radius[i], q_minl[i], q_maxl[i], cands);
        if (nc_i == 0) continue;
        int best = 256, best2 = 256, best_level = -1, best_level2 = -1, best_idx = -1;
        for (int c = 0; c < nc_i; c++) {
            const int k = cands[c];
            if (feat_match[k] >= 0) continue;                 /* already holds a map point with observations */
            const int d = oracle_descriptor_distance(q_desc + 32 * (size_t)i, f_desc + 32 * (size_t)k);
            if (d < best) { best2 = best; best = d; best_level2 = best_level; best_level = f_octave[k]; best_idx = k; }
            else if (d < best2) { best_level2 = f_octave[k]; best2 = d; }
        }
        if (best <= th_dist) {
            if (ratio > 0 && best_level == best_level2 && best > ratio * best2) continue;
            feat_match[best_idx] = i;
            nmatches++;
            if (check_ori) {
                float rot = q_angle[i] - f_angle[best_idx];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rot_bin[best_idx] = bin;
                hist[bin]++;
            }
        }
    }
    if (check_ori) {
        int i1, i2, i3;
        three_maxima(hist, HISTO_LENGTH, &i1, &i2, &i3);
        for (int k = 0; k < N; k++) {
            const int b = rot_bin[k];
            if (b >= 0 && b != i1 && b != i2 && b != i3) { feat_match[k] = -1; nmatches--; }
        }
    }
    free(cell_start); free(cell_items); free(cands); free(rot_bin);
    return nmatches;
}

/* ---------------------------------------------------------------------------------------- */
/* KeyFrame / Sim3 projection family.  The five members (ORBmatcher.cc:292-405 SearchByProjection(KF, Scw, ...), :827-977 Fuse(KF,
 * vpMapPoints, th), :979-1102 Fuse(KF, Scw, ...), :1104-1328 SearchBySim3, :1474-1601 SearchByProjection(Frame, KF, ...)) share a
 * projection block that differs only in which tests are made; the flags name the differences (values = ORBM_PROJ_* of
 * include/orbslamm_b200.h). */
#define PROJ_TWO_STEP 0x01
#define PROJ_NO_DEPTH 0x02
#define PROJ_FRAME_BOUNDS 0x04
#define PROJ_FRAME_UV 0x08
#define PROJ_DIST_CAMERA 0x10
#define PROJ_CHECK_NORMAL 0x20
#define PROJ_LEVEL_PLUS1 0x40

typedef struct {
    float R[9], t[3], R2[9], t2[3], Ow[3];
    float fx, fy, cx, cy, min_x, min_y, max_x, max_y, log_scale_factor, th;
    int32_t flags;
} oracle_projection;

static void rt_apply(const float *R, const float *t, const float *x, float *out)   /* cv::Mat R*x + t: small gemm, then add */
{
    for (int r = 0; r < 3; r++) {
        float s = R[3 * r] * x[0];
        s = s + R[3 * r + 1] * x[1];
        s = s + R[3 * r + 2] * x[2];
        out[r] = s + t[r];
    }
}

void oracle_project_points(const oracle_projection *V, const float *scale_factors, int nlevels, int M, const float *Xw, const float *normal,
                           const float *mf_min_dist, const float *mf_max_dist, uint8_t *q_valid, float *q_uv, float *q_radius,
                           int *q_minl, int *q_maxl, int *q_level)
{
    for (int i = 0; i < M; i++) {
        if (!q_valid[i]) continue;
        q_valid[i] = 0;
        const float *X = Xw + 3 * i;
        float pc[3];
        rt_apply(V->R, V->t, X, pc);                                       /* :326 / :850 / :1017 / :1155 / :1503 */
        if (V->flags & PROJ_TWO_STEP) { float p1[3] = {pc[0], pc[1], pc[2]}; rt_apply(V->R2, V->t2, p1, pc); }   /* :1156, :1236 */
        if (!(V->flags & PROJ_NO_DEPTH) && pc[2] < 0.0f) continue;         /* :329, :853, :1020, :1159 */
        const float invz = 1.0f / pc[2];
        float u, v;
        if (V->flags & PROJ_FRAME_UV) {                                    /* :1510-1511 */
            u = V->fx * pc[0] * invz + V->cx;
            v = V->fy * pc[1] * invz + V->cy;
        } else {                                                           /* :333-338 */
            const float x = pc[0] * invz, y = pc[1] * invz;
            u = V->fx * x + V->cx;
            v = V->fy * y + V->cy;
        }
        if (V->flags & PROJ_FRAME_BOUNDS) {                                /* :1513-1516 */
            if (u < V->min_x || u > V->max_x) continue;
            if (v < V->min_y || v > V->max_y) continue;
        } else if (!(u >= V->min_x && u < V->max_x && v >= V->min_y && v < V->max_y)) continue;   /* KeyFrame::IsInImage */
        const float maxd = 1.2f * mf_max_dist[i], mind = 0.8f * mf_min_dist[i];   /* MapPoint.cc:373-383 */
        float po[3];
        if (V->flags & PROJ_DIST_CAMERA) { po[0] = pc[0]; po[1] = pc[1]; po[2] = pc[2]; }          /* :1179 */
        else { po[0] = X[0] - V->Ow[0]; po[1] = X[1] - V->Ow[1]; po[2] = X[2] - V->Ow[2]; }        /* :347 */
        double s2 = 0;
        for (int k = 0; k < 3; k++) s2 += (double)po[k] * (double)po[k];
        const float dist = (float)sqrt(s2);
        if (dist < mind || dist > maxd) continue;
        if (V->flags & PROJ_CHECK_NORMAL) {                                /* :355-358 */
            double dot = 0;
            for (int k = 0; k < 3; k++) dot += (double)po[k] * (double)normal[3 * i + k];
            if (dot < 0.5 * dist) continue;
        }
        const float ratio = mf_max_dist[i] / dist;                         /* MapPoint::PredictScale */
        const int level = (int)ceilf(logf(ratio) / V->log_scale_factor);
        if (level < 0 || level >= nlevels) continue;                       /* reference: out-of-range index into mvScaleFactors (undefined) */
        q_uv[2 * i] = u; q_uv[2 * i + 1] = v;
        q_radius[i] = V->th * scale_factors[level];
        q_minl[i] = level - 1;
        q_maxl[i] = (V->flags & PROJ_LEVEL_PLUS1) ? level + 1 : level;
        if (q_level) q_level[i] = level;
        q_valid[i] = 1;
    }
}

/* The candidate loop of Fuse (:892-938, :1052-1078) and SearchBySim3 (:1193-1224, :1273-1304): best Hamming distance in the window,
 * no claims.  win_origin2 (may be NULL): KeyFrame::GetFeaturesInArea subtracts the KeyFrame's integer mnMinX / mnMinY while the grid
 * itself was filled by the Frame with its float bounds.  inv_level_sigma2 (may be NULL): the monocular chi-square gate of Fuse. */
void oracle_search_best_in_window(const oracle_grid_params *g, const float *win_origin2, int N, const float *f_xy, const int *f_octave,
                                  const uint8_t *f_desc, int M, const uint8_t *q_valid, const float *q_uv, const float *q_radius,
                                  const int *q_minl, const int *q_maxl, const uint8_t *q_desc, int th_dist,
                                  const float *inv_level_sigma2, double chi2_gate, int *q_best_idx, int *q_best_dist)
{
    const int nc = GRID_COLS * GRID_ROWS;
    int *cell_start = (int *)malloc(sizeof(int) * (nc + 1));
    int *cell_items = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    int *cands = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    oracle_grid_build(g, N, f_xy, cell_start, cell_items);
    oracle_grid_params gw = *g;
    if (win_origin2) { gw.min_x = win_origin2[0]; gw.min_y = win_origin2[1]; }
    for (int i = 0; i < M; i++) {
        q_best_idx[i] = -1; q_best_dist[i] = -1;
        if (!q_valid[i]) continue;
        const float u = q_uv[2 * i], v = q_uv[2 * i + 1];
        const int n = oracle_features_in_area(&gw, cell_start, cell_items, f_xy, f_octave, u, v, q_radius[i], q_minl[i], q_maxl[i], cands);
        int best = 0x7fffffff, best_idx = -1;
        for (int c = 0; c < n; c++) {
            const int k = cands[c];
            if (inv_level_sigma2) {
                const float ex = u - f_xy[2 * k], ey = v - f_xy[2 * k + 1];
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_level_sigma2[f_octave[k]] > chi2_gate) continue;     /* float product against the double 5.99 (:927) */
            }
            const int d = oracle_descriptor_distance(q_desc + 32 * (size_t)i, f_desc + 32 * (size_t)k);
            if (d < best) { best = d; best_idx = k; }
        }
        if (best_idx >= 0) { q_best_dist[i] = best; if (best <= th_dist) q_best_idx[i] = best_idx; }
    }
    free(cell_start); free(cell_items); free(cands);
}

/* ======================================================================================== */
/* SE3 helpers (g2o::SE3Quat on Eigen::Quaterniond), fp64                                    */
typedef struct { double q[4]; /* x y z w */ double t[3]; } se3;

static void quat_from_R(const double R[9], double q[4])    /* Eigen::Quaterniond(Matrix3d) */
{
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}
static void quat_normalize_pos(double q[4])                /* SE3Quat::normalizeRotation, se3quat.h:280-285 */
{
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat_rotate(const double q[4], const double v[3], double out[3])   /* Eigen _transformVector */
{
    double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
    out[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
    out[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}
static void quat_mul(const double a[4], const double b[4], double o[4])
{
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
static void quat_to_R(const double q[4], double R[9])      /* Eigen toRotationMatrix */
{
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static void se3_from_Tcw_f32(const float *T, se3 *s)       /* Converter::toSE3Quat, Converter.cc:37-47 */
{
    double R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = T[4 * r + c];
    quat_from_R(R, s->q);
    quat_normalize_pos(s->q);
    s->t[0] = T[3]; s->t[1] = T[7]; s->t[2] = T[11];
}
static void se3_to_Tcw_f32(const se3 *s, float *T)         /* Converter::toCvMat(SE3Quat), Converter.cc:49-72 */
{
    double R[9];
    quat_to_R(s->q, R);
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T[4 * r + c] = (float)R[3 * r + c]; T[4 * r + 3] = (float)s->t[r]; }
    T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}
static void se3_map(const se3 *s, const double X[3], double out[3])
{
    quat_rotate(s->q, X, out);
    out[0] += s->t[0]; out[1] += s->t[1]; out[2] += s->t[2];
}
static void mat3_mul(const double A[9], const double B[9], double C[9])
{
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
/* SE3Quat::exp(update), update = (omega, upsilon); se3quat.h:223-257 */
static void se3_exp(const double u[6], se3 *out)
{
    const double w[3] = {u[0], u[1], u[2]}, ups[3] = {u[3], u[4], u[5]};
    const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9], R[9], V[9];
    mat3_mul(O, O, O2);
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3);
        for (int i = 0; i < 9; i++) {
            R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
            V[i] = (i % 4 == 0 ? 1.0 : 0.0) + b * O[i] + c * O2[i];
        }
    }
    quat_from_R(R, out->q);
    for (int r = 0; r < 3; r++) out->t[r] = V[3 * r] * ups[0] + V[3 * r + 1] * ups[1] + V[3 * r + 2] * ups[2];
    quat_normalize_pos(out->q);            /* SE3Quat(q, t) ctor normalises */
}
static void se3_mul(const se3 *a, const se3 *b, se3 *o)    /* SE3Quat::operator*, se3quat.h:104-110 */
{
    double rt[3];
    quat_rotate(a->q, b->t, rt);
    se3 r;
    r.t[0] = a->t[0] + rt[0]; r.t[1] = a->t[1] + rt[1]; r.t[2] = a->t[2] + rt[2];
    quat_mul(a->q, b->q, r.q);
    quat_normalize_pos(r.q);
    *o = r;
}

/* ======================================================================================== */
/* Bundle adjustment problem over flat arrays                                                 */
typedef struct {
    int K, P, E;
    se3 *pose; uint8_t *pose_fixed; const double *intr;     /* intr[K*4] fx fy cx cy */
    double *pt;                                              /* [P*3]; pt_fixed => unary (pose-only) edges */
    uint8_t pt_fixed_all;
    const int *e_kf, *e_pt; const double *e_obs;            /* [E], [E], [E*2] */
    const double *e_w;                                       /* information = w * I2 */
    int *e_level; uint8_t *e_robust; double *e_err;          /* stored _error [E*2] */
    double delta, dsqr;
    /* active set / index mapping */
    int *pose_idx, *pt_idx;                                  /* hessian index or -1 */
    int nA, nL;                                              /* active free poses / landmarks */
    int *act_edges; int nAE;
    int *map_pose, *map_pt;                                  /* index -> vertex */
    /* system */
    double *Hpp /*[nA*36]*/, *Hll /*[nL*9]*/, *bp /*[nA*6]*/, *bl /*[nL*3]*/;
    double *Hpl /*[E*18] per active edge slot*/;
    double *x;                                               /* [6nA + 3nL] */
    double *S, *bs;                                          /* reduced system (dense, 6nA) */
    int *first;                                              /* profile */
    double lambda, ni; int nbad;
    const volatile int *stop;
    int lm_iterations_done, lm_trials_done;
} ba_t;

static void ba_compute_error(ba_t *B, int e)
{
    /* EdgeSE3ProjectXYZ::computeError / OnlyPose, types_six_dof_expmap.h:90-95,153-157 */
    double Xc[3];
    se3_map(&B->pose[B->e_kf[e]], &B->pt[3 * B->e_pt[e]], Xc);
    const double *in = &B->intr[4 * B->e_kf[e]];
    const double px = Xc[0] / Xc[2], py = Xc[1] / Xc[2];       /* project2d */
    B->e_err[2 * e] = B->e_obs[2 * e] - (px * in[0] + in[2]);
    B->e_err[2 * e + 1] = B->e_obs[2 * e + 1] - (py * in[1] + in[3]);
}
static double ba_chi2(const ba_t *B, int e)                   /* BaseEdge::chi2, base_edge.h:58-61 */
{
    const double e0 = B->e_err[2 * e], e1 = B->e_err[2 * e + 1], w = B->e_w[e];
    return e0 * (w * e0 + 0.0 * e1) + e1 * (0.0 * e0 + w * e1);
}
static void huber(const ba_t *B, double e, double rho[3])     /* robust_kernel_impl.cpp:78-91 */
{
    if (e <= B->dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
    else { const double s = sqrt(e); rho[0] = 2 * s * B->delta - B->dsqr; rho[1] = B->delta / s; rho[2] = -0.5 * rho[1] / e; }
}
static int ba_depth_positive(const ba_t *B, int e)
{
    double Xc[3];
    se3_map(&B->pose[B->e_kf[e]], &B->pt[3 * B->e_pt[e]], Xc);
    return Xc[2] > 0.0;
}

/* initializeOptimization(level 0) + buildIndexMapping, sparse_optimizer.cpp:166-267 */
static void ba_init_active(ba_t *B)
{
    B->nAE = 0;
    uint8_t *pose_act = (uint8_t *)calloc(B->K, 1), *pt_act = (uint8_t *)calloc(B->P > 0 ? B->P : 1, 1);
    for (int e = 0; e < B->E; e++) {
        if (B->e_level[e] != 0) continue;
        if (B->pt_fixed_all && B->pose_fixed[B->e_kf[e]]) continue;      /* allVerticesFixed */
        B->act_edges[B->nAE++] = e;
        pose_act[B->e_kf[e]] = 1; pt_act[B->e_pt[e]] = 1;
    }
    B->nA = 0; B->nL = 0;
    for (int k = 0; k < B->K; k++) {
        B->pose_idx[k] = -1;
        if (pose_act[k] && !B->pose_fixed[k]) { B->map_pose[B->nA] = k; B->pose_idx[k] = B->nA++; }
    }
    for (int p = 0; p < B->P; p++) {
        B->pt_idx[p] = -1;
        if (!B->pt_fixed_all && pt_act[p]) { B->map_pt[B->nL] = p; B->pt_idx[p] = B->nL++; }
    }
    free(pose_act); free(pt_act);
}

static void ba_compute_active_errors(ba_t *B) { for (int i = 0; i < B->nAE; i++) ba_compute_error(B, B->act_edges[i]); }
static double ba_active_robust_chi2(const ba_t *B)            /* sparse_optimizer.cpp:100-114 */
{
    double chi = 0.0, rho[3];
    for (int i = 0; i < B->nAE; i++) {
        const int e = B->act_edges[i];
        if (B->e_robust[e]) { huber(B, ba_chi2(B, e), rho); chi += rho[0]; }
        else chi += ba_chi2(B, e);
    }
    return chi;
}

/* linearizeOplus of one edge: Jp (2x6, pose) and Jl (2x3, point; untouched for pose-only edges) */
static void ba_edge_jacobians(const ba_t *B, int e, double Jp[12], double Jl[6])
{
        const int kf = B->e_kf[e], p = B->e_pt[e];
        const double *in = &B->intr[4 * kf];
        const double fx = in[0], fy = in[1];
        double Xc[3];
        se3_map(&B->pose[kf], &B->pt[3 * p], Xc);
        const double x = Xc[0], y = Xc[1], z = Xc[2];
        if (B->pt_fixed_all) {                      /* EdgeSE3ProjectXYZOnlyPose::linearizeOplus, .cpp:266-288 */
            const double invz = 1.0 / z, invz_2 = invz * invz;
            Jp[0] = x * y * invz_2 * fx; Jp[1] = -(1 + (x * x * invz_2)) * fx; Jp[2] = y * invz * fx;
            Jp[3] = -invz * fx; Jp[4] = 0; Jp[5] = x * invz_2 * fx;
            Jp[6] = (1 + y * y * invz_2) * fy; Jp[7] = -x * y * invz_2 * fy; Jp[8] = -x * invz * fy;
            Jp[9] = 0; Jp[10] = -invz * fy; Jp[11] = y * invz_2 * fy;
        } else {                                    /* EdgeSE3ProjectXYZ::linearizeOplus, .cpp:103-139 */
            const double z_2 = z * z;
            double R[9];
            quat_to_R(B->pose[kf].q, R);
            const double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
            const double s = -1. / z;
            for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) {
                const double st0 = s * tmp[3 * r], st1 = s * tmp[3 * r + 1], st2 = s * tmp[3 * r + 2];
                Jl[3 * r + c] = st0 * R[c] + st1 * R[3 + c] + st2 * R[6 + c];
            }
            Jp[0] = x * y / z_2 * fx; Jp[1] = -(1 + (x * x / z_2)) * fx; Jp[2] = y / z * fx;
            Jp[3] = -1. / z * fx; Jp[4] = 0; Jp[5] = x / z_2 * fx;
            Jp[6] = (1 + y * y / z_2) * fy; Jp[7] = -x * y / z_2 * fy; Jp[8] = -x / z * fy;
            Jp[9] = 0; Jp[10] = -1. / z * fy; Jp[11] = y / z_2 * fy;
        }
}

/* buildSystem: linearizeOplus + constructQuadraticForm per active edge, block_solver.hpp:506-564 */
static void ba_build_system(ba_t *B)
{
    memset(B->Hpp, 0, sizeof(double) * 36 * (B->nA > 0 ? B->nA : 1));
    memset(B->bp, 0, sizeof(double) * 6 * (B->nA > 0 ? B->nA : 1));
    if (B->nL) { memset(B->Hll, 0, sizeof(double) * 9 * B->nL); memset(B->bl, 0, sizeof(double) * 3 * B->nL); }
    for (int i = 0; i < B->nAE; i++) {
        const int e = B->act_edges[i];
        const int kf = B->e_kf[e], p = B->e_pt[e];
        const int ip = B->pose_idx[kf], il = B->pt_fixed_all ? -1 : B->pt_idx[p];
        double Jp[12], Jl[6];   /* 2x6 pose, 2x3 point */
        ba_edge_jacobians(B, e, Jp, Jl);
        /* constructQuadraticForm, base_binary_edge.hpp:55-120 / base_unary_edge.hpp:43-72 */
        double w = B->e_w[e], rw = 1.0;
        if (B->e_robust[e]) { double rho[3]; huber(B, ba_chi2(B, e), rho); rw = rho[1]; }
        const double wo = rw * w;                              /* weightedOmega = rho' * omega */
        const double r0 = -w * B->e_err[2 * e] * rw, r1 = -w * B->e_err[2 * e + 1] * rw;   /* omega_r */
        if (ip >= 0) {
            double *H = &B->Hpp[36 * ip], *b = &B->bp[6 * ip];
            for (int a = 0; a < 6; a++) {
                b[a] += Jp[a] * r0 + Jp[6 + a] * r1;
                for (int c = 0; c < 6; c++) H[6 * a + c] += Jp[a] * wo * Jp[c] + Jp[6 + a] * wo * Jp[6 + c];
            }
        }
        if (il >= 0) {
            double *H = &B->Hll[9 * il], *b = &B->bl[3 * il];
            for (int a = 0; a < 3; a++) {
                b[a] += Jl[a] * r0 + Jl[3 + a] * r1;
                for (int c = 0; c < 3; c++) H[3 * a + c] += Jl[a] * wo * Jl[c] + Jl[3 + a] * wo * Jl[3 + c];
            }
            double *W = &B->Hpl[18 * i];                       /* 6x3: Jp^T wo Jl */
            if (ip >= 0)
                for (int a = 0; a < 6; a++) for (int c = 0; c < 3; c++) W[3 * a + c] = Jp[a] * wo * Jl[c] + Jp[6 + a] * wo * Jl[3 + c];
        }
    }
}

static void inv3(const double *A, double *I)                  /* Matrix3d::inverse(): cofactors / determinant */
{
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02, id = 1.0 / det;
    I[0] = c00 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = c01 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = c02 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

/* profile LDL^T of the dense symmetric n x n matrix S (row-major, lower part used), in place.
 * returns 0 on a zero pivot (SimplicialLDLT's NumericalIssue), or on a negative one when require_positive
 * (LinearSolverDense's isPositive() gate, linear_solver_dense.h:107-112). */
static int ldlt_solve(double *S, int n, const int *first, const double *b, double *x, int require_positive)
{
    for (int i = 0; i < n; i++) {
        for (int j = first[i]; j <= i; j++) {
            double s = S[(size_t)i * n + j];
            const int k0 = first[i] > first[j] ? first[i] : first[j];
            for (int k = k0; k < j; k++) s -= S[(size_t)i * n + k] * S[(size_t)j * n + k] * S[(size_t)k * n + k];
            if (j < i) S[(size_t)i * n + j] = s / S[(size_t)j * n + j];
            else { if (s == 0.0 || !(s == s) || (require_positive && s < 0.0)) return 0; S[(size_t)i * n + i] = s; }
        }
    }
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = first[i]; k < i; k++) s -= S[(size_t)i * n + k] * x[k]; x[i] = s; }
    for (int i = 0; i < n; i++) x[i] /= S[(size_t)i * n + i];
    for (int i = n - 1; i >= 0; i--) { const double xi = x[i]; for (int k = first[i]; k < i; k++) x[k] -= S[(size_t)i * n + k] * xi; }
    return 1;
}

/* setLambda + solve (Schur) + restoreDiagonal are folded: lambda is added on the fly. block_solver.hpp:358-490,568-608 */
static int ba_solve(ba_t *B)
{
    const int nA = B->nA, nL = B->nL, n = 6 * nA;
    const double lam = B->lambda;
    double *S = B->S, *bs = B->bs;
    memset(S, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < nA; i++)
        for (int a = 0; a < 6; a++) {
            for (int c = 0; c < 6; c++) S[(size_t)(6 * i + a) * n + 6 * i + c] = B->Hpp[36 * i + 6 * a + c];
            S[(size_t)(6 * i + a) * n + 6 * i + a] += lam;
            bs[6 * i + a] = B->bp[6 * i + a];
        }
    for (int i = 0; i < n; i++) B->first[i] = (i / 6) * 6;
    double *Dinv = NULL, *dbl = NULL;
    int *lm_edges_start = NULL, *lm_edges = NULL;
    if (nL) {
        Dinv = (double *)malloc(sizeof(double) * 9 * nL);
        /* per-landmark list of active-edge slots (sorted by pose index like the CCS column) */
        lm_edges_start = (int *)calloc(nL + 1, sizeof(int));
        for (int i = 0; i < B->nAE; i++) { const int e = B->act_edges[i]; if (B->pose_idx[B->e_kf[e]] >= 0) lm_edges_start[B->pt_idx[B->e_pt[e]] + 1]++; }
        for (int l = 0; l < nL; l++) lm_edges_start[l + 1] += lm_edges_start[l];
        lm_edges = (int *)malloc(sizeof(int) * (lm_edges_start[nL] > 0 ? lm_edges_start[nL] : 1));
        int *fill = (int *)calloc(nL, sizeof(int));
        for (int i = 0; i < B->nAE; i++) {
            const int e = B->act_edges[i];
            if (B->pose_idx[B->e_kf[e]] < 0) continue;
            const int l = B->pt_idx[B->e_pt[e]];
            lm_edges[lm_edges_start[l] + fill[l]++] = i;
        }
        free(fill);
        for (int l = 0; l < nL; l++) {             /* sort slots of a landmark by pose index (insertion) */
            int *a = lm_edges + lm_edges_start[l]; const int m = lm_edges_start[l + 1] - lm_edges_start[l];
            for (int u = 1; u < m; u++) { int v = a[u], w = u; while (w > 0 && B->pose_idx[B->e_kf[B->act_edges[a[w - 1]]]] > B->pose_idx[B->e_kf[B->act_edges[v]]]) { a[w] = a[w - 1]; w--; } a[w] = v; }
        }
        for (int l = 0; l < nL; l++) {
            double D[9];
            memcpy(D, &B->Hll[9 * l], sizeof D);
            D[0] += lam; D[4] += lam; D[8] += lam;
            double *Di = &Dinv[9 * l];
            inv3(D, Di);
            const double *bl = &B->bl[3 * l];
            const double db[3] = {Di[0] * bl[0] + Di[1] * bl[1] + Di[2] * bl[2], Di[3] * bl[0] + Di[4] * bl[1] + Di[5] * bl[2],
                                  Di[6] * bl[0] + Di[7] * bl[1] + Di[8] * bl[2]};
            for (int u = lm_edges_start[l]; u < lm_edges_start[l + 1]; u++) {
                const int s1 = lm_edges[u], i1 = B->pose_idx[B->e_kf[B->act_edges[s1]]];
                const double *W1 = &B->Hpl[18 * s1];
                double BD[18];
                for (int a = 0; a < 6; a++) for (int c = 0; c < 3; c++) BD[3 * a + c] = W1[3 * a] * Di[c] + W1[3 * a + 1] * Di[3 + c] + W1[3 * a + 2] * Di[6 + c];
                for (int a = 0; a < 6; a++) bs[6 * i1 + a] -= W1[3 * a] * db[0] + W1[3 * a + 1] * db[1] + W1[3 * a + 2] * db[2];
                for (int v = u; v < lm_edges_start[l + 1]; v++) {
                    const int s2 = lm_edges[v], i2 = B->pose_idx[B->e_kf[B->act_edges[s2]]];
                    const double *W2 = &B->Hpl[18 * s2];
                    /* upper block (i1, i2) -= BD * W2^T ; stored into lower part (i2 rows, i1 cols) transposed */
                    for (int a = 0; a < 6; a++) for (int c = 0; c < 6; c++) {
                        const double val = BD[3 * a] * W2[3 * c] + BD[3 * a + 1] * W2[3 * c + 1] + BD[3 * a + 2] * W2[3 * c + 2];
                        if (i1 == i2) { S[(size_t)(6 * i1 + a) * n + 6 * i1 + c] -= val; }
                        else S[(size_t)(6 * i2 + c) * n + 6 * i1 + a] -= val;
                    }
                    if (i2 != i1) for (int c = 0; c < 6; c++) if (B->first[6 * i2 + c] > 6 * i1) B->first[6 * i2 + c] = 6 * i1;
                }
            }
        }
    }
    double *xp = B->x;
    /* pose-only problems use LinearSolverDense (Eigen LDLT + isPositive gate), BA uses LinearSolverEigen (zero pivot only) */
    const int ok = n > 0 ? ldlt_solve(S, n, B->first, bs, xp, B->pt_fixed_all) : 1;
    if (ok && nL) {
        /* x_l = Dinv (b_l - Hpl^T x_p), block_solver.hpp:463-487 */
        dbl = (double *)malloc(sizeof(double) * 3 * nL);
        memcpy(dbl, B->bl, sizeof(double) * 3 * nL);
        for (int l = 0; l < nL; l++)
            for (int u = lm_edges_start[l]; u < lm_edges_start[l + 1]; u++) {
                const int s1 = lm_edges[u], i1 = B->pose_idx[B->e_kf[B->act_edges[s1]]];
                const double *W1 = &B->Hpl[18 * s1];
                for (int c = 0; c < 3; c++) for (int a = 0; a < 6; a++) dbl[3 * l + c] -= W1[3 * a + c] * xp[6 * i1 + a];
            }
        for (int l = 0; l < nL; l++) {
            const double *Di = &Dinv[9 * l], *c = &dbl[3 * l];
            double *xl = &B->x[n + 3 * l];
            for (int a = 0; a < 3; a++) xl[a] = Di[3 * a] * c[0] + Di[3 * a + 1] * c[1] + Di[3 * a + 2] * c[2];
        }
    }
    free(Dinv); free(dbl); free(lm_edges_start); free(lm_edges);
    return ok;
}

static void ba_update(ba_t *B)                                /* SparseOptimizer::update + oplusImpl */
{
    for (int i = 0; i < B->nA; i++) {
        se3 d, r;
        se3_exp(&B->x[6 * i], &d);
        se3_mul(&d, &B->pose[B->map_pose[i]], &r);
        B->pose[B->map_pose[i]] = r;
    }
    for (int l = 0; l < B->nL; l++) { double *p = &B->pt[3 * B->map_pt[l]]; const double *x = &B->x[6 * B->nA + 3 * l]; p[0] += x[0]; p[1] += x[1]; p[2] += x[2]; }
}

static int ba_terminate(const ba_t *B) { return B->stop && *B->stop; }

enum { LM_OK = 0, LM_TERMINATE = 1 };
/* OptimizationAlgorithmLevenberg::solve, optimization_algorithm_levenberg.cpp:61-164 */
static int ba_lm_iteration(ba_t *B, int iteration)
{
    ba_compute_active_errors(B);
    double currentChi = ba_active_robust_chi2(B), tempChi = currentChi;
    const double iniChi = currentChi;
    ba_build_system(B);
    if (iteration == 0) {                                     /* computeLambdaInit, :166-180 */
        double maxDiag = 0.;
        for (int i = 0; i < B->nA; i++) for (int j = 0; j < 6; j++) maxDiag = fmax(fabs(B->Hpp[36 * i + 7 * j]), maxDiag);
        for (int l = 0; l < B->nL; l++) for (int j = 0; j < 3; j++) maxDiag = fmax(fabs(B->Hll[9 * l + 4 * j]), maxDiag);
        B->lambda = 1e-5 * maxDiag; B->ni = 2; B->nbad = 0;
    }
    double rho = 0;
    int qmax = 0;
    const int nx = 6 * B->nA + 3 * B->nL;
    se3 *pose_bak = (se3 *)malloc(sizeof(se3) * (B->nA > 0 ? B->nA : 1));
    double *pt_bak = (double *)malloc(sizeof(double) * 3 * (B->nL > 0 ? B->nL : 1));
    do {
        for (int i = 0; i < B->nA; i++) pose_bak[i] = B->pose[B->map_pose[i]];            /* push */
        for (int l = 0; l < B->nL; l++) memcpy(&pt_bak[3 * l], &B->pt[3 * B->map_pt[l]], 3 * sizeof(double));
        const int ok2 = ba_solve(B);
        ba_update(B);
        ba_compute_active_errors(B);
        tempChi = ba_active_robust_chi2(B);
        if (!ok2) tempChi = DBL_MAX;
        rho = currentChi - tempChi;
        double scale = 0.;                                   /* computeScale, :182-189 */
        for (int j = 0; j < nx; j++) {
            const double bj = j < 6 * B->nA ? B->bp[j] : B->bl[j - 6 * B->nA];
            scale += B->x[j] * (B->lambda * B->x[j] + bj);
        }
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && isfinite(tempChi)) {
            double alpha = 1. - pow((2 * rho - 1), 3);
            alpha = fmin(alpha, 2. / 3.);
            const double sf = fmax(1. / 3., alpha);
            B->lambda *= sf; B->ni = 2; currentChi = tempChi;
        } else {
            B->lambda *= B->ni; B->ni *= 2;
            for (int i = 0; i < B->nA; i++) B->pose[B->map_pose[i]] = pose_bak[i];        /* pop */
            for (int l = 0; l < B->nL; l++) memcpy(&B->pt[3 * B->map_pt[l]], &pt_bak[3 * l], 3 * sizeof(double));
        }
        qmax++;
        B->lm_trials_done++;
    } while (rho < 0 && qmax < 10 && !ba_terminate(B));
    free(pose_bak); free(pt_bak);
    B->lm_iterations_done++;
    if (qmax == 10 || rho == 0) return LM_TERMINATE;
    if ((iniChi - currentChi) * 1e3 < iniChi) B->nbad++; else B->nbad = 0;
    if (B->nbad >= 3) return LM_TERMINATE;
    return LM_OK;
}

/* SparseOptimizer::optimize, sparse_optimizer.cpp:354-419 */
static int ba_optimize(ba_t *B, int iterations)
{
    if (B->nA + B->nL == 0) return -1;
    int done = 0;
    for (int i = 0; i < iterations && !ba_terminate(B); i++) {
        const int res = ba_lm_iteration(B, i);
        done++;
        if (res != LM_OK) break;
    }
    return done;
}

static void ba_alloc(ba_t *B)
{
    const int K = B->K, P = B->P > 0 ? B->P : 1, E = B->E > 0 ? B->E : 1;
    B->pose_idx = (int *)malloc(sizeof(int) * K); B->pt_idx = (int *)malloc(sizeof(int) * P);
    B->map_pose = (int *)malloc(sizeof(int) * K); B->map_pt = (int *)malloc(sizeof(int) * P);
    B->act_edges = (int *)malloc(sizeof(int) * E);
    B->Hpp = (double *)malloc(sizeof(double) * 36 * K); B->bp = (double *)malloc(sizeof(double) * 6 * K);
    B->Hll = (double *)malloc(sizeof(double) * 9 * P); B->bl = (double *)malloc(sizeof(double) * 3 * P);
    B->Hpl = (double *)malloc(sizeof(double) * 18 * E);
    B->x = (double *)calloc((size_t)6 * K + 3 * P, sizeof(double));
    B->S = (double *)malloc(sizeof(double) * 36 * (size_t)K * K); B->bs = (double *)malloc(sizeof(double) * 6 * K);
    B->first = (int *)malloc(sizeof(int) * 6 * K);
    B->lm_iterations_done = B->lm_trials_done = 0;
}
static void ba_free(ba_t *B)
{
    free(B->pose_idx); free(B->pt_idx); free(B->map_pose); free(B->map_pt); free(B->act_edges);
    free(B->Hpp); free(B->bp); free(B->Hll); free(B->bl); free(B->Hpl); free(B->x); free(B->S); free(B->bs); free(B->first);
}

/* ---------------------------------------------------------------------------------------- */
/* Optimizer::PoseOptimization, Optimizer.cc:262-474 (monocular edges).
 * Tcw: in/out row-major 4x4 f32.  Xw f32[M,3], obs f32[M,2] (kpUn.pt), inv_sigma2 f32[M], K4 fx fy cx cy (float).
 * outlier u8[M] out.  Returns nInitialCorrespondences - nBad (0 if < 3 correspondences). */
int oracle_pose_optimization(float *Tcw, int M, const float *Xw, const float *obs, const float *inv_sigma2,
                             const float *K4, uint8_t *outlier)
{
    if (M < 3) return 0;                                      /* :387-388 */
    ba_t B; memset(&B, 0, sizeof B);
    B.K = 1; B.P = M; B.E = M; B.pt_fixed_all = 1;
    se3 pose; uint8_t fixed = 0;
    double intr[4] = {K4[0], K4[1], K4[2], K4[3]};
    B.pose = &pose; B.pose_fixed = &fixed; B.intr = intr;
    double *pt = (double *)malloc(sizeof(double) * 3 * M), *o = (double *)malloc(sizeof(double) * 2 * M), *w = (double *)malloc(sizeof(double) * M);
    int *ekf = (int *)calloc(M, sizeof(int)), *ept = (int *)malloc(sizeof(int) * M), *lvl = (int *)calloc(M, sizeof(int));
    uint8_t *rob = (uint8_t *)malloc(M); double *err = (double *)calloc(2 * M, sizeof(double));
    for (int i = 0; i < M; i++) {
        for (int c = 0; c < 3; c++) pt[3 * i + c] = Xw[3 * i + c];
        o[2 * i] = obs[2 * i]; o[2 * i + 1] = obs[2 * i + 1];
        w[i] = inv_sigma2[i]; ept[i] = i; rob[i] = 1; outlier[i] = 0;
    }
    B.pt = pt; B.e_kf = ekf; B.e_pt = ept; B.e_obs = o; B.e_w = w; B.e_level = lvl; B.e_robust = rob; B.e_err = err;
    const float deltaMono = sqrtf(5.991f);                    /* const float deltaMono = sqrt(5.991) */
    B.delta = (double)(float)sqrt(5.991); (void)deltaMono;
    B.dsqr = (double)(float)(B.delta * B.delta);          /* RobustKernelHuber keeps delta^2 in a FLOAT member (robust_kernel_impl.h:84, set by setDelta, robust_kernel_impl.cpp:65-69) */
    ba_alloc(&B);
    const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f};
    int nBad = 0;
    for (int it = 0; it < 4; it++) {
        se3_from_Tcw_f32(Tcw, &pose);                         /* reset to the initial pose every round, :400 */
        ba_init_active(&B);
        ba_optimize(&B, 10);
        nBad = 0;
        for (int i = 0; i < M; i++) {
            if (outlier[i]) ba_compute_error(&B, i);
            const float chi2 = (float)ba_chi2(&B, i);
            if (chi2 > chi2Mono[it]) { outlier[i] = 1; lvl[i] = 1; nBad++; }
            else { outlier[i] = 0; lvl[i] = 0; }
            if (it == 2) rob[i] = 0;
        }
        if (M < 10) break;                                    /* optimizer.edges().size()<10, :463 */
    }
    se3_to_Tcw_f32(&pose, Tcw);
    ba_free(&B);
    free(pt); free(o); free(w); free(ekf); free(ept); free(lvl); free(rob); free(err);
    return M - nBad;
}

/* ---------------------------------------------------------------------------------------- */
/* Bundle adjustment over flat arrays: LocalBundleAdjustment (Optimizer.cc:476-801) when two_stage != 0
 * (its0 robust iterations, chi2/depth gating, its1 non-robust iterations) and BundleAdjustment (:68-260) when
 * two_stage == 0 (its0 iterations, robust flag).  poses [K,16] f32 in/out, fixed u8[K], intr f64[K,4],
 * points [P,3] f32 in/out, edges kf/pt i32[E], uv f32[E,2], inv_sigma2 f32[E].
 * e_chi2 f64[E] / e_depth_ok u8[E] / e_outlier u8[E] out: the reference's final check (:734-766).
 * stats (optional, 2 ints): LM iterations and trials executed. */
int oracle_bundle_adjust(int K, float *poses, const uint8_t *fixed, const double *intr, int P, float *points,
                         int E, const int *e_kf, const int *e_pt, const float *e_uv, const float *e_inv_sigma2,
                         int two_stage, int its0, int its1, int robust, const volatile int *stop,
                         double *e_chi2, uint8_t *e_depth_ok, uint8_t *e_outlier, int *stats)
{
    ba_t B; memset(&B, 0, sizeof B);
    B.K = K; B.P = P; B.E = E; B.pt_fixed_all = 0; B.stop = stop;
    se3 *pose = (se3 *)malloc(sizeof(se3) * K);
    uint8_t *fx = (uint8_t *)malloc(K);
    for (int k = 0; k < K; k++) { se3_from_Tcw_f32(poses + 16 * k, &pose[k]); fx[k] = fixed[k]; }
    double *pt = (double *)malloc(sizeof(double) * 3 * (P > 0 ? P : 1));
    for (int i = 0; i < 3 * P; i++) pt[i] = points[i];
    double *o = (double *)malloc(sizeof(double) * 2 * (E > 0 ? E : 1)), *w = (double *)malloc(sizeof(double) * (E > 0 ? E : 1));
    int *lvl = (int *)calloc(E > 0 ? E : 1, sizeof(int));
    uint8_t *rob = (uint8_t *)malloc(E > 0 ? E : 1); double *err = (double *)calloc(2 * (E > 0 ? E : 1), sizeof(double));
    for (int e = 0; e < E; e++) { o[2 * e] = e_uv[2 * e]; o[2 * e + 1] = e_uv[2 * e + 1]; w[e] = e_inv_sigma2[e]; rob[e] = (uint8_t)(two_stage ? 1 : robust); }
    B.pose = pose; B.pose_fixed = fx; B.intr = intr; B.pt = pt; B.e_kf = e_kf; B.e_pt = e_pt; B.e_obs = o; B.e_w = w;
    B.e_level = lvl; B.e_robust = rob; B.e_err = err;
    /* LocalBundleAdjustment: const float thHuberMono = sqrt(5.991) (Optimizer.cc:592); BundleAdjustment: const float thHuber2D = sqrt(5.99) (:106) */
    B.delta = two_stage ? (double)(float)sqrt(5.991) : (double)(float)sqrt(5.99);
    B.dsqr = (double)(float)(B.delta * B.delta);   /* RobustKernelHuber keeps delta^2 in a FLOAT member (robust_kernel_impl.h:84, set by setDelta, robust_kernel_impl.cpp:65-69) */
    ba_alloc(&B);
    int do_more = 1;
    if (stop && *stop) {                                      /* :678-680: return before touching anything */
        ba_free(&B);
        free(pose); free(fx); free(pt); free(o); free(w); free(lvl); free(rob); free(err);
        return 1;
    }
    {
        ba_init_active(&B);
        ba_optimize(&B, its0);
        if (two_stage) {
            if (stop && *stop) do_more = 0;
            if (do_more) {
                for (int e = 0; e < E; e++) {                 /* :691-705 */
                    if (ba_chi2(&B, e) > 5.991 || !ba_depth_positive(&B, e)) lvl[e] = 1;
                    rob[e] = 0;
                }
                ba_init_active(&B);
                ba_optimize(&B, its1);
            }
        }
    }
    for (int e = 0; e < E; e++) {                             /* :734-766 */
        const double c = ba_chi2(&B, e); const int d = ba_depth_positive(&B, e);
        if (e_chi2) e_chi2[e] = c;
        if (e_depth_ok) e_depth_ok[e] = (uint8_t)d;
        if (e_outlier) e_outlier[e] = (uint8_t)(c > 5.991 || !d);
    }
    /* every local KF (free, or fixed because mnId==0: fixed[k]==1) is written back through the SE3Quat round trip,
     * :785-791; fixed cameras outside the local window (fixed[k]==2) are never touched */
    for (int k = 0; k < K; k++) if (fixed[k] != 2) se3_to_Tcw_f32(&pose[k], poses + 16 * k);
    for (int i = 0; i < 3 * P; i++) points[i] = (float)pt[i];
    if (stats) { stats[0] = B.lm_iterations_done; stats[1] = B.lm_trials_done; }
    ba_free(&B);
    free(pose); free(fx); free(pt); free(o); free(w); free(lvl); free(rob); free(err);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Frame::UndistortKeyPoints (Frame.cc:404-434) = cv::undistortPoints(pts, K, distCoef, R = I, P = K) with the default
 * termination criteria (5 iterations, no epsilon test).  OpenCV is un-vendored; restated from cvUndistortPointsInternal and
 * pinned against cv2.undistortPoints (tests/test_oracle_opencv_pin.py).  dist5 = k1, k2, p1, p2, k3 (k3 = 0 for the 4-entry
 * DistCoef of the reference's YAML files).  When k1 == 0 the reference copies the keypoints (Frame.cc:406-410). */
void oracle_undistort_points(const float *K4, const float *dist5, int n, const float *xy, float *xy_un)
{
    if (dist5[0] == 0.0f) { memcpy(xy_un, xy, sizeof(float) * 2 * (size_t)n); return; }
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k1 = dist5[0], k2 = dist5[1], p1 = dist5[2], p2 = dist5[3], k3 = dist5[4];
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx; y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
            const double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0. * r2 + 0. * r2 * r2;
            x = (x0 - dX) * icdist; y = (y0 - dY) * icdist;
        }
        const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        xy_un[2 * i] = (float)(xx * ww); xy_un[2 * i + 1] = (float)(yy * ww);
    }
}

/* Frame::isInFrustum (Frame.cc:269-325) with MapPoint::PredictScale (MapPoint.cc:385-394) and
 * Get{Min,Max}DistanceInvariance (MapPoint.cc:373-383) over flat arrays.  Ow = camera centre (Frame::mOw).
 * bounds4 = mnMinX, mnMinY, mnMaxX, mnMaxY.  cv::Mat algebra: R*P+t as the small-gemm path (fp32, left to right);
 * cv::norm and Mat::dot accumulate in double.  Outputs are written only for points in view (mbTrackInView). */
void oracle_is_in_frustum(const float *Tcw, const float *Ow, const float *K4, const float *bounds4, float log_scale_factor, float cos_limit,
                          int M, const float *Xw, const float *normal, const float *mf_min_dist, const float *mf_max_dist,
                          uint8_t *in_view, float *proj_xy, int *pred_level, float *view_cos)
{
    for (int i = 0; i < M; i++) {
        in_view[i] = 0;
        const float *P = Xw + 3 * i;
        float pc[3];
        for (int r = 0; r < 3; r++) {
            float s = Tcw[4 * r] * P[0];
            s = s + Tcw[4 * r + 1] * P[1];
            s = s + Tcw[4 * r + 2] * P[2];
            pc[r] = s + Tcw[4 * r + 3];
        }
        if (pc[2] < 0.0f) continue;
        const float invz = 1.0f / pc[2];
        const float u = K4[0] * pc[0] * invz + K4[2];
        const float v = K4[1] * pc[1] * invz + K4[3];
        if (u < bounds4[0] || u > bounds4[2]) continue;
        if (v < bounds4[1] || v > bounds4[3]) continue;
        const float maxd = 1.2f * mf_max_dist[i], mind = 0.8f * mf_min_dist[i];
        const float po[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        double s2 = 0;
        for (int k = 0; k < 3; k++) s2 += (double)po[k] * (double)po[k];
        const float dist = (float)sqrt(s2);
        if (dist < mind || dist > maxd) continue;
        double dot = 0;
        for (int k = 0; k < 3; k++) dot += (double)po[k] * (double)normal[3 * i + k];
        const float vc = (float)(dot / dist);
        if (vc < cos_limit) continue;
        const float ratio = mf_max_dist[i] / dist;
        in_view[i] = 1;
        proj_xy[2 * i] = u; proj_xy[2 * i + 1] = v;
        pred_level[i] = (int)ceilf(logf(ratio) / log_scale_factor);
        view_cos[i] = vc;
    }
}

/* ======================================================================================== */
/* Optimizer::OptimizeSim3, Optimizer.cc:1348-1543: one VertexSim3Expmap against fixed points, EdgeSim3ProjectXYZ (x1 = S12 X2) and
 * EdgeInverseSim3ProjectXYZ (x2 = S21 X1) with Huber kernels, BlockSolverX + LinearSolverDense (7x7 LDLT), g2o's Levenberg.  The
 * edges have no analytic Jacobian in this g2o (types_seven_dof_expmap.h:147,169): BaseBinaryEdge::linearizeOplus differentiates
 * numerically, central differences with delta = 1e-9 through oplusImpl (base_binary_edge.hpp:131-205).
 *   g2o::Sim3 (types/sim3.h): r quaternion (x y z w), t, s; exp map :45-104; map :106-108; inverse :184-187; operator* :214-220.
 *   VertexSim3Expmap::oplusImpl (types_seven_dof_expmap.h:54-63): update[6] = 0 when _fix_scale; estimate <- Sim3(update) * estimate. */
typedef struct { double q[4]; double t[3]; double s; } sim3;

static void sim3_exp(const double u[7], sim3 *out)                 /* Sim3(const Vector7d&), sim3.h:45-104 */
{
    const double w[3] = {u[0], u[1], u[2]}, sigma = u[6];
    const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9], R[9], W[9];
    mat3_mul(O, O, O2);
    const double s = exp(sigma), eps = 0.00001;
    double A, Bc, C;
    if (fabs(sigma) < eps) {
        C = 1;
        if (theta < eps) {
            A = 1. / 2.; Bc = 1. / 6.;
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i];
        } else {
            const double theta2 = theta * theta;
            A = (1 - cos(theta)) / (theta2);
            Bc = (theta - sin(theta)) / (theta2 * theta);
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + sin(theta) / theta * O[i] + (1 - cos(theta)) / (theta * theta) * O2[i];
        }
    } else {
        C = (s - 1) / sigma;
        if (theta < eps) {
            const double sigma2 = sigma * sigma;
            A = ((sigma - 1) * s + 1) / sigma2;
            Bc = ((0.5 * sigma2 - sigma + 1) * s) / (sigma2 * sigma);
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i];
        } else {
            for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + sin(theta) / theta * O[i] + (1 - cos(theta)) / (theta * theta) * O2[i];
            const double a = s * sin(theta), b = s * cos(theta), theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
            A = (a * sigma + (1 - b) * theta) / (theta * c);
            Bc = (C - ((b - 1) * sigma + a * theta) / (c)) * 1. / (theta2);
        }
    }
    quat_from_R(R, out->q);
    for (int i = 0; i < 9; i++) W[i] = A * O[i] + Bc * O2[i] + C * (i % 4 == 0 ? 1.0 : 0.0);
    for (int r = 0; r < 3; r++) out->t[r] = W[3 * r] * u[3] + W[3 * r + 1] * u[4] + W[3 * r + 2] * u[5];
    out->s = s;
}

static void sim3_map(const sim3 *S, const double x[3], double out[3])          /* s*(r*xyz) + t */
{
    double rx[3];
    quat_rotate(S->q, x, rx);
    for (int k = 0; k < 3; k++) out[k] = S->s * rx[k] + S->t[k];
}

static void sim3_inverse(const sim3 *S, sim3 *out)                              /* Sim3(r.conjugate(), r.conjugate()*((-1./s)*t), 1./s) */
{
    const double qc[4] = {-S->q[0], -S->q[1], -S->q[2], S->q[3]};
    const double f = -1. / S->s, st[3] = {f * S->t[0], f * S->t[1], f * S->t[2]};
    quat_rotate(qc, st, out->t);
    memcpy(out->q, qc, sizeof qc);
    out->s = 1. / S->s;
}

static void sim3_mul(const sim3 *a, const sim3 *b, sim3 *o)                     /* operator*, sim3.h:214-220 */
{
    sim3 r;
    quat_mul(a->q, b->q, r.q);
    double rt[3];
    quat_rotate(a->q, b->t, rt);
    for (int k = 0; k < 3; k++) r.t[k] = a->s * rt[k] + a->t[k];
    r.s = a->s * b->s;
    *o = r;
}

typedef struct {
    int N; const float *P1c, *P2c, *obs1, *obs2, *w1, *w2; double K1[4], K2[4];
    int fix_scale; double delta, dsqr;
    uint8_t *active;             /* per correspondence: both edges still in the graph */
    double *err;                 /* [N,4]: stored _error of e12 and e21 */
} s3_t;

static void s3_errors(const s3_t *B, const sim3 *S, int i, double e[4])
{
    sim3 Si;
    double X2[3] = {B->P2c[3 * i], B->P2c[3 * i + 1], B->P2c[3 * i + 2]}, X1[3] = {B->P1c[3 * i], B->P1c[3 * i + 1], B->P1c[3 * i + 2]}, p[3];
    sim3_map(S, X2, p);                                             /* EdgeSim3ProjectXYZ::computeError */
    e[0] = (double)B->obs1[2 * i] - (p[0] / p[2] * B->K1[0] + B->K1[2]);
    e[1] = (double)B->obs1[2 * i + 1] - (p[1] / p[2] * B->K1[1] + B->K1[3]);
    sim3_inverse(S, &Si);                                           /* EdgeInverseSim3ProjectXYZ::computeError */
    sim3_map(&Si, X1, p);
    e[2] = (double)B->obs2[2 * i] - (p[0] / p[2] * B->K2[0] + B->K2[2]);
    e[3] = (double)B->obs2[2 * i + 1] - (p[1] / p[2] * B->K2[1] + B->K2[3]);
}

static void s3_oplus(const s3_t *B, const sim3 *S, const double *upd, sim3 *out)
{
    double u[7];
    memcpy(u, upd, sizeof u);
    if (B->fix_scale) u[6] = 0;
    sim3 d;
    sim3_exp(u, &d);
    sim3_mul(&d, S, out);
}

static double s3_chi2(double e0, double e1, double w) { return e0 * (w * e0 + 0.0 * e1) + e1 * (0.0 * e0 + w * e1); }

static double s3_active_errors(s3_t *B, const sim3 *S)             /* computeActiveErrors + activeRobustChi2 */
{
    double chi = 0;
    for (int i = 0; i < B->N; i++) {
        if (!B->active[i]) continue;
        double *e = &B->err[4 * i];
        s3_errors(B, S, i, e);
        for (int k = 0; k < 2; k++) {
            const double c = s3_chi2(e[2 * k], e[2 * k + 1], (double)(k ? B->w2[i] : B->w1[i]));
            double rho[3];
            if (c <= B->dsqr) { rho[0] = c; } else { rho[0] = 2 * sqrt(c) * B->delta - B->dsqr; }
            chi += rho[0];
        }
    }
    return chi;
}

static int ldlt7(const double *H, const double *b, double lambda, double *x)    /* LinearSolverDense: Eigen LDLT + isPositive */
{
    /* Eigen::LDLT (Eigen/src/Cholesky/LDLT.h, unblocked): symmetric pivoting on the largest remaining |diagonal|, then the rank-1 update of the
     * trailing block; isPositive() = no negative pivot; solve = P^T L^-T D^-1 L^-1 P b */
    double A[49], L[49], D[7], y[7];
    int perm[7];
    memcpy(A, H, sizeof A);
    for (int i = 0; i < 7; i++) for (int j = i + 1; j < 7; j++) A[7 * i + j] = A[7 * j + i];      /* only the lower triangle is referenced */
    for (int i = 0; i < 7; i++) { A[8 * i] += lambda; perm[i] = i; }
    for (int i = 0; i < 49; i++) L[i] = (i % 8 == 0) ? 1.0 : 0.0;
    int positive = 1;
    for (int k = 0; k < 7; k++) {
        int p = k;
        for (int i = k + 1; i < 7; i++) if (fabs(A[8 * i]) > fabs(A[8 * p])) p = i;
        if (p != k) {
            for (int j = 0; j < 7; j++) { const double t = A[7 * k + j]; A[7 * k + j] = A[7 * p + j]; A[7 * p + j] = t; }
            for (int i = 0; i < 7; i++) { const double t = A[7 * i + k]; A[7 * i + k] = A[7 * i + p]; A[7 * i + p] = t; }
            for (int j = 0; j < k; j++) { const double t = L[7 * k + j]; L[7 * k + j] = L[7 * p + j]; L[7 * p + j] = t; }
            const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        const double d = A[8 * k]; D[k] = d;
        if (d < 0) positive = 0;
        if (d == 0) continue;
        for (int i = k + 1; i < 7; i++) L[7 * i + k] = A[7 * i + k] / d;
        for (int i = k + 1; i < 7; i++) for (int j = k + 1; j < 7; j++) A[7 * i + j] -= L[7 * i + k] * A[7 * k + j];
    }
    if (!positive) return 0;
    for (int i = 0; i < 7; i++) y[i] = b[perm[i]];
    for (int i = 0; i < 7; i++) { double s = y[i]; for (int k = 0; k < i; k++) s -= L[7 * i + k] * y[k]; y[i] = s; }
    for (int i = 0; i < 7; i++) y[i] = D[i] != 0 ? y[i] / D[i] : 0.0;
    for (int i = 6; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 7; k++) s -= L[7 * k + i] * y[k]; y[i] = s; }
    for (int i = 0; i < 7; i++) x[perm[i]] = y[i];
    return 1;
}

/* SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg::solve on the single Sim3 vertex */
static void s3_optimize(s3_t *B, sim3 *S, int iterations, int *stats)
{
    int any = 0;
    for (int i = 0; i < B->N; i++) any |= B->active[i];
    if (!any) return;
    double lambda = 0, ni = 2;
    int nbad = 0;
    for (int it = 0; it < iterations; it++) {
        double currentChi = s3_active_errors(B, S), tempChi;
        const double iniChi = currentChi;
        double H[49], b[7], x[7];
        memset(H, 0, sizeof H); memset(b, 0, sizeof b);
        for (int i = 0; i < B->N; i++) {                            /* buildSystem: linearizeOplus (numeric) + constructQuadraticForm */
            if (!B->active[i]) continue;
            double J[4][7];
            const double delta = 1e-9, scalar = 1.0 / (2 * delta);
            for (int d = 0; d < 7; d++) {
                double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[4], em[4];
                sim3 Sp;
                add[d] = delta; s3_oplus(B, S, add, &Sp); s3_errors(B, &Sp, i, ep);
                add[d] = -delta; s3_oplus(B, S, add, &Sp); s3_errors(B, &Sp, i, em);
                for (int r = 0; r < 4; r++) J[r][d] = scalar * (ep[r] - em[r]);
            }
            const double *e = &B->err[4 * i];
            for (int k = 0; k < 2; k++) {
                const double w = (double)(k ? B->w2[i] : B->w1[i]);
                const double e0 = e[2 * k], e1 = e[2 * k + 1];
                const double c = s3_chi2(e0, e1, w);
                double rho1 = 1.;
                if (c > B->dsqr) rho1 = B->delta / sqrt(c);
                /* constructQuadraticForm, base_binary_edge.hpp:55-120, in g2o's evaluation order: omega_r = -(information * error), then *= rho'; the FULL
                 * block B^T (rho' information) B is accumulated entry by entry ((J_a w) J_c and (J_c w) J_a round differently; Eigen's LDLT reads the lower
                 * triangle) -- pinned against the reference's object code */
                const double wo = rho1 * w, r0 = -(w * e0) * rho1, r1 = -(w * e1) * rho1;
                const double *J0 = J[2 * k], *J1 = J[2 * k + 1];
                for (int a = 0; a < 7; a++) {
                    b[a] += J0[a] * r0 + J1[a] * r1;
                    for (int cc = 0; cc < 7; cc++) H[7 * a + cc] += J0[a] * wo * J0[cc] + J1[a] * wo * J1[cc];
                }
            }
        }
        if (it == 0) {                                              /* computeLambdaInit */
            double mx = 0;
            for (int a = 0; a < 7; a++) mx = fmax(fabs(H[8 * a]), mx);
            lambda = 1e-5 * mx; ni = 2; nbad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const sim3 backup = *S;
            const int ok2 = ldlt7(H, b, lambda, x);
            sim3 Sn;
            s3_oplus(B, S, x, &Sn);
            *S = Sn;
            tempChi = s3_active_errors(B, S);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0.;
            for (int j = 0; j < 7; j++) scale += x[j] * (lambda * x[j] + b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha); ni = 2; currentChi = tempChi;
            } else { lambda *= ni; ni *= 2; *S = backup; }
            qmax++;
            if (stats) stats[1]++;
        } while (rho < 0 && qmax < 10);
        if (stats) stats[0]++;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nbad++; else nbad = 0;
        if (nbad >= 3) break;
    }
}

/* sim3_io: r (x y z w), t, s of g2oS12 in/out.  P1c / P2c f32[N,3]: the matched map points in their own camera frames (R1w*P3D1w+t1w,
 * Optimizer.cc:1411-1424, float as cv::Mat computes them); valid u8[N]: the pair makes a correspondence (:1398-1433); obs1 / obs2 f32[N,2]:
 * kpUn.pt in KF1 / KF2; w1 / w2: mvInvLevelSigma2 of the keypoint octaves.  inlier u8[N] out: vpMatches1[i] is kept.
 * Returns nIn (0 with g2oS12 untouched when fewer than 10 correspondences survive the first pass, :1497-1498).  stats: LM iterations, trials. */
int oracle_optimize_sim3(double *sim3_io, int N, const uint8_t *valid, const float *P1c, const float *P2c, const float *obs1, const float *obs2,
                         const float *w1, const float *w2, const float *K1, const float *K2, float th2, int fix_scale, uint8_t *inlier, int *stats)
{
    s3_t B; memset(&B, 0, sizeof B);
    B.N = N; B.P1c = P1c; B.P2c = P2c; B.obs1 = obs1; B.obs2 = obs2; B.w1 = w1; B.w2 = w2; B.fix_scale = fix_scale;
    for (int k = 0; k < 4; k++) { B.K1[k] = K1[k]; B.K2[k] = K2[k]; }
    const float deltaHuber = sqrtf(th2);                            /* const float deltaHuber = sqrt(th2) */
    B.delta = deltaHuber; B.dsqr = (double)(float)(B.delta * B.delta);   /* RobustKernelHuber keeps delta^2 in a FLOAT member (robust_kernel_impl.h:84, set by setDelta, robust_kernel_impl.cpp:65-69) */
    B.active = (uint8_t *)malloc(N > 0 ? N : 1); B.err = (double *)calloc(4 * (size_t)(N > 0 ? N : 1), sizeof(double));
    if (stats) stats[0] = stats[1] = 0;
    sim3 S;
    memcpy(S.q, sim3_io, 4 * sizeof(double)); memcpy(S.t, sim3_io + 4, 3 * sizeof(double)); S.s = sim3_io[7];
    int nCorr = 0;
    for (int i = 0; i < N; i++) { B.active[i] = valid[i] ? 1 : 0; inlier[i] = B.active[i]; nCorr += B.active[i]; }
    s3_optimize(&B, &S, 5, stats);
    int nBad = 0;
    for (int i = 0; i < N; i++) {
        if (!B.active[i]) continue;
        const double *e = &B.err[4 * i];
        if (s3_chi2(e[0], e[1], (double)w1[i]) > th2 || s3_chi2(e[2], e[3], (double)w2[i]) > th2) { B.active[i] = 0; inlier[i] = 0; nBad++; }
    }
    const int more = nBad > 0 ? 10 : 5;
    int nIn = 0;
    if (nCorr - nBad >= 10) {
        s3_optimize(&B, &S, more, stats);
        for (int i = 0; i < N; i++) {
            if (!B.active[i]) continue;
            const double *e = &B.err[4 * i];
            if (s3_chi2(e[0], e[1], (double)w1[i]) > th2 || s3_chi2(e[2], e[3], (double)w2[i]) > th2) inlier[i] = 0;
            else nIn++;
        }
        memcpy(sim3_io, S.q, 4 * sizeof(double)); memcpy(sim3_io + 4, S.t, 3 * sizeof(double)); sim3_io[7] = S.s;
    }
    free(B.active); free(B.err);
    return nIn;
}

/* ======================================================================================== */
/* Vocabulary-bucket matchers and SearchForInitialization (matcher, continued)              */
/* A DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) is passed as CSR: ascending node ids, start offsets, feature indices in
 * push order.  The three bucket matchers walk the two maps like a merge join (ORBmatcher.cc:176-269, :548-634, :695-790) and only ever
 * compare features of equal node id, so the restatement iterates over the common nodes.
 *   mode 0: SearchByBoW(KeyFrame*, Frame&, ..) :159-290 and SearchByBoW(KeyFrame*, KeyFrame*, ..) :524-657 -- best / second best over the
 *           still unclaimed side-2 features of the bucket, accept if best <= TH_LOW and (float)best < ratio * (float)second, claim.
 *   mode 1: SearchForTriangulation :659-825 (monocular: mvuRight < 0, bOnlyStereo = false) -- no claims (vbMatched2 is never set), best
 *           distance <= TH_LOW with the LAST candidate winning ties (dist > bestDist -> skip), epipole-distance and epipolar-line gates.
 * elig1 / elig2: the feature takes part (has a good map point for mode 0 / has none for mode 1).  match12[idx1] = idx2 or -1. */
typedef struct { int n_nodes; const int *nodes; const int *start; const int *items; } oracle_featvec;
typedef struct { const float *xy1, *xy2; const int *octave2; const float *F12; float ex, ey; const float *scale_factors2, *level_sigma2_2; } oracle_epipolar;

static int check_dist_epipolar_line(const float *p1, const float *p2, const float *F12, float sigma2)     /* ORBmatcher.cc:140-157 */
{
    const float a = p1[0] * F12[0] + p1[1] * F12[3] + F12[6];
    const float b = p1[0] * F12[1] + p1[1] * F12[4] + F12[7];
    const float c = p1[0] * F12[2] + p1[1] * F12[5] + F12[8];
    const float num = a * p2[0] + b * p2[1] + c;
    const float den = a * a + b * b;
    if (den == 0) return 0;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2;
}

int oracle_search_by_bow(int mode, int N1, const uint8_t *desc1, const float *angle1, const uint8_t *elig1, const oracle_featvec *fv1,
                         int N2, const uint8_t *desc2, const float *angle2, const uint8_t *elig2, const oracle_featvec *fv2,
                         float ratio, int check_ori, const oracle_epipolar *ep, int *match12)
{
    uint8_t *matched2 = (uint8_t *)calloc(N2 > 0 ? N2 : 1, 1);
    int *rot_bin = (int *)malloc(sizeof(int) * (N1 > 0 ? N1 : 1));
    int hist[HISTO_LENGTH];
    memset(hist, 0, sizeof hist);
    for (int i = 0; i < N1; i++) { match12[i] = -1; rot_bin[i] = -1; }
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0, a = 0, b = 0;
    while (a < fv1->n_nodes && b < fv2->n_nodes) {
        if (fv1->nodes[a] < fv2->nodes[b]) { a++; continue; }           /* lower_bound on the other map */
        if (fv1->nodes[a] > fv2->nodes[b]) { b++; continue; }
        for (int u = fv1->start[a]; u < fv1->start[a + 1]; u++) {
            const int idx1 = fv1->items[u];
            if (!elig1[idx1]) continue;
            int best1 = mode == 0 ? 256 : TH_LOW, best2 = 256, best_idx = -1;
            for (int v = fv2->start[b]; v < fv2->start[b + 1]; v++) {
                const int idx2 = fv2->items[v];
                if (mode == 0) {
                    if (matched2[idx2] || !elig2[idx2]) continue;
                    const int d = oracle_descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (d < best1) { best2 = best1; best1 = d; best_idx = idx2; }
                    else if (d < best2) best2 = d;
                } else {
                    if (matched2[idx2] || !elig2[idx2]) continue;
                    const int d = oracle_descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (d > TH_LOW || d > best1) continue;
                    const float distex = ep->ex - ep->xy2[2 * idx2], distey = ep->ey - ep->xy2[2 * idx2 + 1];
                    if (distex * distex + distey * distey < 100 * ep->scale_factors2[ep->octave2[idx2]]) continue;
                    if (check_dist_epipolar_line(ep->xy1 + 2 * idx1, ep->xy2 + 2 * idx2, ep->F12, ep->level_sigma2_2[ep->octave2[idx2]])) { best_idx = idx2; best1 = d; }
                }
            }
            int accept;
            if (mode == 0) accept = best1 <= TH_LOW && (float)best1 < ratio * (float)best2;
            else accept = best_idx >= 0;
            if (!accept) continue;
            match12[idx1] = best_idx;
            if (mode == 0) matched2[best_idx] = 1;
            nmatches++;
            if (check_ori) {
                float rot = angle1[idx1] - angle2[best_idx];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rot_bin[idx1] = bin; hist[bin]++;
            }
        }
        a++; b++;
    }
    if (check_ori) {
        int i1, i2, i3;
        three_maxima(hist, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < N1; i++) { const int bn = rot_bin[i]; if (bn >= 0 && bn != i1 && bn != i2 && bn != i3) { match12[i] = -1; nmatches--; } }
    }
    free(matched2); free(rot_bin);
    return nmatches;
}

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:407-522: level-0 keypoints of F1 search a
 * window around their previous match in F2 (GetFeaturesInArea(x, y, windowSize, 0, 0)); a candidate already matched with a smaller-or-equal
 * distance is skipped, a better one is stolen (its former owner loses the match but stays in the rotation histogram, :471-476, :497-505).
 * prev_matched f32[N1,2] in/out (:517-519).  Returns nmatches. */
int oracle_search_for_initialization(const oracle_grid_params *g, int N1, const float *xy1_unused, const int *octave1, const float *angle1, const uint8_t *desc1,
                                     int N2, const float *xy2, const int *octave2, const float *angle2, const uint8_t *desc2,
                                     float *prev_matched, int window, float ratio, int check_ori, int *matches12)
{
    (void)xy1_unused;
    const int nc = GRID_COLS * GRID_ROWS;
    int *cell_start = (int *)malloc(sizeof(int) * (nc + 1));
    int *cell_items = (int *)malloc(sizeof(int) * (N2 > 0 ? N2 : 1)), *cands = (int *)malloc(sizeof(int) * (N2 > 0 ? N2 : 1));
    int *matched_dist = (int *)malloc(sizeof(int) * (N2 > 0 ? N2 : 1)), *matches21 = (int *)malloc(sizeof(int) * (N2 > 0 ? N2 : 1));
    int *rot_bin = (int *)malloc(sizeof(int) * (N1 > 0 ? N1 : 1));
    int hist[HISTO_LENGTH];
    memset(hist, 0, sizeof hist);
    oracle_grid_build(g, N2, xy2, cell_start, cell_items);
    for (int i = 0; i < N2; i++) { matched_dist[i] = 0x7fffffff; matches21[i] = -1; }
    for (int i = 0; i < N1; i++) { matches12[i] = -1; rot_bin[i] = -1; }
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int level1 = octave1[i1];
        if (level1 > 0) continue;
        const int n = oracle_features_in_area(g, cell_start, cell_items, xy2, octave2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window, level1, level1, cands);
        if (n == 0) continue;
        int best = 0x7fffffff, best2 = 0x7fffffff, best_idx = -1;
        for (int c = 0; c < n; c++) {
            const int i2 = cands[c];
            const int d = oracle_descriptor_distance(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
            if (matched_dist[i2] <= d) continue;
            if (d < best) { best2 = best; best = d; best_idx = i2; }
            else if (d < best2) best2 = d;
        }
        if (best <= TH_LOW) {
            if (best < (float)best2 * ratio) {
                if (matches21[best_idx] >= 0) { matches12[matches21[best_idx]] = -1; nmatches--; }
                matches12[i1] = best_idx; matches21[best_idx] = i1; matched_dist[best_idx] = best;
                nmatches++;
                if (check_ori) {
                    float rot = angle1[i1] - angle2[best_idx];
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rot_bin[i1] = bin; hist[bin]++;
                }
            }
        }
    }
    if (check_ori) {
        int i1, i2, i3;
        three_maxima(hist, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < N1; i++) {
            const int bn = rot_bin[i];
            if (bn >= 0 && bn != i1 && bn != i2 && bn != i3 && matches12[i] >= 0) { matches12[i] = -1; nmatches--; }
        }
    }
    for (int i = 0; i < N1; i++) if (matches12[i] >= 0) { prev_matched[2 * i] = xy2[2 * matches12[i]]; prev_matched[2 * i + 1] = xy2[2 * matches12[i] + 1]; }
    free(cell_start); free(cell_items); free(cands); free(matched_dist); free(matches21); free(rot_bin);
    return nmatches;
}

/* ======================================================================================== */
/* DBoW2 vocabulary transform: TemplatedVocabulary<TDescriptor,F>::transform(features, v, fv, levelsup), S/Thirdparty/DBoW2/DBoW2/
 * TemplatedVocabulary.h:1127-1177 (TF_IDF weighting, L1 norm: the only configuration ORBSLAMM loads) with the per-feature tree walk
 * :1217-1262, BowVector::addWeight / normalize (BowVector.cpp:34-84) and FORB::distance (FORB.cpp:81-101 = the same bit-hack Hamming).
 * The tree is given flat: children of node i = child_ids[child_start[i] .. child_start[i+1]) in Node::children order; leaves carry
 * word_id / weight.  Outputs: per-feature word / node (-1 = stopped word), the BowVector as ascending (word, value) pairs and the
 * FeatureVector as CSR.  Returns the number of words in the BowVector. */
int oracle_vocab_transform(int L, const uint8_t *node_desc, const int *child_start, const int *child_ids, const int *word_id, const double *weight,
                           int N, const uint8_t *desc, int levelsup, int *word_of, int *node_of, int *bow_ids, double *bow_vals,
                           int *fv_nodes, int *fv_start, int *fv_items, int *fv_count)
{
    const int nid_level = L - levelsup;
    int nb = 0, nf = 0;
    for (int i = 0; i < N; i++) {
        const uint8_t *f = desc + 32 * (size_t)i;
        int nid = 0, final_id = 0, level = 0;                      /* transform(feature, id, w, &nid, levelsup), :1217-1262 */
        do {
            ++level;
            const int c0 = child_start[final_id], c1 = child_start[final_id + 1];
            final_id = child_ids[c0];
            double best_d = oracle_descriptor_distance(f, node_desc + 32 * (size_t)final_id);
            for (int c = c0 + 1; c < c1; c++) {
                const int id = child_ids[c];
                const double d = oracle_descriptor_distance(f, node_desc + 32 * (size_t)id);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (level == nid_level) nid = final_id;
        } while (child_start[final_id] != child_start[final_id + 1]);
        const double w = weight[final_id];
        word_of[i] = -1; node_of[i] = -1;
        if (!(w > 0)) continue;                                    /* stopped word */
        word_of[i] = word_id[final_id]; node_of[i] = nid;
        {   /* v.addWeight(id, w): std::map lower_bound, += or insert (BowVector.cpp:34-46) */
            int lo = 0, hi = nb;
            while (lo < hi) { const int mid = (lo + hi) / 2; if (bow_ids[mid] < word_of[i]) lo = mid + 1; else hi = mid; }
            if (lo < nb && bow_ids[lo] == word_of[i]) bow_vals[lo] += w;
            else {
                memmove(bow_ids + lo + 1, bow_ids + lo, sizeof(int) * (nb - lo)); memmove(bow_vals + lo + 1, bow_vals + lo, sizeof(double) * (nb - lo));
                bow_ids[lo] = word_of[i]; bow_vals[lo] = w; nb++;
            }
        }
        {   /* fv.addFeature(nid, i): node list kept ascending, the feature lists are rebuilt below */
            int lo = 0, hi = nf;
            while (lo < hi) { const int mid = (lo + hi) / 2; if (fv_nodes[mid] < nid) lo = mid + 1; else hi = mid; }
            if (!(lo < nf && fv_nodes[lo] == nid)) { memmove(fv_nodes + lo + 1, fv_nodes + lo, sizeof(int) * (nf - lo)); fv_nodes[lo] = nid; nf++; }
        }
    }
    /* feature lists per node in push order (= ascending feature index) */
    int pos = 0;
    for (int a = 0; a < nf; a++) {
        fv_start[a] = pos;
        for (int i = 0; i < N; i++) if (node_of[i] == fv_nodes[a]) fv_items[pos++] = i;
    }
    fv_start[nf] = pos;
    *fv_count = nf;
    /* must = mustNormalize(L1) -> v.normalize(L1), BowVector.cpp:62-84 */
    double norm = 0.0;
    for (int a = 0; a < nb; a++) norm += fabs(bow_vals[a]);
    if (norm > 0.0) for (int a = 0; a < nb; a++) bow_vals[a] /= norm;
    return nb;
}

/* ======================================================================================== */
/* Sim3Solver (S/src/Sim3Solver.cc): the data-parallel part of the RANSAC -- FromCameraToImage (:419-437), Project (:392-417) and CheckInliers
 * (:340-365) -- for n_hyp hypotheses (T12, T21 as produced by ComputeSim3, row-major 4x4 float) over N correspondences.  All float; cv::Mat R*P+t
 * = small gemm then add; Mat::dot accumulates in double and the result is stored in a float; the error thresholds are the reference's
 * std::vector<size_t> entries, i.e. 9.210 * sigma^2 TRUNCATED to an integer (:88-89, Sim3Solver.h:78-79). */
void oracle_sim3_max_error(int N, const int *octave, const float *level_sigma2, int *max_err)
{
    for (int i = 0; i < N; i++) max_err[i] = (int)(size_t)(9.210 * level_sigma2[octave[i]]);
}

void oracle_sim3_from_camera_to_image(int N, const float *X3Dc, const float *K4, float *p2d)
{
    for (int i = 0; i < N; i++) {
        const float invz = 1 / (X3Dc[3 * i + 2]);
        const float x = X3Dc[3 * i] * invz, y = X3Dc[3 * i + 1] * invz;
        p2d[2 * i] = K4[0] * x + K4[2]; p2d[2 * i + 1] = K4[1] * y + K4[3];
    }
}

static void sim3_project(const float *T, const float *K4, const float *P, float *uv)
{
    float pc[3];
    for (int r = 0; r < 3; r++) {
        float s = T[4 * r] * P[0];
        s = s + T[4 * r + 1] * P[1];
        s = s + T[4 * r + 2] * P[2];
        pc[r] = s + T[4 * r + 3];
    }
    const float invz = 1 / (pc[2]);
    const float x = pc[0] * invz, y = pc[1] * invz;
    uv[0] = K4[0] * x + K4[2]; uv[1] = K4[1] * y + K4[3];
}

void oracle_sim3_check_inliers(int n_hyp, const float *T12, const float *T21, int N, const float *X3Dc1, const float *X3Dc2, const float *P1im1,
                               const float *P2im2, const int *max_err1, const int *max_err2, const float *K1, const float *K2,
                               uint8_t *inliers, int *n_inliers)
{
    for (int h = 0; h < n_hyp; h++) {
        int n = 0;
        for (int i = 0; i < N; i++) {
            float p2im1[2], p1im2[2];
            sim3_project(T12 + 16 * h, K1, X3Dc2 + 3 * i, p2im1);          /* Project(mvX3Dc2, vP2im1, mT12i, mK1) */
            sim3_project(T21 + 16 * h, K2, X3Dc1 + 3 * i, p1im2);          /* Project(mvX3Dc1, vP1im2, mT21i, mK2) */
            const float d1[2] = {P1im1[2 * i] - p2im1[0], P1im1[2 * i + 1] - p2im1[1]};
            const float d2[2] = {p1im2[0] - P2im2[2 * i], p1im2[1] - P2im2[2 * i + 1]};
            const float err1 = (float)((double)d1[0] * (double)d1[0] + (double)d1[1] * (double)d1[1]);
            const float err2 = (float)((double)d2[0] * (double)d2[0] + (double)d2[1] * (double)d2[1]);
            const int in = err1 < (float)max_err1[i] && err2 < (float)max_err2[i];
            inliers[(size_t)h * N + i] = (uint8_t)in;
            n += in;
        }
        n_inliers[h] = n;
    }
}

/* ======================================================================================== */
/* Optimizer::OptimizeEssentialGraph / MMOptimizeEssentialGraph, numeric core (S/src/Optimizer.cc:804-1067, 1069-1346): a pose graph of
 * VertexSim3Expmap vertices and EdgeSim3 edges (error = log(Sji * Siw * Sjw^-1), identity information, no robust kernel; g2o/types/
 * types_seven_dof_expmap.h:64-84, sim3.h:110-181), numeric Jacobians for both vertices (base_binary_edge.hpp:131-205, delta = 1e-9), g2o's
 * Levenberg with setUserLambdaInit(1e-16) and 20 iterations, BlockSolver_7_3 + LinearSolverEigen (a direct sparse Cholesky; here a dense LDLT,
 * equal up to fp64 rounding).  The graph assembly (which keyframes, which edges) is the shim's / caller's job. */
static void lu3_solve(const double A[9], const double b[3], double x[3])          /* Eigen PartialPivLU 3x3 */
{
    double M[3][4] = {{A[0], A[1], A[2], b[0]}, {A[3], A[4], A[5], b[1]}, {A[6], A[7], A[8], b[2]}};
    for (int c = 0; c < 3; c++) {
        int p = c;
        for (int r = c + 1; r < 3; r++) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
        if (p != c) for (int k = 0; k < 4; k++) { const double t = M[c][k]; M[c][k] = M[p][k]; M[p][k] = t; }
        for (int r = c + 1; r < 3; r++) { const double f = M[r][c] / M[c][c]; for (int k = c; k < 4; k++) M[r][k] -= f * M[c][k]; }
    }
    for (int r = 2; r >= 0; r--) { double s = M[r][3]; for (int k = r + 1; k < 3; k++) s -= M[r][k] * x[k]; x[r] = s / M[r][r]; }
}

static void sim3_log(const sim3 *S, double res[7])                                 /* Sim3::log, sim3.h:110-181 */
{
    const double sigma = log(S->s);
    double R[9], omega[3], O[9], O2[9], W[9];
    quat_to_R(S->q, R);
    const double d = 0.5 * (R[0] + R[4] + R[8] - 1), eps = 0.00001;
    const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};                  /* deltaR(R) */
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
        C = (S->s - 1) / sigma;
        if (d > 1 - eps) {
            const double sigma2 = sigma * sigma;
            for (int k = 0; k < 3; k++) omega[k] = 0.5 * dR[k];
            A = ((sigma - 1) * S->s + 1) / (sigma2); B = ((0.5 * sigma2 - sigma + 1) * S->s) / (sigma2 * sigma);
        } else {
            const double theta = acos(d);
            for (int k = 0; k < 3; k++) omega[k] = theta / (2 * sqrt(1 - d * d)) * dR[k];
            const double theta2 = theta * theta, a = S->s * sin(theta), b = S->s * cos(theta), c = theta2 + sigma * sigma;
            A = (a * sigma + (1 - b) * theta) / (theta * c); B = (C - ((b - 1) * sigma + a * theta) / (c)) * 1. / (theta2);
        }
    }
    O[0] = 0; O[1] = -omega[2]; O[2] = omega[1]; O[3] = omega[2]; O[4] = 0; O[5] = -omega[0]; O[6] = -omega[1]; O[7] = omega[0]; O[8] = 0;
    /* W = A*Omega + B*Omega*Omega + C*I evaluates left to right: (B*Omega)*Omega -- pinned against the reference's object code (1 ulp apart from B*(Omega*Omega)) */
    double BO[9];
    for (int i = 0; i < 9; i++) BO[i] = B * O[i];
    mat3_mul(BO, O, O2);
    for (int i = 0; i < 9; i++) W[i] = A * O[i] + O2[i] + C * (i % 4 == 0 ? 1.0 : 0.0);
    double ups[3];
    lu3_solve(W, S->t, ups);
    for (int k = 0; k < 3; k++) { res[k] = omega[k]; res[k + 3] = ups[k]; }
    res[6] = sigma;
}

static void pg_edge_error(const sim3 *meas, const sim3 *Si, const sim3 *Sj, double e[7])      /* EdgeSim3::computeError */
{
    sim3 Sjinv, a, b;
    sim3_inverse(Sj, &Sjinv);
    sim3_mul(meas, Si, &a);
    sim3_mul(&a, &Sjinv, &b);
    sim3_log(&b, e);
}

static void pg_oplus(const sim3 *S, const double *upd, int fix_scale, sim3 *out)
{
    double u[7];
    memcpy(u, upd, sizeof u);
    if (fix_scale) u[6] = 0;
    sim3 d;
    sim3_exp(u, &d);
    sim3_mul(&d, S, out);
}

static int ldlt_dense(double *A, int n, const double *b, double *x)              /* unpivoted LDL^T, zero pivot -> failure */
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k] * A[(size_t)k * n + k];
            if (j < i) A[(size_t)i * n + j] = s / A[(size_t)j * n + j];
            else { if (s == 0.0 || !isfinite(s)) return 0; A[(size_t)i * n + i] = s; }
        }
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * x[k]; x[i] = s; }
    for (int i = 0; i < n; i++) x[i] /= A[(size_t)i * n + i];
    for (int i = n - 1; i >= 0; i--) { const double xi = x[i]; for (int k = 0; k < i; k++) x[k] -= A[(size_t)i * n + k] * xi; }
    return 1;
}

/* sim3 [K,8] in/out (r xyzw, t, s), fixed u8[K]; edges: vertex[0] = e_i, vertex[1] = e_j, measurement e_meas [E,8] (= Sji).
 * stats (may be NULL): LM iterations, LM trials, failed factorisations.  Returns the number of iterations executed. */
int oracle_optimize_pose_graph(int K, double *sim3_io, const uint8_t *fixed, int E, const int *e_i, const int *e_j, const double *e_meas, int fix_scale,
                               int iterations, double lambda_init, int *stats)
{
    sim3 *V = (sim3 *)malloc(sizeof(sim3) * (K > 0 ? K : 1)), *bak = (sim3 *)malloc(sizeof(sim3) * (K > 0 ? K : 1)), *M = (sim3 *)malloc(sizeof(sim3) * (E > 0 ? E : 1));
    int *hidx = (int *)malloc(sizeof(int) * (K > 0 ? K : 1));
    int nA = 0;
    for (int k = 0; k < K; k++) { memcpy(V[k].q, sim3_io + 8 * k, 32); memcpy(V[k].t, sim3_io + 8 * k + 4, 24); V[k].s = sim3_io[8 * k + 7]; hidx[k] = fixed[k] ? -1 : nA++; }
    for (int e = 0; e < E; e++) { memcpy(M[e].q, e_meas + 8 * e, 32); memcpy(M[e].t, e_meas + 8 * e + 4, 24); M[e].s = e_meas[8 * e + 7]; }
    const int n = 7 * nA;
    double *H = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (n > 0 ? n : 1)), *Hw = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (n > 0 ? n : 1));
    double *b = (double *)malloc(sizeof(double) * (n > 0 ? n : 1)), *x = (double *)calloc(n > 0 ? n : 1, sizeof(double)), *err = (double *)malloc(sizeof(double) * 7 * (E > 0 ? E : 1));
    int its = 0, trials = 0, fails = 0, nbad = 0;
    double lambda = 0, ni = 2;
#define PG_ACTIVE(e) (hidx[e_i[e]] >= 0 || hidx[e_j[e]] >= 0)
#define PG_CHI2(out) do { out = 0; for (int e = 0; e < E; e++) { if (!PG_ACTIVE(e)) continue; pg_edge_error(&M[e], &V[e_i[e]], &V[e_j[e]], &err[7 * e]); for (int k = 0; k < 7; k++) out += err[7 * e + k] * err[7 * e + k]; } } while (0)
    for (int it = 0; it < iterations && n > 0; it++) {
        double currentChi, tempChi;
        PG_CHI2(currentChi);
        const double iniChi = currentChi;
        memset(H, 0, sizeof(double) * (size_t)n * n); memset(b, 0, sizeof(double) * n);
        for (int e = 0; e < E; e++) {
            if (!PG_ACTIVE(e)) continue;
            double J[2][7][7];                                                   /* [vertex][row][col] */
            const int vid[2] = {e_i[e], e_j[e]};
            for (int s = 0; s < 2; s++) {
                if (hidx[vid[s]] < 0) continue;
                for (int d = 0; d < 7; d++) {
                    double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[7], em[7];
                    sim3 P;
                    add[d] = 1e-9; pg_oplus(&V[vid[s]], add, fix_scale, &P);
                    pg_edge_error(&M[e], s == 0 ? &P : &V[vid[0]], s == 1 ? &P : &V[vid[1]], ep);
                    add[d] = -1e-9; pg_oplus(&V[vid[s]], add, fix_scale, &P);
                    pg_edge_error(&M[e], s == 0 ? &P : &V[vid[0]], s == 1 ? &P : &V[vid[1]], em);
                    for (int r = 0; r < 7; r++) J[s][r][d] = (1.0 / (2 * 1e-9)) * (ep[r] - em[r]);
                }
            }
            for (int s = 0; s < 2; s++) {
                const int hs = hidx[vid[s]];
                if (hs < 0) continue;
                for (int a = 0; a < 7; a++) { double g = 0; for (int r = 0; r < 7; r++) g += J[s][r][a] * err[7 * e + r]; b[7 * hs + a] -= g; }
                for (int t = 0; t < 2; t++) {
                    const int ht = hidx[vid[t]];
                    if (ht < 0) continue;
                    for (int a = 0; a < 7; a++) for (int c = 0; c < 7; c++) { double g = 0; for (int r = 0; r < 7; r++) g += J[s][r][a] * J[t][r][c]; H[(size_t)(7 * hs + a) * n + 7 * ht + c] += g; }
                }
            }
        }
        if (it == 0) { lambda = lambda_init > 0 ? lambda_init : 0; if (!(lambda_init > 0)) { double mx = 0; for (int a = 0; a < n; a++) mx = fmax(fabs(H[(size_t)a * n + a]), mx); lambda = 1e-5 * mx; } ni = 2; nbad = 0; }
        double rho = 0;
        int qmax = 0;
        do {
            memcpy(bak, V, sizeof(sim3) * K);
            memcpy(Hw, H, sizeof(double) * (size_t)n * n);
            for (int a = 0; a < n; a++) Hw[(size_t)a * n + a] += lambda;
            const int ok2 = ldlt_dense(Hw, n, b, x);
            if (!ok2) fails++;
            for (int k = 0; k < K; k++) if (hidx[k] >= 0) { sim3 nv; pg_oplus(&V[k], &x[7 * hidx[k]], fix_scale, &nv); V[k] = nv; }
            PG_CHI2(tempChi);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0.;
            for (int a = 0; a < n; a++) scale += x[a] * (lambda * x[a] + b[a]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha); ni = 2; currentChi = tempChi;
            } else { lambda *= ni; ni *= 2; memcpy(V, bak, sizeof(sim3) * K); }
            qmax++; trials++;
        } while (rho < 0 && qmax < 10);
        its++;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nbad++; else nbad = 0;
        if (nbad >= 3) break;
    }
    for (int k = 0; k < K; k++) { memcpy(sim3_io + 8 * k, V[k].q, 32); memcpy(sim3_io + 8 * k + 4, V[k].t, 24); sim3_io[8 * k + 7] = V[k].s; }
    if (stats) { stats[0] = its; stats[1] = trials; stats[2] = fails; }
    free(V); free(bak); free(M); free(hidx); free(H); free(Hw); free(b); free(x); free(err);
    return its;
}

/* ======================================================================================== */
/* Probes: single primitives of the restatement, exported so that tests can compare them with the reference's own g2o object code
 * (oracle/_ref/libref_optimizer.so, ref_g2o_* in oracle/ref_optimizer_capi.cc) on random inputs.                                   */
static void probe_ba1(ba_t *B, se3 *pose, uint8_t *fixed, double *pt, int *ekf, int *ept, double *obs, double *w, double *err, const double *q, const double *t,
                      const double *Xw, const double *o2, double weight, const double *K4, int only_pose)
{
    memset(B, 0, sizeof *B);
    memcpy(pose->q, q, 4 * sizeof(double)); memcpy(pose->t, t, 3 * sizeof(double));
    quat_normalize_pos(pose->q);                         /* SE3Quat(const Quaterniond&, const Vector3d&) normalises, se3quat.h:66-69 */
    *fixed = 0; memcpy(pt, Xw, 3 * sizeof(double)); *ekf = 0; *ept = 0; obs[0] = o2[0]; obs[1] = o2[1]; *w = weight;
    B->K = 1; B->P = 1; B->E = 1; B->pose = pose; B->pose_fixed = fixed; B->intr = K4; B->pt = pt; B->pt_fixed_all = (uint8_t)only_pose;
    B->e_kf = ekf; B->e_pt = ept; B->e_obs = obs; B->e_w = w; B->e_err = err;
}
void oracle_probe_se3_exp(const double *u6, double *q, double *t) { se3 s; se3_exp(u6, &s); memcpy(q, s.q, 4 * sizeof(double)); memcpy(t, s.t, 3 * sizeof(double)); }
void oracle_probe_se3_oplus(double *q, double *t, const double *u6)
{
    se3 T, d, r; memcpy(T.q, q, 4 * sizeof(double)); memcpy(T.t, t, 3 * sizeof(double));
    quat_normalize_pos(T.q);                             /* the SE3Quat(q, t) constructor of the caller normalises */
    se3_exp(u6, &d); se3_mul(&d, &T, &r);
    memcpy(q, r.q, 4 * sizeof(double)); memcpy(t, r.t, 3 * sizeof(double));
}
void oracle_probe_converter_to_se3quat(const float *Tcw, double *q, double *t) { se3 s; se3_from_Tcw_f32(Tcw, &s); memcpy(q, s.q, 4 * sizeof(double)); memcpy(t, s.t, 3 * sizeof(double)); }
void oracle_probe_converter_to_cvmat(const double *q, const double *t, float *Tcw) { se3 s; memcpy(s.q, q, 4 * sizeof(double)); memcpy(s.t, t, 3 * sizeof(double)); quat_normalize_pos(s.q); se3_to_Tcw_f32(&s, Tcw); }
void oracle_probe_edge_se3(const double *q, const double *t, const double *Xw, const double *obs2, double inv_sigma2, const double *K4, int only_pose,
                           double *err2, double *chi2, int *depth_ok, double *Jpoint6, double *Jpose12)
{
    ba_t B; se3 pose; uint8_t fixed; double pt[3], obs[2], w, err[2]; int ekf, ept;
    probe_ba1(&B, &pose, &fixed, pt, &ekf, &ept, obs, &w, err, q, t, Xw, obs2, inv_sigma2, K4, only_pose);
    ba_compute_error(&B, 0);
    err2[0] = err[0]; err2[1] = err[1];
    if (chi2) *chi2 = ba_chi2(&B, 0);
    if (depth_ok) *depth_ok = ba_depth_positive(&B, 0);
    double Jl[6] = {0, 0, 0, 0, 0, 0};
    ba_edge_jacobians(&B, 0, Jpose12, Jl);
    if (Jpoint6) memcpy(Jpoint6, Jl, sizeof Jl);
}
void oracle_probe_huber(double e2, double delta, double *rho3) { ba_t B; memset(&B, 0, sizeof B); B.delta = delta; B.dsqr = (double)(float)(delta * delta); huber(&B, e2, rho3); }
static void probe_get_sim3(const double *s8, sim3 *S) { memcpy(S->q, s8, 4 * sizeof(double)); memcpy(S->t, s8 + 4, 3 * sizeof(double)); S->s = s8[7]; }
static void probe_put_sim3(const sim3 *S, double *s8) { memcpy(s8, S->q, 4 * sizeof(double)); memcpy(s8 + 4, S->t, 3 * sizeof(double)); s8[7] = S->s; }
void oracle_probe_sim3_exp(const double *u7, double *s8) { sim3 S; sim3_exp(u7, &S); probe_put_sim3(&S, s8); }
void oracle_probe_sim3_log(const double *s8, double *u7) { sim3 S; probe_get_sim3(s8, &S); sim3_log(&S, u7); }
void oracle_probe_sim3_inverse(const double *s8, double *o8) { sim3 S, O; probe_get_sim3(s8, &S); sim3_inverse(&S, &O); probe_put_sim3(&O, o8); }
void oracle_probe_sim3_mul(const double *a8, const double *b8, double *o8) { sim3 A, Bm, O; probe_get_sim3(a8, &A); probe_get_sim3(b8, &Bm); sim3_mul(&A, &Bm, &O); probe_put_sim3(&O, o8); }
void oracle_probe_sim3_map(const double *s8, const double *x3, double *o3) { sim3 S; probe_get_sim3(s8, &S); sim3_map(&S, x3, o3); }
void oracle_probe_sim3_oplus(double *s8, const double *u7, int fix_scale) { sim3 S, O; probe_get_sim3(s8, &S); pg_oplus(&S, u7, fix_scale, &O); probe_put_sim3(&O, s8); }
/* EdgeSim3: error and the numeric Jacobians of both vertices (central differences through oplus, delta = 1e-9), row-major 7x7 */
void oracle_probe_edge_sim3(const double *meas8, const double *si8, const double *sj8, int fix_scale, double *err7, double *Ji49, double *Jj49)
{
    sim3 M, Sv[2];
    probe_get_sim3(meas8, &M); probe_get_sim3(si8, &Sv[0]); probe_get_sim3(sj8, &Sv[1]);
    pg_edge_error(&M, &Sv[0], &Sv[1], err7);
    for (int s = 0; s < 2; s++)
        for (int d = 0; d < 7; d++) {
            double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[7], em[7];
            sim3 Pp;
            add[d] = 1e-9; pg_oplus(&Sv[s], add, fix_scale, &Pp); pg_edge_error(&M, s == 0 ? &Pp : &Sv[0], s == 1 ? &Pp : &Sv[1], ep);
            add[d] = -1e-9; pg_oplus(&Sv[s], add, fix_scale, &Pp); pg_edge_error(&M, s == 0 ? &Pp : &Sv[0], s == 1 ? &Pp : &Sv[1], em);
            for (int r = 0; r < 7; r++) (s == 0 ? Ji49 : Jj49)[7 * r + d] = (1.0 / (2 * 1e-9)) * (ep[r] - em[r]);
        }
}
/* EdgeSim3ProjectXYZ (inverse = 0) / EdgeInverseSim3ProjectXYZ (inverse = 1) of OptimizeSim3: error and the numeric Jacobian wrt the Sim3 vertex (2x7) */
void oracle_probe_edge_sim3_project(int inverse, const double *s8, const double *K1, const double *K2, const float *X3, const float *obs, int fix_scale,
                                    double *err2, double *Jsim14)
{
    s3_t B; memset(&B, 0, sizeof B);
    B.N = 1; B.P1c = X3; B.P2c = X3; B.obs1 = obs; B.obs2 = obs; B.fix_scale = fix_scale;
    memcpy(B.K1, K1, sizeof B.K1); memcpy(B.K2, K2, sizeof B.K2);
    sim3 S; probe_get_sim3(s8, &S);
    double e[4];
    s3_errors(&B, &S, 0, e);
    err2[0] = e[2 * inverse]; err2[1] = e[2 * inverse + 1];
    for (int d = 0; d < 7; d++) {
        double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[4], em[4];
        sim3 Sp;
        add[d] = 1e-9; s3_oplus(&B, &S, add, &Sp); s3_errors(&B, &Sp, 0, ep);
        add[d] = -1e-9; s3_oplus(&B, &S, add, &Sp); s3_errors(&B, &Sp, 0, em);
        for (int r = 0; r < 2; r++) Jsim14[7 * r + d] = (1.0 / (2 * 1e-9)) * (ep[2 * inverse + r] - em[2 * inverse + r]);
    }
}
