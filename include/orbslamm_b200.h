/*
 * orbslamm_b200.h -- C-ABI of the B200-native ORBSLAMM hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * The C++ classes in orbslamm_b200/host/ (ORBextractor / ORBmatcher / Optimizer,
 * same signatures as the reference headers) are thin marshalling shims over these
 * entry points.  Every function returns an int status: 0 = OK, <0 = error
 * (see ORBS_E_*); no exception crosses this boundary.  There is NO CPU fallback:
 * if no CUDA device is usable every entry point returns ORBS_E_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference checkout,
 * S/ = SingleRobotScenario/):
 *   orbx_*     S/include/ORBextractor.h:45-111   (ctor, operator(), getters, mvImagePyramid)
 *   orbm_*     S/include/ORBmatcher.h:37-102     (every member: DescriptorDistance, SearchByProjection x4, SearchByBoW x2,
 *                                                 SearchForInitialization, SearchForTriangulation, SearchBySim3, Fuse x2)
 *              S/src/Frame.cc:230-245,269-325,327-392,404-434  (AssignFeaturesToGrid, isInFrustum, GetFeaturesInArea / PosInGrid, UndistortKeyPoints)
 *   orbv_*     S/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1262  (transform, as called by Frame::ComputeBoW, S/src/Frame.cc:395-402)
 *   orbo_*     S/include/Optimizer.h:37-68       (every member: PoseOptimization, LocalBundleAdjustment, BundleAdjustment, OptimizeSim3,
 *                                                 OptimizeEssentialGraph; M/: MMGlobalBundleAdjustemnt, MMOptimizeEssentialGraph)
 *              S/include/Sim3Solver.h:33-129     (CheckInliers / Project / FromCameraToImage of the RANSAC)
 *   orbf_*     the fused per-frame front end (extract -> project -> search -> pose optimisation) used by the e2e path
 */
#ifndef ORBSLAMM_B200_H
#define ORBSLAMM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBS_OK 0
#define ORBS_E_INVALID (-1)   /* bad argument (null pointer, non-positive size, ...) */
#define ORBS_E_CUDA (-2)      /* CUDA runtime error / no device (orbs_last_error() has the text) */
#define ORBS_E_CAPACITY (-3)  /* caller-provided output capacity too small */
#define ORBS_E_SHAPE (-4)     /* image shape the reference algorithm is undefined for */
#define ORBS_E_STATE (-5)     /* call order violated (e.g. download before extract) */

#define ORBS_MAX_LEVELS 16
#define ORBS_FRAME_GRID_COLS 64   /* S/include/Frame.h:37-38 */
#define ORBS_FRAME_GRID_ROWS 48

/* thread-local text of the last error on this thread ("" if none) */
const char *orbs_last_error(void);
/* library/ABI version and the device the library would use; both cheap, no compute */
int orbs_version(void);
int orbs_device_count(void);

/* ------------------------------------------------------------------ */
/* ORB extractor  (replaces S/src/ORBextractor.cc)                       */
typedef struct orbx_handle orbx_handle;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST),
 * ORBextractor.cc:410-470.  `device` = CUDA ordinal.  Thresholds must be >= 1. */
int orbx_create(orbx_handle **out, int nfeatures, float scale_factor, int nlevels,
                int ini_th_fast, int min_th_fast, int device);
int orbx_destroy(orbx_handle *h);

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (ORBextractor.h:63-85) + mnFeaturesPerLevel.  Any
 * pointer may be NULL.  Arrays must hold nlevels entries. */
int orbx_get_tables(const orbx_handle *h, int *nlevels, float *scale_factor, float *scale,
                    float *inv_scale, float *sigma2, float *inv_sigma2, int *features_per_level);

/* Upper bound of keypoints one frame of this shape can return (sum over levels of
 * max(N_l + 3, 4 * nIni_l)); use it to size the output arrays. */
int orbx_max_keypoints(orbx_handle *h, int width, int height, int *max_kp);

/* ORBextractor::operator()(image, mask, keypoints, descriptors), ORBextractor.cc:1043-1105,
 * for n_frames independent images of one shape.  HOST buffers in, HOST buffers out:
 *   images      n_frames images, image f at images + f*frame_stride, rows `stride` bytes apart
 *   kp_* / desc per-frame slabs of `cap` entries: frame f writes [f*cap, f*cap + counts[f])
 *               kp_xy f32[.,2] (level-0 pixels), kp_angle f32 deg, kp_response f32,
 *               kp_octave i32, kp_size f32, desc u8[.,32]   (cv::KeyPoint fields; class_id = -1)
 *   counts      i32[n_frames]
 * Empty image (width or height 0) -> counts = 0, status OK (ORBextractor.cc:1046-1047). */
int orbx_extract(orbx_handle *h, const uint8_t *images, int n_frames, int width, int height,
                 int stride, size_t frame_stride, float *kp_xy, float *kp_angle, float *kp_response,
                 int32_t *kp_octave, float *kp_size, uint8_t *desc, int cap, int32_t *counts);

/* Same, but the images are already in DEVICE memory and the results stay on the device
 * inside the handle (see orbx_device_results).  Asynchronous on the handle's stream.
 * Any base pointer / stride works.  With a 4-byte (resp. 16-byte) aligned base, stride and frame_stride the rows are read with word loads
 * (resp. by the TMA engine) up to the aligned end of the row: every row, the last one included, must then be backed by `stride` bytes
 * (a buffer of n_frames * frame_stride bytes with frame_stride >= height * stride, as cudaMallocPitch / a padded tensor gives). */
int orbx_extract_device(orbx_handle *h, const uint8_t *d_images, int n_frames, int width,
                        int height, int stride, size_t frame_stride);

/* HOST images in (uploaded in chunks, overlapped with the extraction), results stay on the device like
 * orbx_extract_device.  Asynchronous: the host image buffer must stay valid until orbx_synchronize / orbx_download. */
int orbx_extract_host_async(orbx_handle *h, const uint8_t *images, int n_frames, int width, int height,
                            int stride, size_t frame_stride);

/* Device-resident results of the last extract call.  slab = per-frame capacity (entries). */
typedef struct {
    int n_frames, slab;
    const float *kp_xy;        /* [n_frames*slab, 2] */
    const float *kp_angle;     /* [n_frames*slab] */
    const float *kp_response;  /* [n_frames*slab] */
    const int32_t *kp_octave;  /* [n_frames*slab] */
    const float *kp_size;      /* [n_frames*slab] */
    const uint8_t *desc;       /* [n_frames*slab, 32] */
    const int32_t *counts;     /* [n_frames] */
    const int32_t *level_counts; /* [n_frames, nlevels] */
} orbx_device_view;
int orbx_device_results(orbx_handle *h, orbx_device_view *view);

/* Copy the device-resident results of the last orbx_extract_device to host slabs
 * (same layout as orbx_extract) and synchronise. */
int orbx_download(orbx_handle *h, float *kp_xy, float *kp_angle, float *kp_response,
                  int32_t *kp_octave, float *kp_size, uint8_t *desc, int cap, int32_t *counts);

/* mvImagePyramid[level] of frame `frame` of the last call (ORBextractor.h:87,
 * ORBextractor.cc:1107-1132).  border = 0: the level itself (lw x lh);
 * border = 1: with the 19-px BORDER_REFLECT_101 frame ((lw+38) x (lh+38)).
 * Blocks until the data is on the host. */
int orbx_level_size(orbx_handle *h, int width, int height, int level, int *lw, int *lh);
int orbx_get_pyramid_level(orbx_handle *h, int frame, int level, int border, uint8_t *out,
                           int out_stride);

/* Stage inspection (tests): the FAST candidates of one level of one frame of the last call, i.e.
 * vToDistributeKeys of ORBextractor.cc:775-829 as (x, y, response) int triplets relative to minBorder,
 * in no particular order.  xys may be NULL to query the count. */
int orbx_get_candidates(orbx_handle *h, int frame, int level, int32_t *xys, int cap, int *n_out);

/* The CUDA stream (cudaStream_t) this handle launches on, for event timing. */
void *orbx_stream(orbx_handle *h);
/* Launch on a caller-owned stream instead (so that several handles can be chained without events). */
int orbx_set_stream(orbx_handle *h, void *cuda_stream);
int orbx_synchronize(orbx_handle *h);
/* number of kernels this handle has launched since creation (bench bookkeeping) */
long long orbx_kernel_launches(const orbx_handle *h);
/* Per-kernel CUDA-event timing on the handle's stream (bench roofline).  Kernel ids: 0 resize_level,
 * 1 fast_cells, 2 blur7, 3 octree, 4 orient_describe.  set_profiling(1) resets the accumulators. */
int orbx_set_profiling(orbx_handle *h, int enabled);
int orbx_get_kernel_times(orbx_handle *h, double *total_ms, long long *counts, int n);

/* ------------------------------------------------------------------ */
/* ORB matcher  (replaces S/src/ORBmatcher.cc + the Frame grid it searches) */
typedef struct orbm_handle orbm_handle;

#define ORBS_MEM_HOST 0     /* the array arguments are host pointers (copied in/out, call blocks) */
#define ORBS_MEM_DEVICE 1   /* the array arguments are device pointers (asynchronous on the handle's stream) */
#define ORBM_TH_HIGH 100    /* ORBmatcher::TH_HIGH / TH_LOW / HISTO_LENGTH, ORBmatcher.cc:37-39 */
#define ORBM_TH_LOW 50
#define ORBM_HISTO_LENGTH 30

int orbm_create(orbm_handle **out, int device);
int orbm_destroy(orbm_handle *h);
void *orbm_stream(orbm_handle *h);
int orbm_set_stream(orbm_handle *h, void *cuda_stream);
int orbm_synchronize(orbm_handle *h);
long long orbm_kernel_launches(const orbm_handle *h);

/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:1649-1665) for all pairs: out[i*m + j] = Hamming(a[i], b[j]).
 * a u8[n,32], b u8[m,32], out i32[n,m]. */
int orbm_descriptor_distance(orbm_handle *h, const uint8_t *a, int n, const uint8_t *b, int m, int32_t *out, int memspace);

/* MapPoint::ComputeDistinctiveDescriptors (S/src/MapPoint.cc:242-307) for a batch of map points: point p owns the descriptor rows start[p] .. start[p+1]-1 of
 * desc (its observations in std::map<KeyFrame*, size_t> order, bad keyframes left out); best_idx[p] = the row (relative to start[p]) with the least median
 * Hamming distance to the point's rows (sorted row, element (int)(0.5 * (N - 1)), self distance 0 included; first row on ties), -1 for a point without rows.
 * The reference calls the member once per map point from LocalMapping::ProcessNewKeyFrame / CreateNewMapPoints / SearchInNeighbors, LoopClosing and
 * MultiMapper (O(N^2) DescriptorDistance calls each); the drop-in collects the points of one such loop and makes one call.
 *   desc u8[start[n_points], 32] (16-byte aligned), start i32[n_points + 1], best_idx i32[n_points]. */
int orbm_distinctive_descriptors(orbm_handle *h, int n_points, const uint8_t *desc, const int32_t *start, int32_t *best_idx, int memspace);

/* Frame::UndistortKeyPoints (S/src/Frame.cc:404-434): cv::undistortPoints(pts, K, distCoef, R = I, P = K) with OpenCV's default
 * criteria (5 fixed-point iterations in fp64); a copy when dist5[0] == 0 (Frame.cc:406-410).
 *   kp_xy f32[n_frames*slab,2] (Frame::mvKeys[i].pt), counts i32[n_frames]; K4 f32[4] = fx fy cx cy and dist5 f32[5] = k1 k2 p1 p2 k3
 *   are HOST pointers (k3 = 0 for the 4-entry DistCoef of the reference's YAML files); kp_xy_un out (Frame::mvKeysUn[i].pt). */
int orbm_undistort_keypoints(orbm_handle *h, int n_frames, const float *kp_xy, const int32_t *counts, int slab, const float *K4,
                             const float *dist5, float *kp_xy_un, int memspace);

/* Frame::isInFrustum (S/src/Frame.cc:269-325) with MapPoint::PredictScale / Get{Min,Max}DistanceInvariance
 * (S/src/MapPoint.cc:373-394) for counts[f] map points per frame: positive depth, projection inside the image bounds, distance
 * inside [0.8 mfMinDistance, 1.2 mfMaxDistance], viewing cosine >= limit (0.5 in Tracking::SearchLocalPoints, Tracking.cc:1233).
 *   Tcw f32[n_frames,16]; Ow f32[n_frames,3] = Frame::mOw; K4 / bounds4 (mnMinX mnMinY mnMaxX mnMaxY) HOST pointers;
 *   Xw, normal f32[n_frames*slab,3]; mf_min_distance / mf_max_distance f32[.] = MapPoint::mfMinDistance / mfMaxDistance.
 *   out: in_view u8[.] (mbTrackInView), proj_xy f32[.,2] (mTrackProjX/Y), pred_level i32[.] (mnTrackScaleLevel), view_cos f32[.]
 *   (mTrackViewCos) -- written for points in view; these feed orbm_search_by_projection for the local-map search. */
int orbm_is_in_frustum(orbm_handle *h, int n_frames, const float *Tcw, const float *Ow, const float *K4, const float *bounds4,
                       float log_scale_factor, float viewing_cos_limit, const float *Xw, const float *normal,
                       const float *mf_min_distance, const float *mf_max_distance, const int32_t *counts, int slab,
                       uint8_t *in_view, float *proj_xy, int32_t *pred_level, float *view_cos, int memspace);

/* Frame::AssignFeaturesToGrid (S/src/Frame.cc:230-245, PosInGrid :382-392): the 64x48 feature grid of n_frames frames as CSR,
 *   cell = ix * 48 + iy; cell_start i32[n_frames, 64*48 + 1]; cell_items i32[n_frames, f_slab] feature indices, ascending inside a cell
 *   (the order of the reference's mGrid[ix][iy] vectors).  Features that fall outside the grid are in no cell. */
int orbm_assign_features_to_grid(orbm_handle *h, int n_frames, const float *bounds4, const float *f_xy, const int32_t *f_counts, int f_slab,
                                 int32_t *cell_start, int32_t *cell_items, int memspace);

/* Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (S/src/Frame.cc:327-380) / KeyFrame::GetFeaturesInArea(x, y, r) (S/src/KeyFrame.cc:
 * 618-657; pass win_origin2 = the keyframe's integer mnMinX, mnMinY and no level arrays) for q_counts[f] queries per frame.
 *   q_xyr f32[n_frames*q_slab,3] = x, y, r; q_minl / q_maxl i32[.] (may be NULL = -1, -1); f_octave may be NULL without level limits.
 *   out_idx i32[n_frames*q_slab, cap]: the indices in the reference's order (cells ix-major, ascending index inside a cell), truncated
 *   at cap; out_count i32[.]: the full size of the returned vector. */
int orbm_get_features_in_area(orbm_handle *h, int n_frames, const float *bounds4, const float *win_origin2, const float *f_xy, const int32_t *f_octave,
                              const int32_t *f_counts, int f_slab, const float *q_xyr, const int32_t *q_minl, const int32_t *q_maxl,
                              const int32_t *q_counts, int q_slab, int cap, int32_t *out_idx, int32_t *out_count, int memspace);

/* Projection block of SearchByProjection(CurrentFrame, LastFrame, th, bMono=true), ORBmatcher.cc:1336-1391:
 * per frame f and last-frame slot i (valid[i] != 0: the slot holds a non-outlier map point) project Xw with the
 * current pose, keep it if the depth is positive and (u, v) lies inside the image bounds, and emit the search
 * window.  fp32 arithmetic restated from OpenCV's small gemm (sequential fp32, no FMA).
 *   K4 f32[4] = fx fy cx cy and bounds4 f32[4] = mnMinX mnMinY mnMaxX mnMaxY are always HOST pointers;
 *   Tcw f32[n_frames,16] row-major;
 *   scale_factors f32[nlevels]; Xw f32[n_frames*q_slab,3]; last_octave i32[.]; q_counts i32[n_frames];
 *   q_valid u8[.] in/out; q_uv f32[.,2], q_radius f32[.], q_minl/q_maxl i32[.] out (octave-1 / octave+1). */
int orbm_project_last_frame(orbm_handle *h, int n_frames, const float *Tcw, const float *K4, const float *bounds4,
                            const float *scale_factors, int nlevels, const float *Xw, const int32_t *last_octave,
                            const int32_t *q_counts, int q_slab, float th, uint8_t *q_valid, float *q_uv,
                            float *q_radius, int32_t *q_minl, int32_t *q_maxl, int memspace);

/* The search loop shared by SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:45-129: ratio > 0,
 * check_ori = 0) and SearchByProjection(Cur, Last, th, mono) (ORBmatcher.cc:1353-1467: ratio <= 0, check_ori = 1),
 * for n_frames independent frames.  Per frame the 64x48 grid of Frame::AssignFeaturesToGrid is rebuilt on the
 * device, every valid query visits Frame::GetFeaturesInArea(u, v, r, minLevel, maxLevel), skips features that
 * already hold a map point, takes the best (and second best) Hamming distance, accepts it if best <= th_dist
 * (and, with ratio > 0, unless both are on the same level and best > ratio * second), and claims the feature.
 * The reference's loop is sequential (an accepted query hides its feature from all later queries); the result
 * here is identical to that sequential order.  With check_ori the 30-bin rotation histogram keeps the three
 * dominant bins (ComputeThreeMaxima, ORBmatcher.cc:1603-1644).
 *   features: f_xy f32[n_frames*f_slab,2] (mvKeysUn.pt), f_octave i32, f_angle f32, f_desc u8[.,32], f_counts i32[n_frames]
 *   queries:  q_valid u8[n_frames*q_slab], q_uv f32[.,2], q_radius f32, q_minl/q_maxl i32, q_angle f32, q_desc u8[.,32],
 *             q_counts i32[n_frames]
 *   feat_match i32[n_frames*f_slab] in/out: in  < 0 = free, >= 0 = feature already holds a map point (skipped);
 *                                            out = index of the query now assigned to the feature (Frame::mvpMapPoints)
 *   nmatches i32[n_frames] out (the function's return value per frame). */
int orbm_search_by_projection(orbm_handle *h, int n_frames, const float *bounds4,
                              const float *f_xy, const int32_t *f_octave, const float *f_angle, const uint8_t *f_desc,
                              const int32_t *f_counts, int f_slab,
                              const uint8_t *q_valid, const float *q_uv, const float *q_radius, const int32_t *q_minl,
                              const int32_t *q_maxl, const float *q_angle, const uint8_t *q_desc, const int32_t *q_counts,
                              int q_slab, int th_dist, float ratio, int check_ori,
                              int32_t *feat_match, int32_t *nmatches, int memspace);

/* ---- KeyFrame / Sim3 projection family (ORBmatcher.cc:292-405, 827-977, 979-1102, 1104-1328, 1474-1601) ----------------------
 * These five members share one shape: transform every candidate map point into the target view, run the visibility tests,
 * predict the pyramid level (MapPoint::PredictScale, MapPoint.cc:385-394), search the target's feature grid in a radius
 * th * mvScaleFactors[level] and keep the best Hamming distance.  They differ in which tests they make; `flags` selects them. */
#define ORBM_PROJ_TWO_STEP     0x01 /* x = R2 (R X + t) + t2   (SearchBySim3: R1w,t1w then sR21,t21; :1155-1156, :1235-1236) */
#define ORBM_PROJ_NO_DEPTH     0x02 /* no "z < 0 -> skip" test (SearchByProjection(Frame&, KeyFrame*, ...), :1503-1511) */
#define ORBM_PROJ_FRAME_BOUNDS 0x04 /* reject u < min || u > max (Frame, :1513-1516); default is KeyFrame::IsInImage
                                       (x >= min && x < max, KeyFrame.cc:659-662) */
#define ORBM_PROJ_FRAME_UV     0x08 /* u = fx*xc*invz + cx (:1510-1511); default x = xc*invz, u = fx*x + cx (:334-338) */
#define ORBM_PROJ_DIST_CAMERA  0x10 /* distance = |x| in the target camera (SearchBySim3, :1179); default |X - Ow| */
#define ORBM_PROJ_CHECK_NORMAL 0x20 /* reject PO.dot(Pn) < 0.5*dist (:355-358, :886-889, :1037-1040) */
#define ORBM_PROJ_LEVEL_PLUS1  0x40 /* levels [l-1, l+1] (GetFeaturesInArea(u,v,r,l-1,l+1), :1532); default [l-1, l] */

typedef struct orbm_projection {    /* one per target view; always a HOST array */
    float R[9], t[3];               /* Rcw, tcw of the target (Scw's rotation / scale and translation / scale for the Sim3 variants) */
    float R2[9], t2[3];             /* second transform (ORBM_PROJ_TWO_STEP) */
    float Ow[3];                    /* camera centre (unused with ORBM_PROJ_DIST_CAMERA) */
    float fx, fy, cx, cy;
    float min_x, min_y, max_x, max_y; /* image bounds of the target: KeyFrame::mnMinX.. (ints) or Frame::mnMinX.. */
    float log_scale_factor;         /* mfLogScaleFactor of the target */
    float th;                       /* radius = th * scale_factors[level] */
    int32_t flags;
} orbm_projection;

/* Projection half.  q_valid in: the map point is a candidate (not bad, not already found, ...); out: and it passed every test.
 *   Xw f32[n_views*q_slab,3]; normal f32[.,3] (may be NULL without ORBM_PROJ_CHECK_NORMAL); mf_min_distance / mf_max_distance f32[.]
 *   = MapPoint::mfMinDistance / mfMaxDistance (the 0.8 / 1.2 factors of Get{Min,Max}DistanceInvariance are applied inside);
 *   out: q_uv f32[.,2], q_radius f32[.], q_minl / q_maxl i32[.], q_level i32[.] (may be NULL).
 * A predicted level outside [0, nlevels) makes the reference index mvScaleFactors out of range (undefined behaviour, PredictScale
 * is not clamped in this version); such points are dropped here. */
int orbm_project_points(orbm_handle *h, int n_views, const orbm_projection *proj, const float *scale_factors, int nlevels,
                        const float *Xw, const float *normal, const float *mf_min_distance, const float *mf_max_distance,
                        const int32_t *q_counts, int q_slab, uint8_t *q_valid, float *q_uv, float *q_radius, int32_t *q_minl,
                        int32_t *q_maxl, int32_t *q_level, int memspace);

/* orbm_search_by_projection for a KeyFrame target: the feature grid was built by the Frame (float bounds grid_bounds4, Frame.cc:
 * 230-245) but KeyFrame::GetFeaturesInArea offsets by the KeyFrame's INTEGER mnMinX / mnMinY (KeyFrame.h, KeyFrame.cc:618-657);
 * win_origin2 = those two values (NULL = same as the grid).  Everything else as orbm_search_by_projection. */
int orbm_search_by_projection_kf(orbm_handle *h, int n_frames, const float *grid_bounds4, const float *win_origin2,
                                 const float *f_xy, const int32_t *f_octave, const float *f_angle, const uint8_t *f_desc,
                                 const int32_t *f_counts, int f_slab,
                                 const uint8_t *q_valid, const float *q_uv, const float *q_radius, const int32_t *q_minl,
                                 const int32_t *q_maxl, const float *q_angle, const uint8_t *q_desc, const int32_t *q_counts,
                                 int q_slab, int th_dist, float ratio, int check_ori,
                                 int32_t *feat_match, int32_t *nmatches, int memspace);

/* Search half without claims (Fuse :892-938 and :1052-1078, SearchBySim3 :1193-1224 / :1273-1304): every valid query takes the
 * feature with the smallest Hamming distance in its window (first visited wins ties), independently of all other queries.
 *   inv_level_sigma2 f32[nlevels] (HOST, may be NULL): with it a candidate is skipped when
 *   ((u-kpx)^2 + (v-kpy)^2) * inv_level_sigma2[octave] > chi2_gate   (Fuse, monocular branch :918-928, chi2_gate = 5.99)
 *   out: q_best_idx i32[.] = feature index or -1 when the window is empty / best distance > th_dist; q_best_dist i32[.]. */
int orbm_search_best_in_window(orbm_handle *h, int n_frames, const float *grid_bounds4, const float *win_origin2,
                               const float *f_xy, const int32_t *f_octave, const uint8_t *f_desc, const int32_t *f_counts, int f_slab,
                               const uint8_t *q_valid, const float *q_uv, const float *q_radius, const int32_t *q_minl,
                               const int32_t *q_maxl, const uint8_t *q_desc, const int32_t *q_counts, int q_slab, int th_dist,
                               const float *inv_level_sigma2, int nlevels, float chi2_gate,
                               int32_t *q_best_idx, int32_t *q_best_dist, int memspace);

/* ---- vocabulary-bucket matchers: SearchByBoW(KeyFrame*, Frame&, ..) ORBmatcher.cc:159-290, SearchByBoW(KeyFrame*, KeyFrame*, ..) :524-657,
 * SearchForTriangulation :659-825 ---------------------------------------------------------------------------------------------------------
 * A DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>, Frame::mFeatVec / KeyFrame::mFeatVec) is passed as CSR per keyframe:
 *   fv_nodes i32[n_pairs, fv_slab] ascending node ids, fv_start i32[n_pairs, fv_slab + 1] offsets into fv_items, fv_items i32[n_pairs, slab]
 *   feature indices in the order the map's vectors hold them, fv_counts i32[n_pairs] number of nodes.
 * Side 1 is the keyframe whose features are walked (the reference's outer loop), side 2 the one searched.
 *   ORBM_BOW_MATCH: elig1[i] / elig2[i] = the feature takes part (side 1: holds a good map point; side 2: all features for the Frame variant,
 *     holds a good map point for the KeyFrame variant).  Best and second best Hamming distance over the still unclaimed side-2 features of
 *     the node; accepted if best <= TH_LOW and (float)best < ratio * (float)second; the accepted feature is claimed.
 *   ORBM_BOW_TRIANGULATION (monocular keyframes, bOnlyStereo = false): elig = the feature holds NO map point; no claims (the reference never
 *     sets vbMatched2); smallest distance <= TH_LOW among the candidates that pass the epipole-distance gate (:747-753) and
 *     CheckDistEpipolarLine (:140-157), the last one winning ties (:742).  Needs `epi`.
 * With check_ori the rotation histogram keeps the three dominant bins.  Out: match12 i32[n_pairs*slab1] = side-2 feature index or -1,
 * nmatches i32[n_pairs]. */
#define ORBM_BOW_MATCH 0
#define ORBM_BOW_TRIANGULATION 1
typedef struct orbm_epipolar {       /* array members follow `memspace`; scale_factors2 / level_sigma2_2 are HOST arrays of nlevels floats */
    const float *xy1, *xy2;          /* mvKeysUn.pt of side 1 [n_pairs*slab1,2] / side 2 [n_pairs*slab2,2] */
    const int32_t *octave2;          /* [n_pairs*slab2] */
    const float *F12;                /* [n_pairs,9] row-major fundamental matrix (LocalMapping::ComputeF12) */
    const float *epipole;            /* [n_pairs,2] ex, ey: camera centre of keyframe 1 projected into keyframe 2 (:665-673) */
    const float *scale_factors2, *level_sigma2_2;
    int32_t nlevels;
} orbm_epipolar;
int orbm_search_by_bow(orbm_handle *h, int n_pairs, int mode,
                       const uint8_t *desc1, const float *angle1, const uint8_t *elig1, const int32_t *counts1, int slab1,
                       const int32_t *fv1_nodes, const int32_t *fv1_start, const int32_t *fv1_items, const int32_t *fv1_counts, int fv1_slab,
                       const uint8_t *desc2, const float *angle2, const uint8_t *elig2, const int32_t *counts2, int slab2,
                       const int32_t *fv2_nodes, const int32_t *fv2_start, const int32_t *fv2_items, const int32_t *fv2_counts, int fv2_slab,
                       float ratio, int check_ori, const orbm_epipolar *epi, int32_t *match12, int32_t *nmatches, int memspace);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:407-522) for n_pairs frame pairs: the
 * level-0 keypoints of F1 search GetFeaturesInArea(prev.x, prev.y, windowSize, 0, 0) of F2; best / second best with the ratio test; a feature
 * of F2 already matched with a smaller-or-equal distance is skipped, a better match steals it (:471-476); rotation histogram (:497-512).
 *   prev_matched f32[n_pairs*slab1,2] in/out (updated with the matched keypoint positions, :517-519); matches12 i32[n_pairs*slab1] out. */
int orbm_search_for_initialization(orbm_handle *h, int n_pairs, const float *bounds4,
                                   const int32_t *octave1, const float *angle1, const uint8_t *desc1, const int32_t *counts1, int slab1,
                                   const float *xy2, const int32_t *octave2, const float *angle2, const uint8_t *desc2, const int32_t *counts2, int slab2,
                                   float *prev_matched, int window, float ratio, int check_ori, int32_t *matches12, int32_t *nmatches, int memspace);

/* ------------------------------------------------------------------ */
/* DBoW2 vocabulary transform  (replaces S/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1262 transform, BowVector.cpp:34-84,
 * FORB.cpp:81-101, as called by Frame::ComputeBoW / KeyFrame::ComputeBoW, S/src/Frame.cc:395-402: transform(descs, mBowVec, mFeatVec, 4)) */
typedef struct orbv_handle orbv_handle;

/* The vocabulary tree as flat HOST arrays (copied to HBM once): node 0 is the root; node_desc u8[n_nodes,32]; the children of node i are
 * child_ids[child_start[i] .. child_start[i+1]) in the order of Node::children (file order of ORBvoc.txt, TemplatedVocabulary.h:1370-1419);
 * a node without children is a word with id word_id[i] and weight weight[i] (the idf part; TF_IDF weighting, L1 scoring as ORBSLAMM
 * loads it).  k, L as in the file header. */
int orbv_create(orbv_handle **out, int device, int k, int L, int n_nodes, const uint8_t *node_desc, const int32_t *child_start,
                const int32_t *child_ids, const int32_t *word_id, const double *weight);
int orbv_destroy(orbv_handle *h);
long long orbv_kernel_launches(const orbv_handle *h);

/* transform(features, v, fv, levelsup) for n_frames frames: every descriptor walks the tree (at each level the child with the smallest
 * Hamming distance, the first one on ties); words with weight 0 are stopped and contribute to neither vector.
 *   desc u8[n_frames*slab,32], counts i32[n_frames];
 *   out: word_of_feature / node_of_feature i32[n_frames*slab] (may be NULL; -1 = stopped);
 *        BowVector: bow_ids i32[n_frames*slab] ascending word ids, bow_vals f64[.] (sum of the word weight over its features in feature
 *        order, then divided by the L1 norm taken in ascending word order -- the std::map loops of the reference), bow_counts i32[n_frames];
 *        FeatureVector as the CSR orbm_search_by_bow takes: fv_nodes i32[n_frames*slab], fv_start i32[n_frames*(slab+1)],
 *        fv_items i32[n_frames*slab] (feature indices, ascending inside a node), fv_counts i32[n_frames]. */
int orbv_transform(orbv_handle *h, int n_frames, const uint8_t *desc, const int32_t *counts, int slab, int levelsup,
                   int32_t *word_of_feature, int32_t *node_of_feature, int32_t *bow_ids, double *bow_vals, int32_t *bow_counts,
                   int32_t *fv_nodes, int32_t *fv_start, int32_t *fv_items, int32_t *fv_counts, int memspace);

/* ------------------------------------------------------------------ */
/* Optimizer  (replaces S/src/Optimizer.cc and the g2o LM / Schur / LDLT stack it drives; all fp64 inside,
 * float32 poses and points at the boundary like cv::Mat / Converter.cc)                                      */
typedef struct orbo_handle orbo_handle;

int orbo_create(orbo_handle **out, int device);
int orbo_destroy(orbo_handle *h);
void *orbo_stream(orbo_handle *h);
int orbo_set_stream(orbo_handle *h, void *cuda_stream);
int orbo_synchronize(orbo_handle *h);
long long orbo_kernel_launches(const orbo_handle *h);

/* Optimizer::PoseOptimization(Frame*), Optimizer.cc:262-474 (monocular edges), for n_frames independent frames:
 * 4 rounds x 10 LM iterations on EdgeSE3ProjectXYZOnlyPose edges, Huber sqrt(5.991), chi2 gating after each
 * round, pose reset to the initial estimate at the start of every round.
 *   Tcw f32[n_frames,16] in/out (row-major 4x4); K4 f32[4] HOST pointer (fx fy cx cy);
 *   per frame slab of `slab` correspondences, counts i32[n_frames]:
 *   Xw f32[.,3] (MapPoint::GetWorldPos), obs f32[.,2] (mvKeysUn.pt), inv_sigma2 f32[.] (mvInvLevelSigma2[octave]);
 *   outlier u8[.] out (Frame::mvbOutlier); n_inliers i32[n_frames] out (the return value; 0 if < 3 correspondences). */
int orbo_pose_optimization(orbo_handle *h, int n_frames, float *Tcw, const float *K4, const float *Xw, const float *obs,
                           const float *inv_sigma2, const int32_t *counts, int slab, uint8_t *outlier, int32_t *n_inliers,
                           int memspace);

/* Same optimisation, but the correspondences are gathered on the device from a matcher result exactly like the edge
 * construction loop of Optimizer.cc:303-384: feature i of frame f contributes an edge iff feat_match[i] >= 0 (it holds
 * map point / query feat_match[i]), with obs = f_xy[i], Xw = q_Xw[feat_match[i]], information = inv_level_sigma2[octave].
 *   f_outlier u8[n_frames*f_slab] out (Frame::mvbOutlier, 0 for features without a map point); n_inliers i32[n_frames];
 *   n_edges i32[n_frames] optional out (nInitialCorrespondences).  K4 and inv_level_sigma2... K4 is a HOST pointer. */
int orbo_pose_optimization_matched(orbo_handle *h, int n_frames, float *Tcw, const float *K4, const float *f_xy,
                                   const int32_t *f_octave, const int32_t *f_counts, int f_slab, const int32_t *feat_match,
                                   const float *q_Xw, const int32_t *q_counts, int q_slab, const float *inv_level_sigma2,
                                   int nlevels, uint8_t *f_outlier, int32_t *n_inliers, int32_t *n_edges, int memspace);

/* Optimizer::LocalBundleAdjustment (Optimizer.cc:476-801; two_stage = 1: its0 robust iterations, chi2 / depth gating,
 * its1 non-robust iterations) and Optimizer::BundleAdjustment (Optimizer.cc:68-260; two_stage = 0: its0 iterations
 * with `robust`) over flat arrays (HOST pointers):
 *   poses f32[K,16] in/out; fixed u8[K]: 0 free, 1 fixed but written back (mnId == 0), 2 fixed camera (never written);
 *   intr f64[K,4] fx fy cx cy per keyframe; points f32[P,3] in/out;
 *   edges: e_kf i32[E], e_pt i32[E], e_uv f32[E,2], e_inv_sigma2 f32[E], in ANY order (the per-edge outputs come back in the caller's order);
 *          a list already grouped by ascending e_pt -- what walking the map points produces, Optimizer.cc:595-675 -- skips the stable scatter;
 *   stop_flag: optional host flag, ONE BYTE (a C++ `bool *` such as &mbAbortBA / &mbStopGBA is passed as is), watched by the
 *              calling thread while the LM slots run on the device and seen by the device before every trial decision
 *              (g2o: terminate() per iteration and per trial, sparse_optimizer.h:188, levenberg.cpp:149);
 *   outputs (optional): e_chi2 f64[E], e_depth_ok u8[E], e_outlier u8[E] = the final check of Optimizer.cc:734-766;
 *   stats i32[4]: LM iterations, LM trials, Cholesky failures, 1 if aborted before the first iteration.
 * Returns ORBS_OK, or 1 if the stop flag was already set on entry (nothing touched, Optimizer.cc:678-680). */
int orbo_bundle_adjust(orbo_handle *h, int K, float *poses, const uint8_t *fixed, const double *intr, int P, float *points,
                       int E, const int32_t *e_kf, const int32_t *e_pt, const float *e_uv, const float *e_inv_sigma2,
                       int two_stage, int its0, int its1, int robust, const volatile uint8_t *stop_flag,
                       double *e_chi2, uint8_t *e_depth_ok, uint8_t *e_outlier, int32_t *stats);

/* Numeric core of Optimizer::OptimizeEssentialGraph / MMOptimizeEssentialGraph (S/src/Optimizer.cc:804-1067, 1069-1346): Levenberg over K
 * VertexSim3Expmap vertices and E EdgeSim3 edges (error = log(Sji * Siw * Sjw^-1), identity information, numeric Jacobians with delta = 1e-9 for
 * both vertices), lambda_init = 1e-16 and 20 iterations in the reference; the normal equations go through the tiled sparse Cholesky.
 * HOST pointers.  sim3 f64[K,8] in/out = r (x y z w), t, s of every vertex (CorrectedSim3 or Sim3(Rcw, tcw, 1), :845-869); fixed u8[K] (the loop
 * keyframe, :866-867); edge e joins vertex[0] = e_i[e] and vertex[1] = e_j[e] with measurement e_meas f64[E,8] = Sji (:895-905; the caller builds the
 * loop / spanning-tree / loop-edge / covisibility edge set of :880-1000); fix_scale = bFixScale; lambda_init <= 0 selects g2o's default
 * (1e-5 * max diagonal).  stats i32[3] (may be NULL): LM iterations, LM trials, failed factorisations. */
int orbo_optimize_pose_graph(orbo_handle *h, int K, double *sim3, const uint8_t *fixed, int E, const int32_t *e_i, const int32_t *e_j,
                             const double *e_meas, int fix_scale, int iterations, double lambda_init, int32_t *stats);

/* Sim3Solver::ComputeSim3 (S/src/Sim3Solver.cc:226-338: Horn's closed form on three point pairs) for n_hyp RANSAC min sets in one launch.
 *   X1, X2 f32[n_hyp, 3, 3]: the three points (rows) of every min set in camera-1 / camera-2 coordinates (mvX3Dc1[idx], mvX3Dc2[idx]);
 *   out: T12, T21 f32[n_hyp, 16] (mT12i, mT21i, row-major 4x4), Rts f32[n_hyp, 13] = mR12i (9), mt12i (3), ms12i -- may be NULL.
 * Float arithmetic in OpenCV's evaluation order; cv::eigen is OpenCV-build dependent, so parity is a float tolerance (1e-4 relative on the matrices), not
 * bit-exactness: the drop-in uses this entry only when built with -DORBSLAMM_DEVICE_COMPUTE_SIM3 and otherwise keeps the reference's host ComputeSim3. */
int orbo_sim3_compute(orbo_handle *h, int n_hyp, const float *X1, const float *X2, int fix_scale, float *T12, float *T21, float *Rts, int memspace);

/* Sim3Solver (S/src/Sim3Solver.cc), the data-parallel part of the RANSAC.  ComputeSim3 (Horn's closed form on three points, :226-338) stays with
 * the caller: the random index triples do not depend on the inlier counts, so all hypotheses of a solver (<= mRansacMaxIts = 300) can be
 * generated first and checked in ONE call; the caller then walks the counts in order and applies the rule of iterate() (:176-194) unchanged.
 *   orbo_sim3_prepare: the constructor's per-correspondence data for one keyframe (:84-103, :419-437): max_err i32[N] = (size_t)(9.210 *
 *     mvLevelSigma2[octave]) -- the reference stores the thresholds in std::vector<size_t>, i.e. truncated -- and p2d f32[N,2] =
 *     FromCameraToImage(X3Dc, K).  level_sigma2 f32[nlevels] and K4 are HOST arrays.
 *   orbo_sim3_check_inliers: CheckInliers (:340-365) with Project (:392-417) for n_hyp hypotheses: T12 / T21 f32[n_hyp,16] (mT12i / mT21i,
 *     row-major), X3Dc1 / X3Dc2 f32[N,3] (mvX3Dc1 / 2), P1im1 / P2im2 f32[N,2], K1 / K2 f32[4] HOST.
 *     out: inliers u8[n_hyp,N] (mvbInliersi per hypothesis), n_inliers i32[n_hyp] (mnInliersi). */
int orbo_sim3_prepare(orbo_handle *h, int N, const float *X3Dc, const int32_t *octave, const float *level_sigma2, int nlevels, const float *K4,
                      int32_t *max_err, float *p2d, int memspace);
int orbo_sim3_check_inliers(orbo_handle *h, int n_hyp, const float *T12, const float *T21, int N, const float *X3Dc1, const float *X3Dc2, const float *P1im1,
                            const float *P2im2, const int32_t *max_err1, const int32_t *max_err2, const float *K1, const float *K2, uint8_t *inliers,
                            int32_t *n_inliers, int memspace);

/* Optimizer::OptimizeSim3(pKF1, pKF2, vpMatches1, g2oS12, th2, bFixScale) (S/src/Optimizer.cc:1348-1543) for n_pairs independent keyframe
 * pairs: one VertexSim3Expmap against fixed points, EdgeSim3ProjectXYZ (x1 = S12 X2) + EdgeInverseSim3ProjectXYZ (x2 = S21 X1) per
 * correspondence with Huber kernels (delta = sqrt(th2)), g2o Levenberg on the dense 7x7 system, numeric Jacobians (central differences,
 * delta = 1e-9, base_binary_edge.hpp:131-205): 5 iterations, drop pairs with chi2 > th2 on either edge, 10 (5 if none dropped) more.
 *   sim3 f64[n_pairs,8] in/out = g2o::Sim3 members r (x y z w), t, s of g2oS12 -- left untouched when fewer than 10 correspondences
 *     survive the first pass (the reference returns 0 there, :1497-1498);
 *   valid u8[n_pairs*slab]: slot i is a correspondence (vpMatches1[i] set, both points good, :1398-1433); P1c / P2c f32[.,3]: the two map
 *     points in their own camera frames (R1w*P3D1w+t1w, float, :1411-1424); obs1 / obs2 f32[.,2] = kpUn.pt in KF1 / KF2; inv_sigma2_1/2
 *     f32[.] = mvInvLevelSigma2[octave]; K1 / K2 f32[n_pairs,4] = fx fy cx cy; counts i32[n_pairs];
 *   out: inlier u8[.] (vpMatches1[i] stays set), n_inliers i32[n_pairs] (the return value), lm_stats i32[n_pairs,2] (may be NULL):
 *     LM iterations and trials executed. */
int orbo_optimize_sim3(orbo_handle *h, int n_pairs, double *sim3, const uint8_t *valid, const float *P1c, const float *P2c, const float *obs1,
                       const float *obs2, const float *inv_sigma2_1, const float *inv_sigma2_2, const float *K1, const float *K2, const int32_t *counts,
                       int slab, float th2, int fix_scale, uint8_t *inlier, int32_t *n_inliers, int32_t *lm_stats, int memspace);

/* Multi-GPU bundle adjustment: one process per GPU, keyframe poses replicated, MAP POINTS (with all
 * their observations) sharded over the ranks.  After orbo_comm_init the handle's orbo_bundle_adjust becomes a
 * collective: every rank passes ALL keyframes (same order, same poses) but only ITS points and edges; each rank builds
 * its partial reduced pose system as packed nonzero tiles + right-hand side in ONE buffer, ONE exchange per LM trial sums them over
 * the ranks (reduce-scatter + all-gather kernels over NVLink peer memory, csrc/peer_reduce.cuh; ncclAllReduce where the ranks cannot
 * map each other's memory), every rank factors the same system redundantly and back-substitutes its own points.
 * Five scalars (chi2, the two gain-ratio parts, stop flag, Cholesky failure) are summed after the trial so that all ranks take
 * identical LM decisions on their device-resident control blocks; the keyframe activity mask and the tile adjacency are
 * reduced once per call.  Poses come back
 * identical on every rank; point / edge outputs are the rank's shard.
 *   orbo_comm_unique_id: rank 0 creates the 128-byte NCCL id, the caller distributes it (MPI, torch.distributed, files);
 *   orbo_comm_init: collective. */
int orbo_comm_unique_id(uint8_t *id128);
int orbo_comm_init(orbo_handle *h, int nranks, int rank, const uint8_t *id128);
/* how the sharded solve exchanges the reduced system: 0 = single GPU, 1 = NCCL all-reduce, 2 = NVLink peer-memory kernels (csrc/peer_reduce.cuh; chosen by
 * orbo_comm_init when every rank can map every other rank's buffers through cudaIpc, ORBS_NO_PEER=1 keeps NCCL). */
int orbo_comm_mode(const orbo_handle *h);

/* ------------------------------------------------------------------ */
/* Fused front end: what Tracking::TrackWithMotionModel (S/src/Tracking.cc:912-973) does per frame -- Frame ctor ->
 * ORBextractor::operator(), ORBmatcher::SearchByProjection(Cur, Last, th, mono), Optimizer::PoseOptimization -- for
 * n_frames independent streams in one HOST-buffer call.  The handle borrows the three handles (which keep working on
 * their own) and puts them on one stream; intermediate results stay in HBM.
 *   in : images (as orbx_extract); K4, scale_factors[nlevels], inv_level_sigma2[nlevels] (HOST);
 *        last-frame map points per stream: q_Xw f32[.,3], q_octave, q_angle, q_desc u8[.,32], q_valid u8, q_counts, q_slab;
 *        Tcw f32[n_frames,16]: predicted pose in, optimised pose out; image bounds are (0, 0, width, height).
 *   out: keypoints / descriptors / counts (as orbx_extract, cap >= orbx_max_keypoints), feat_match i32[n_frames*cap]
 *        (index of the last-frame slot matched to each feature or -1), nmatches, outlier u8[n_frames*cap], n_inliers. */
typedef struct orbf_handle orbf_handle;
int orbf_create(orbf_handle **out, orbx_handle *ex, orbm_handle *mt, orbo_handle *po, int device);
int orbf_destroy(orbf_handle *h);
int orbf_track_frames(orbf_handle *h, const uint8_t *images, int n_frames, int width, int height, int stride, size_t frame_stride,
                      const float *K4, const float *scale_factors, const float *inv_level_sigma2, int nlevels,
                      const float *q_Xw, const int32_t *q_octave, const float *q_angle, const uint8_t *q_desc, const uint8_t *q_valid,
                      const int32_t *q_counts, int q_slab, float th_proj, int th_dist, int check_ori,
                      float *Tcw, float *kp_xy, float *kp_angle, float *kp_response, int32_t *kp_octave, float *kp_size, uint8_t *desc,
                      int cap, int32_t *counts, int32_t *feat_match, int32_t *nmatches, uint8_t *outlier, int32_t *n_inliers);

/* Bench bookkeeping for the last orbo_bundle_adjust call: out4 = { seconds inside the LM loops (graph resident on the
 * device), seconds of the whole call, seconds of host graph layout + H2D, leading dimension of the reduced system }. */
int orbo_last_ba_timing(orbo_handle *h, double *out4);
/* Sparse structure of the reduced system in the last orbo_bundle_adjust call (what LinearSolverEigen's sparse LDLT with its
 * fill-reducing ordering exploits in the reference, linear_solver_eigen.h:94-124): out3 = { tile rows (64x64 tiles, 10 keyframes
 * each), structurally nonzero tiles of L after the nested-dissection tile ordering, levels of the elimination DAG }. */
int orbo_last_ba_structure(orbo_handle *h, long long *out3);
/* Per-kernel CUDA-event timing of the BA kernels.  ids: 0 errors, 1 build_points, 2 build_poses, 3 schur, 4 the sparse tiled
 * Cholesky factorisation (all k_chol_panel / k_chol_update launches of one solve), 5-6 unused, 7 triangular solves, 8 backsub, 9 update,
 * 10 memset of the reduced system. */
int orbo_set_profiling(orbo_handle *h, int enabled);
int orbo_get_kernel_times(orbo_handle *h, double *total_ms, long long *counts, int n);

#ifdef __cplusplus
}
#endif
#endif /* ORBSLAMM_B200_H */
