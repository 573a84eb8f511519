#include <cstdlib>
// orb_extractor.cu -- host driver and C-ABI of the ORB extractor (orbx_*).
//
// Replaces ORBextractor (S/include/ORBextractor.h:45-111, S/src/ORBextractor.cc).  The handle owns
// a CUDA stream and all device workspaces; a batch of frames is processed by a fixed sequence of
// launches (see orb_kernels.cu).  No CPU fallback exists: every entry point fails with ORBS_E_CUDA
// if the device is unusable.
#include <math.h>
#include <string.h>
#include <vector>
#include <mutex>
#include <algorithm>
#include "orb_extractor.cuh"
#include "orb_kernels.cuh"   // kernels live in this translation unit

namespace orbs {

static thread_local std::string g_last_error;
void set_last_error(const std::string &s) { g_last_error = s; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    cudaGetLastError();
    return ORBS_E_CUDA;
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

}  // namespace orbs

using namespace orbs;

struct orbx_handle {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;        // the 7x7 blur runs here, beside the quad-tree kernel (launch_pipeline)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool own_stream = true;
    std::vector<cudaEvent_t> chunk_events;
    // parameters and tables (ORBextractor.cc:410-470)
    int nfeatures = 0, nlevels = 0, ini_th = 0, min_th = 0;
    double scale_factor = 0;
    float scale[ORBS_MAX_LEVELS], inv_scale[ORBS_MAX_LEVELS], sigma2[ORBS_MAX_LEVELS], inv_sigma2[ORBS_MAX_LEVELS];
    int features_per_level[ORBS_MAX_LEVELS];
    int umax[kHalfPatch + 1];
    // plan for the current image shape
    bool have_plan = false;
    ExtractPlan plan;
    int octree_smem = 0;
    DevBuf d_cells, d_tiles, d_rs_tab;
    bool rs_words[ORBS_MAX_LEVELS] = {};   // level l's resize may use k_resize_level<true> (the 4 source columns of every thread within 8 bytes of an aligned word)
    // per-batch workspaces
    int batch_cap = 0;
    DevBuf d_stage;      // staged host images (64-byte pitch)
    DevBuf d_packed;     // host images as uploaded (flat DMA), re-pitched on the device
    size_t stage_pitch = 0, stage_frame = 0;
    DevBuf d_pyr, d_blur, d_cand, d_knode, d_lvl_kp, d_counts /* cand_count | lvl_count | counts | err */;
    DevBuf d_kp_xy, d_kp_angle, d_kp_resp, d_kp_oct, d_kp_size, d_desc;
    PinnedBuf h_counts;
    // state of the last call
    int last_frames = 0;
    const uint8_t *last_img0 = nullptr;
    int last_pitch0 = 0;
    size_t last_frame0 = 0;
    long long launches = 0;
    KernelTimer timer;   // ids: 0 resize, 1 fast_cells, 2 blur7, 3 octree, 4 orient_describe
    // TMA descriptors: one (x, y, frame) u8 tensor map per pyramid level and consumer (the box size is part of the descriptor)
    LevelTmaps tmaps_fast, tmaps_blur;
    unsigned tmap_levels = 0;            // bit l: level l has valid tensor maps
    const void *tmap_pyr_base = nullptr; int tmap_pyr_frames = 0;
    const void *tmap_img0 = nullptr; int tmap_img0_pitch = 0, tmap_img0_frames = 0; size_t tmap_img0_frame = 0;
    std::mutex mu;
};

// cuTensorMapEncodeTiled through the runtime (no -lcuda): u8 tensor (x, y, frame), box (box_w, box_h, 1), no swizzle, zero fill
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

static bool encode_level_map(CUtensorMap *m, const void *base, int w, int h, size_t pitch, size_t frame_bytes, int nframes, int box_w, int box_h)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || ((uintptr_t)base & 15) || (pitch & 15) || (frame_bytes & 15) || nframes <= 0 || box_w > 256 || box_h > 256 || (box_w & 15) || box_w <= 0 || box_h <= 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int kBlurBoxW = 112, kBlurBoxH = kBlurTH + 6;      // blur tile box: 16-byte left margin + 96 columns, 38 rows

// (re)encode the maps whose buffers changed: levels >= 1 live in d_pyr (batch_cap frames), level 0 is the caller's / staged image batch
static void refresh_tmaps(orbx_handle *h, const uint8_t *d_img0_all, int pitch0, size_t frame0, int n_frames_total)
{
    const ExtractPlan &P = h->plan;
    if (h->tmap_pyr_base != h->d_pyr.p || h->tmap_pyr_frames != h->batch_cap) {
        h->tmap_levels &= 1u;
        for (int l = 1; l < P.nlevels; l++) {
            const LevelPlan &L = P.lv[l];
            const uint8_t *base = h->d_pyr.as<uint8_t>() + L.pyr_off;
            if (L.cell_count > 0 && encode_level_map(&h->tmaps_fast.m[l], base, L.w, L.h, L.pitch, P.pyr_frame_bytes, h->batch_cap, L.fast_box_w, L.fast_box_h) &&
                encode_level_map(&h->tmaps_blur.m[l], base, L.w, L.h, L.pitch, P.pyr_frame_bytes, h->batch_cap, kBlurBoxW, kBlurBoxH))
                h->tmap_levels |= 1u << l;
        }
        h->tmap_pyr_base = h->d_pyr.p; h->tmap_pyr_frames = h->batch_cap;
    }
    if (h->tmap_img0 != d_img0_all || h->tmap_img0_pitch != pitch0 || h->tmap_img0_frame != frame0 || h->tmap_img0_frames != n_frames_total) {
        const LevelPlan &L = P.lv[0];
        h->tmap_levels &= ~1u;
        const size_t fb = n_frames_total > 1 ? frame0 : align_up((size_t)pitch0 * L.h, 16);
        if (L.cell_count > 0 && pitch0 >= (int)align_up(L.w, 16) && encode_level_map(&h->tmaps_fast.m[0], d_img0_all, L.w, L.h, (size_t)pitch0, fb, n_frames_total, L.fast_box_w, L.fast_box_h) &&
            encode_level_map(&h->tmaps_blur.m[0], d_img0_all, L.w, L.h, (size_t)pitch0, fb, n_frames_total, kBlurBoxW, kBlurBoxH))
            h->tmap_levels |= 1u;
        h->tmap_img0 = d_img0_all; h->tmap_img0_pitch = pitch0; h->tmap_img0_frame = frame0; h->tmap_img0_frames = n_frames_total;
    }
}

static int build_plan(orbx_handle *h, int width, int height)
{
    h->tmap_pyr_base = nullptr; h->tmap_img0 = nullptr; h->tmap_levels = 0;

    ExtractPlan &P = h->plan;
    memset(&P, 0, sizeof P);
    P.nlevels = h->nlevels; P.width = width; P.height = height;
    P.ini_th = h->ini_th; P.min_th = h->min_th;
    memcpy(P.umax, h->umax, sizeof P.umax);
    std::vector<CellDesc> cells;
    std::vector<TileDesc> tiles;
    std::vector<int2> rs;
    size_t pyr_off = 0, cand_off = 0;
    int kp_off = 0, kp_slab = 0;
    int max_node_cap = 1;
    for (int l = 0; l < h->nlevels; l++) {
        LevelPlan &L = P.lv[l];
        L.w = cv_round_f((float)width * h->inv_scale[l]);      // ORBextractor.cc:1111-1112
        L.h = cv_round_f((float)height * h->inv_scale[l]);
        ORBS_REQUIRE(L.w >= 1 && L.h >= 1 && L.w < 32768 && L.h < 32768, ORBS_E_SHAPE, "image shape unsupported (level collapses or >= 32768 px)");
        L.pitch = (int)align_up(L.w, 64);
        L.nfeat = h->features_per_level[l];
        L.scale = h->scale[l];
        L.patch = (int)(kPatch * h->scale[l]);
        L.pyr_off = pyr_off;
        pyr_off += align_up((size_t)L.pitch * L.h, 256);
        // FAST cells (ORBextractor.cc:771-829)
        const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        L.bw = maxBX - kMinBorder; L.bh = maxBY - kMinBorder;
        L.cell_first = (int)cells.size();
        L.cand_cap = 0;
        const float fw = (float)L.bw, fh = (float)L.bh;
        const int nCols = L.bw > 0 ? (int)(fw / 30.f) : 0, nRows = L.bh > 0 ? (int)(fh / 30.f) : 0;
        if (nCols >= 1 && nRows >= 1) {
            const int wCell = (int)ceilf(fw / nCols), hCell = (int)ceilf(fh / nRows);
            ORBS_REQUIRE(nCols < 1024 && nRows < 1024, ORBS_E_SHAPE, "image too large (more than 1023 FAST cells per side)");
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(kMinBorder + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(kMinBorder + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    CellDesc c;
                    c.level = (short)l; c.x0 = (short)iniX; c.y0 = (short)iniY;
                    c.cw = (short)((int)maxX - (int)iniX); c.ch = (short)((int)maxY - (int)iniY);
                    c.ci = (short)i; c.cj = (short)j; c.addx = (short)(j * wCell); c.addy = (short)(i * hCell);
                    if (c.cw > 66 || c.ch > 66) { set_last_error("internal: FAST cell larger than 66 px"); return ORBS_E_SHAPE; }
                    if (c.cw > 6 && c.ch > 6) {
                        cells.push_back(c);
                        L.cand_cap += ((c.cw - 6 + 1) / 2) * ((c.ch - 6 + 1) / 2);
                    }
                }
            }
        }
        L.cell_count = (int)cells.size() - L.cell_first;
        {   // TMA box of this level's FAST cells: widest 16-byte aligned row span x tallest cell
            int mw = 16, mh = 1;
            for (int c = L.cell_first; c < (int)cells.size(); c++) {
                const int xa = (cells[c].x0 - 1) & ~15;
                mw = std::max(mw, (int)align_up((size_t)(cells[c].x0 - 1 - xa) + cells[c].cw + 1, 16)); mh = std::max(mh, (int)cells[c].ch);
            }
            L.fast_box_w = mw; L.fast_box_h = mh;
        }
        // quad-tree roots (ORBextractor.cc:542-545)
        if (L.cell_count > 0) {
            L.n_ini = (int)roundf((float)L.bw / (float)L.bh);
            ORBS_REQUIRE(L.n_ini >= 1, ORBS_E_SHAPE, "image more than twice as tall as wide: the reference's quad-tree has no root");
            L.hx = (float)L.bw / (float)L.n_ini;
        } else { L.n_ini = 1; L.hx = (float)(L.bw > 0 ? L.bw : 1); }
        L.node_cap = L.nfeat + 3 > 4 * L.n_ini ? L.nfeat + 3 : 4 * L.n_ini;
        if (L.cell_count == 0) L.node_cap = 1;
        if (L.node_cap > max_node_cap) max_node_cap = L.node_cap;
        L.kp_off = kp_off; kp_off += L.node_cap;
        kp_slab += L.cell_count ? L.node_cap : 0;
        L.cand_off = cand_off; cand_off += (size_t)L.cand_cap;
        // blur tiles
        L.tile_first = (int)tiles.size();
        if (L.cell_count > 0)
            for (int y = 0; y < L.h; y += kBlurTH)
                for (int x = 0; x < L.w; x += kBlurTW) tiles.push_back(TileDesc{(short)l, (short)x, (short)y});
        L.tile_count = (int)tiles.size() - L.tile_first;
        // resize tables (cv::resize INTER_LINEAR, fixed point; level l from level l-1)
        if (l > 0) {
            const LevelPlan &S = P.lv[l - 1];
            for (int axis = 0; axis < 2; axis++) {
                const int ssize = axis ? S.h : S.w, dsize = axis ? L.h : L.w;
                while (rs.size() & 3) rs.push_back(make_int2(0, 0));        // 32-byte aligned tables: 4 entries per vector load
                (axis ? L.rs_y_off : L.rs_x_off) = (int)rs.size();
                const double inv = (double)dsize / ssize, sc = 1. / inv;
                for (int d = 0; d < dsize; d++) {
                    float fx = (float)((d + 0.5) * sc - 0.5);
                    int sx = (int)floorf(fx);
                    fx -= sx;
                    if (sx < 0) { fx = 0; sx = 0; }
                    if (sx >= ssize - 1) { fx = 0; sx = ssize - 1; }
                    const int a0 = (short)cv_round_f((1.f - fx) * 2048.f), a1 = (short)cv_round_f(fx * 2048.f);
                    rs.push_back(make_int2(sx, (a1 << 16) | (a0 & 0xffff)));
                }
                if (axis == 0) {
                    bool ok = true;
                    const int2 *tx = rs.data() + L.rs_x_off;
                    for (int d = 0; d < dsize && ok; d += 4) ok = tx[std::min(d + 3, dsize - 1)].x - (tx[d].x & ~3) <= 7;
                    h->rs_words[l] = ok;
                }
            }
        }
    }
    P.pyr_frame_bytes = pyr_off;
    P.cand_frame_entries = cand_off ? cand_off : 1;
    P.lvl_slab = kp_off;
    P.kp_slab = kp_slab > 0 ? kp_slab : 1;
    P.total_cells = (int)cells.size();
    P.total_tiles = (int)tiles.size();
    // octree shared memory: 64 B per node + scan scratch
    h->octree_smem = max_node_cap * 64 + (512 + 1) * 4 + 64;
    ORBS_REQUIRE(h->octree_smem <= 220 * 1024, ORBS_E_INVALID, "nfeatures too large for the on-chip quad-tree (per-level quota > ~3400)");
    // the opt-in limit is per device and per FUNCTION (shared by every extractor handle, e.g. Tracking's mpIniORBextractor with
    // 2 * nFeatures next to mpORBextractorLeft): always the static upper bound, so a smaller handle never lowers it
    ORBS_CUDA(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    if (!cells.empty()) {
        if (int rc = h->d_cells.reserve(cells.size() * sizeof(CellDesc))) return rc;
        ORBS_CUDA(cudaMemcpyAsync(h->d_cells.p, cells.data(), cells.size() * sizeof(CellDesc), cudaMemcpyHostToDevice, h->stream));
    }
    if (!tiles.empty()) {
        if (int rc = h->d_tiles.reserve(tiles.size() * sizeof(TileDesc))) return rc;
        ORBS_CUDA(cudaMemcpyAsync(h->d_tiles.p, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, h->stream));
    }
    if (!rs.empty()) {
        if (int rc = h->d_rs_tab.reserve(rs.size() * sizeof(int2))) return rc;
        ORBS_CUDA(cudaMemcpyAsync(h->d_rs_tab.p, rs.data(), rs.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    }
    ORBS_CUDA(cudaStreamSynchronize(h->stream));   // host vectors go out of scope
    h->have_plan = true;
    h->batch_cap = 0;
    return ORBS_OK;
}

static int ensure_plan(orbx_handle *h, int width, int height)
{
    if (h->have_plan && h->plan.width == width && h->plan.height == height) return ORBS_OK;
    h->have_plan = false;
    return build_plan(h, width, height);
}

static int ensure_batch(orbx_handle *h, int n)
{
    if (n <= h->batch_cap) return ORBS_OK;
    const ExtractPlan &P = h->plan;
    int rc;
    if ((rc = h->d_pyr.reserve(P.pyr_frame_bytes * n))) return rc;
    if ((rc = h->d_blur.reserve(P.pyr_frame_bytes * n))) return rc;
    if ((rc = h->d_cand.reserve(P.cand_frame_entries * n * sizeof(uint2)))) return rc;
    if ((rc = h->d_knode.reserve(P.cand_frame_entries * n * sizeof(unsigned)))) return rc;
    if ((rc = h->d_lvl_kp.reserve((size_t)P.lvl_slab * n * sizeof(uint2)))) return rc;
    if ((rc = h->d_counts.reserve(((size_t)2 * P.nlevels * n + n + 16) * sizeof(int)))) return rc;
    const size_t slab = (size_t)P.kp_slab * n;
    if ((rc = h->d_kp_xy.reserve(slab * 8))) return rc;
    if ((rc = h->d_kp_angle.reserve(slab * 4))) return rc;
    if ((rc = h->d_kp_resp.reserve(slab * 4))) return rc;
    if ((rc = h->d_kp_oct.reserve(slab * 4))) return rc;
    if ((rc = h->d_kp_size.reserve(slab * 4))) return rc;
    if ((rc = h->d_desc.reserve(slab * 32))) return rc;
    if ((rc = h->h_counts.reserve(((size_t)n * (1 + P.nlevels) + 16) * sizeof(int)))) return rc;
    h->batch_cap = n;
    return ORBS_OK;
}

// device pointers into d_counts
static inline int *cand_count_ptr(orbx_handle *h) { return h->d_counts.as<int>(); }
static inline int *lvl_count_ptr(orbx_handle *h, int n) { return h->d_counts.as<int>() + (size_t)h->plan.nlevels * n; }
static inline int *counts_ptr(orbx_handle *h, int n) { return h->d_counts.as<int>() + (size_t)2 * h->plan.nlevels * n; }
static inline int *err_ptr(orbx_handle *h, int n) { return counts_ptr(h, n) + n; }

static int reset_counts(orbx_handle *h)
{
    const int nb = h->batch_cap;
    ORBS_CUDA(cudaMemsetAsync(h->d_counts.p, 0, ((size_t)2 * h->plan.nlevels * nb + nb + 16) * sizeof(int), h->stream));
    return ORBS_OK;
}

// frames [f0, f0 + n) of the batch whose first image is d_img0_all; every per-frame workspace is addressed by frame index,
// so chunks of one batch can be processed back to back (host path: chunk c+1 is uploaded while chunk c computes)
static int launch_pipeline(orbx_handle *h, const uint8_t *d_img0_all, int f0, int n, int pitch0, size_t frame0)
{
    const ExtractPlan &P = h->plan;
    cudaStream_t st = h->stream;
    const int nb = h->batch_cap;     // layout of d_counts uses the allocated capacity
    const uint8_t *d_img0 = d_img0_all + (size_t)f0 * frame0;
    uint8_t *pyr = h->d_pyr.as<uint8_t>() + (size_t)f0 * P.pyr_frame_bytes;
    uint8_t *blur = h->d_blur.as<uint8_t>() + (size_t)f0 * P.pyr_frame_bytes;
    uint2 *cand = h->d_cand.as<uint2>() + (size_t)f0 * P.cand_frame_entries;
    unsigned *knode = h->d_knode.as<unsigned>() + (size_t)f0 * P.cand_frame_entries;
    uint2 *lvl_kp = h->d_lvl_kp.as<uint2>() + (size_t)f0 * P.lvl_slab;
    int *cand_count = cand_count_ptr(h) + (size_t)f0 * P.nlevels, *lvl_count = lvl_count_ptr(h, nb) + (size_t)f0 * P.nlevels;
    int *counts = counts_ptr(h, nb) + f0;
    const size_t ko = (size_t)f0 * P.kp_slab;
    // pyramid chain
    for (int l = 1; l < P.nlevels; l++) {
        const LevelPlan &S = P.lv[l - 1], &D = P.lv[l];
        const uint8_t *src; int spitch; size_t sframe;
        if (l == 1) { src = d_img0; spitch = pitch0; sframe = frame0; }
        else { src = pyr + S.pyr_off; spitch = S.pitch; sframe = P.pyr_frame_bytes; }
        dim3 grid((D.w + 255) / 256, (D.h + 3) / 4, n);                     // 64 x 4 threads, 4 output pixels per thread
        h->timer.begin(0, st);
        // word loads need 4-byte aligned source rows (always true for pyramid levels; level 0 may be a caller-owned device buffer)
        const bool words = h->rs_words[l] && ((((uintptr_t)src | (uintptr_t)spitch | (uintptr_t)sframe) & 3) == 0);
        (words ? k_resize_level<true> : k_resize_level<false>)<<<grid, 256, 0, st>>>(src, S.w, S.h, spitch, sframe, pyr + D.pyr_off, D.w, D.h, D.pitch, P.pyr_frame_bytes,
                                                                                      h->d_rs_tab.as<int2>() + D.rs_x_off, h->d_rs_tab.as<int2>() + D.rs_y_off);
        h->timer.end(st);
        h->launches++;
    }
    if (P.total_cells > 0) {
        h->timer.begin(1, st);
        // pyramid levels >= 1 are always 64-byte aligned; level 0 goes through the TMA engine when the caller's buffer allows it
        const bool l0_bulk = (((uintptr_t)d_img0 | (uintptr_t)pitch0 | (uintptr_t)frame0) & 15) == 0 && pitch0 >= (int)align_up(P.lv[0].w, 16);
        const unsigned tma_levels = (l0_bulk ? 1u : 0u) | 0xfffffffeu;
        k_fast_cells<<<dim3(P.total_cells, n), 128, 0, st>>>(P, h->tmaps_fast, h->tmap_levels, f0, tma_levels, h->d_cells.as<CellDesc>(), d_img0, pitch0, frame0, pyr, cand,
                                                             cand_count, err_ptr(h, nb));
        h->timer.end(st);
        // The blur needs only the pyramid, the quad-tree only the FAST candidates, the descriptors both: the blur (issue-bound, 90 % of the issue slots)
        // goes to a side stream and runs BESIDE the quad-tree kernel (barrier-latency bound, ~37 %) instead of before it.
        cudaStream_t sb = st;
        if (h->side_stream) {
            ORBS_CUDA(cudaEventRecord(h->ev_fork, st));
            ORBS_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
            sb = h->side_stream;
        }
        h->timer.begin(2, sb);
        k_blur7<<<dim3(P.total_tiles, n), 256, 0, sb>>>(P, h->tmaps_blur, h->tmap_levels, f0, tma_levels, h->d_tiles.as<TileDesc>(), d_img0, pitch0, frame0, pyr, blur);
        h->timer.end(sb);
        if (h->side_stream) ORBS_CUDA(cudaEventRecord(h->ev_join, sb));
        h->timer.begin(3, st);
        k_octree<<<dim3(P.nlevels, n), 512, h->octree_smem, st>>>(P, cand, cand_count, knode, lvl_kp, lvl_count, err_ptr(h, nb));
        h->timer.end(st);
        if (h->side_stream) ORBS_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
        const int warps_per_block = 8;
        h->timer.begin(4, st);
        k_orient_describe<<<dim3((P.kp_slab + warps_per_block - 1) / warps_per_block, n), warps_per_block * 32, 0, st>>>(
            P, d_img0, pitch0, frame0, pyr, blur, lvl_kp, lvl_count, h->d_kp_xy.as<float2>() + ko, h->d_kp_angle.as<float>() + ko,
            h->d_kp_resp.as<float>() + ko, h->d_kp_oct.as<int>() + ko, h->d_kp_size.as<float>() + ko, h->d_desc.as<uint8_t>() + 32 * ko, counts);
        h->timer.end(st);
        h->launches += 4;
    }
    ORBS_CUDA(cudaGetLastError());
    h->last_frames = f0 + n; h->last_img0 = d_img0_all; h->last_pitch0 = pitch0; h->last_frame0 = frame0;
    return ORBS_OK;
}

// host images -> staged device copy, uploaded in chunks on a copy stream so that the upload of chunk c+1 overlaps the
// extraction of chunk c
static int upload_and_extract(orbx_handle *h, const uint8_t *images, int n_frames, int width, int height, int stride, size_t frame_stride)
{
    const size_t pitch = align_up(width, 64), frame = pitch * height;
    if (int rc = h->d_stage.reserve(frame * n_frames)) return rc;
    if (!h->copy_stream) ORBS_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (int rc = reset_counts(h)) return rc;
    refresh_tmaps(h, h->d_stage.as<uint8_t>(), (int)pitch, frame, n_frames);
    const int chunk = n_frames > 16 ? 16 : n_frames;
    const int nchunks = (n_frames + chunk - 1) / chunk;
    while ((int)h->chunk_events.size() < nchunks + 1) { cudaEvent_t e; ORBS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->chunk_events.push_back(e); }
    // the copy stream must not overwrite the staging buffer while an earlier call still reads it
    ORBS_CUDA(cudaEventRecord(h->chunk_events[nchunks], h->stream));
    ORBS_CUDA(cudaStreamWaitEvent(h->copy_stream, h->chunk_events[nchunks], 0));
    // contiguous frames: one flat DMA per chunk into d_packed, re-pitched on the device (k_repitch) -- full PCIe rate
    const bool flat = frame_stride == (size_t)stride * height;
    if (flat) { if (int rc = h->d_packed.reserve(frame_stride * n_frames + 64)) return rc; }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * chunk, n = std::min(chunk, n_frames - f0);
        if (flat) {
            ORBS_CUDA(cudaMemcpyAsync(h->d_packed.as<uint8_t>() + f0 * frame_stride, images + f0 * frame_stride, frame_stride * n, cudaMemcpyHostToDevice, h->copy_stream));
        } else {
            for (int f = f0; f < f0 + n; f++)
                ORBS_CUDA(cudaMemcpy2DAsync(h->d_stage.as<uint8_t>() + f * frame, pitch, images + f * frame_stride, stride, width, height,
                                            cudaMemcpyHostToDevice, h->copy_stream));
        }
        ORBS_CUDA(cudaEventRecord(h->chunk_events[c], h->copy_stream));
    }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * chunk, n = std::min(chunk, n_frames - f0);
        ORBS_CUDA(cudaStreamWaitEvent(h->stream, h->chunk_events[c], 0));
        if (flat) {
            const int threads = (int)(pitch >> 4) * height;
            k_repitch<<<dim3((threads + 255) / 256, n), 256, 0, h->stream>>>(h->d_packed.as<uint8_t>() + f0 * frame_stride, stride, frame_stride,
                                                                             h->d_stage.as<uint8_t>() + f0 * frame, (int)pitch, frame, width, height);
            h->launches++;
        }
        if (int rc = launch_pipeline(h, h->d_stage.as<uint8_t>(), f0, n, (int)pitch, frame)) return rc;
    }
    return ORBS_OK;
}

static int download_results(orbx_handle *h, float *kp_xy, float *kp_angle, float *kp_response, int32_t *kp_octave,
                            float *kp_size, uint8_t *desc, int cap, int32_t *counts)
{
    const ExtractPlan &P = h->plan;
    const int n = h->last_frames, nb = h->batch_cap;
    ORBS_REQUIRE(n > 0, ORBS_E_STATE, "no extraction result to download");
    cudaStream_t st = h->stream;
    int *hc = h->h_counts.as<int>();
    ORBS_CUDA(cudaMemcpyAsync(hc, counts_ptr(h, nb), (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    // err flag sits right after counts[nb]; fetch separately when n != nb
    int *herr = hc + n + 1;
    ORBS_CUDA(cudaMemcpyAsync(herr, err_ptr(h, nb), sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaStreamSynchronize(st));
    if (*herr != 0) { set_last_error(*herr == 1 ? "internal: FAST candidate buffer overflow" : *herr == 3 ? "internal: bulk copy of a FAST cell did not complete" : "internal: quad-tree node capacity exceeded"); return ORBS_E_CAPACITY; }
    for (int f = 0; f < n; f++) {
        const int c = hc[f];
        counts[f] = c;
        if (c > cap) { set_last_error("output capacity too small; size with orbx_max_keypoints()"); return ORBS_E_CAPACITY; }
    }
    int maxc = 0;
    for (int f = 0; f < n; f++) maxc = std::max(maxc, hc[f]);
    if (maxc > 0) {
        // one strided copy per array: rows = frames, row width = the largest keypoint count of the batch
        auto copy2d = [&](void *dst, const void *src, size_t elem) -> int {
            ORBS_CUDA(cudaMemcpy2DAsync(dst, (size_t)cap * elem, src, (size_t)P.kp_slab * elem, (size_t)maxc * elem, n, cudaMemcpyDeviceToHost, st));
            return ORBS_OK;
        };
        int rc = ORBS_OK;
        if (kp_xy && (rc = copy2d(kp_xy, h->d_kp_xy.p, 8))) return rc;
        if (kp_angle && (rc = copy2d(kp_angle, h->d_kp_angle.p, 4))) return rc;
        if (kp_response && (rc = copy2d(kp_response, h->d_kp_resp.p, 4))) return rc;
        if (kp_octave && (rc = copy2d(kp_octave, h->d_kp_oct.p, 4))) return rc;
        if (kp_size && (rc = copy2d(kp_size, h->d_kp_size.p, 4))) return rc;
        if (desc && (rc = copy2d(desc, h->d_desc.p, 32))) return rc;
    }
    ORBS_CUDA(cudaStreamSynchronize(st));
    return ORBS_OK;
}

extern "C" {

const char *orbs_last_error(void) { return g_last_error.c_str(); }
int orbs_version(void) { return 100; }
int orbs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int orbx_create(orbx_handle **out, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device)
{
    ORBS_REQUIRE(out, ORBS_E_INVALID, "orbx_create: null out pointer");
    *out = nullptr;
    ORBS_REQUIRE(nfeatures > 0 && nlevels >= 1 && nlevels <= ORBS_MAX_LEVELS && scale_factor > 1.0f, ORBS_E_INVALID,
                 "orbx_create: need nfeatures > 0, 1 <= nlevels <= 16, scaleFactor > 1");
    ORBS_REQUIRE(ini_th >= 1 && ini_th <= 254 && min_th >= 1 && min_th <= 254, ORBS_E_INVALID, "orbx_create: FAST thresholds must be in [1, 254]");
    ORBS_CUDA(cudaSetDevice(device));
    orbx_handle *h = new orbx_handle();
    h->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    {
        const char *off = getenv("ORBS_NO_SIDE_STREAM");
        if (!(off && off[0] == '1') && cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) == cudaSuccess) {
            if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
                cudaStreamDestroy(h->side_stream); h->side_stream = nullptr;
            }
        } else h->side_stream = nullptr;
        cudaGetLastError();
    }
    h->nfeatures = nfeatures; h->nlevels = nlevels; h->ini_th = ini_th; h->min_th = min_th;
    h->scale_factor = scale_factor;                       // member is double (ORBextractor.h:98)
    h->scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        h->scale[i] = (float)(h->scale[i - 1] * h->scale_factor);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < nlevels; i++) { h->inv_scale[i] = 1.0f / h->scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    const float factor = (float)(1.0f / h->scale_factor);
    float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        h->features_per_level[l] = cv_round_f(desired);
        sum += h->features_per_level[l];
        desired *= factor;
    }
    h->features_per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    int v, v0;
    const int vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= vmax; ++v) h->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
        h->umax[v] = v0;
        ++v0;
    }
    *out = h;
    return ORBS_OK;
}

int orbx_destroy(orbx_handle *h)
{
    if (!h) return ORBS_OK;
    cudaSetDevice(h->device);
    if (h->stream && h->own_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    else cudaDeviceSynchronize();
    DevBuf *bufs[] = {&h->d_cells, &h->d_tiles, &h->d_rs_tab, &h->d_stage, &h->d_packed, &h->d_pyr, &h->d_blur, &h->d_cand, &h->d_knode, &h->d_lvl_kp,
                      &h->d_counts, &h->d_kp_xy, &h->d_kp_angle, &h->d_kp_resp, &h->d_kp_oct, &h->d_kp_size, &h->d_desc};
    for (DevBuf *b : bufs) b->release();
    h->h_counts.release();
    h->timer.release();
    for (cudaEvent_t e : h->chunk_events) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->side_stream) { cudaStreamSynchronize(h->side_stream); cudaStreamDestroy(h->side_stream); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return ORBS_OK;
}

int orbx_get_tables(const orbx_handle *h, int *nlevels, float *scale_factor, float *scale, float *inv_scale, float *sigma2,
                    float *inv_sigma2, int *features_per_level)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    if (nlevels) *nlevels = h->nlevels;
    if (scale_factor) *scale_factor = (float)h->scale_factor;
    for (int i = 0; i < h->nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->inv_scale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
        if (features_per_level) features_per_level[i] = h->features_per_level[i];
    }
    return ORBS_OK;
}

int orbx_max_keypoints(orbx_handle *h, int width, int height, int *max_kp)
{
    ORBS_REQUIRE(h && max_kp, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(width > 0 && height > 0, ORBS_E_INVALID, "non-positive image size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    if (int rc = ensure_plan(h, width, height)) return rc;
    *max_kp = h->plan.kp_slab;
    return ORBS_OK;
}

int orbx_level_size(orbx_handle *h, int width, int height, int level, int *lw, int *lh)
{
    ORBS_REQUIRE(h && lw && lh && level >= 0 && level < h->nlevels, ORBS_E_INVALID, "bad argument");
    *lw = cv_round_f((float)width * h->inv_scale[level]);
    *lh = cv_round_f((float)height * h->inv_scale[level]);
    return ORBS_OK;
}

int orbx_extract_device(orbx_handle *h, const uint8_t *d_images, int n_frames, int width, int height, int stride, size_t frame_stride)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    ORBS_REQUIRE(n_frames > 0, ORBS_E_INVALID, "n_frames must be positive");
    ORBS_REQUIRE(width > 0 && height > 0, ORBS_E_INVALID, "use orbx_extract for empty images");
    ORBS_REQUIRE(d_images && stride >= width && (n_frames == 1 || frame_stride >= (size_t)stride * (height - 1) + width), ORBS_E_INVALID, "bad image pointer or strides");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    if (int rc = ensure_plan(h, width, height)) return rc;
    if (int rc = ensure_batch(h, n_frames)) return rc;
    if (int rc = reset_counts(h)) return rc;
    refresh_tmaps(h, d_images, stride, frame_stride, n_frames);
    return launch_pipeline(h, d_images, 0, n_frames, stride, frame_stride);
}

int orbx_extract(orbx_handle *h, const uint8_t *images, int n_frames, int width, int height, int stride, size_t frame_stride,
                 float *kp_xy, float *kp_angle, float *kp_response, int32_t *kp_octave, float *kp_size, uint8_t *desc, int cap,
                 int32_t *counts)
{
    ORBS_REQUIRE(h && counts, ORBS_E_INVALID, "null handle or counts");
    ORBS_REQUIRE(n_frames > 0, ORBS_E_INVALID, "n_frames must be positive");
    if (width <= 0 || height <= 0) {               // empty image: silent return (ORBextractor.cc:1046-1047)
        for (int f = 0; f < n_frames; f++) counts[f] = 0;
        return ORBS_OK;
    }
    ORBS_REQUIRE(images && stride >= width, ORBS_E_INVALID, "bad image pointer or stride");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    if (int rc = ensure_plan(h, width, height)) return rc;
    if (int rc = ensure_batch(h, n_frames)) return rc;
    if (int rc = upload_and_extract(h, images, n_frames, width, height, stride, frame_stride)) return rc;
    return download_results(h, kp_xy, kp_angle, kp_response, kp_octave, kp_size, desc, cap, counts);
}

int orbx_extract_host_async(orbx_handle *h, const uint8_t *images, int n_frames, int width, int height, int stride, size_t frame_stride)
{
    ORBS_REQUIRE(h && images, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && width > 0 && height > 0 && stride >= width, ORBS_E_INVALID, "bad image shape");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    if (int rc = ensure_plan(h, width, height)) return rc;
    if (int rc = ensure_batch(h, n_frames)) return rc;
    return upload_and_extract(h, images, n_frames, width, height, stride, frame_stride);
}

int orbx_device_results(orbx_handle *h, orbx_device_view *view)
{
    ORBS_REQUIRE(h && view, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(h->last_frames > 0, ORBS_E_STATE, "no extraction has run");
    view->n_frames = h->last_frames; view->slab = h->plan.kp_slab;
    view->kp_xy = h->d_kp_xy.as<float>(); view->kp_angle = h->d_kp_angle.as<float>();
    view->kp_response = h->d_kp_resp.as<float>(); view->kp_octave = h->d_kp_oct.as<int>();
    view->kp_size = h->d_kp_size.as<float>(); view->desc = h->d_desc.as<uint8_t>();
    view->counts = counts_ptr(h, h->batch_cap); view->level_counts = lvl_count_ptr(h, h->batch_cap);
    return ORBS_OK;
}

int orbx_download(orbx_handle *h, float *kp_xy, float *kp_angle, float *kp_response, int32_t *kp_octave, float *kp_size,
                  uint8_t *desc, int cap, int32_t *counts)
{
    ORBS_REQUIRE(h && counts, ORBS_E_INVALID, "null handle or counts");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    return download_results(h, kp_xy, kp_angle, kp_response, kp_octave, kp_size, desc, cap, counts);
}

int orbx_get_pyramid_level(orbx_handle *h, int frame, int level, int border, uint8_t *out, int out_stride)
{
    ORBS_REQUIRE(h && out, ORBS_E_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_REQUIRE(h->last_frames > 0 && frame >= 0 && frame < h->last_frames, ORBS_E_STATE, "no such frame in the last call");
    ORBS_REQUIRE(level >= 0 && level < h->nlevels, ORBS_E_INVALID, "bad level");
    ORBS_CUDA(cudaSetDevice(h->device));
    const LevelPlan &L = h->plan.lv[level];
    const uint8_t *src; int pitch;
    if (level == 0) { src = h->last_img0 + (size_t)frame * h->last_frame0; pitch = h->last_pitch0; }
    else { src = h->d_pyr.as<uint8_t>() + (size_t)frame * h->plan.pyr_frame_bytes + L.pyr_off; pitch = L.pitch; }
    if (!border) {
        ORBS_REQUIRE(out_stride >= L.w, ORBS_E_INVALID, "out_stride too small");
        ORBS_CUDA(cudaMemcpy2DAsync(out, out_stride, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
    } else {
        const int W = L.w + 2 * kEdge, H = L.h + 2 * kEdge;
        ORBS_REQUIRE(out_stride >= W, ORBS_E_INVALID, "out_stride too small");
        DevBuf tmp;
        if (int rc = tmp.reserve((size_t)W * H)) return rc;
        k_border_copy<<<dim3((W + 255) / 256, H), 256, 0, h->stream>>>(src, L.w, L.h, pitch, tmp.as<uint8_t>(), kEdge);
        h->launches++;
        cudaError_t e = cudaMemcpy2DAsync(out, out_stride, tmp.p, W, W, H, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        tmp.release();
        if (e != cudaSuccess) return cuda_fail(e, "pyramid border copy", __FILE__, __LINE__);
        return ORBS_OK;
    }
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    return ORBS_OK;
}

int orbx_get_candidates(orbx_handle *h, int frame, int level, int32_t *xys, int cap, int *n_out)
{
    ORBS_REQUIRE(h && n_out, ORBS_E_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_REQUIRE(h->last_frames > 0 && frame >= 0 && frame < h->last_frames, ORBS_E_STATE, "no such frame in the last call");
    ORBS_REQUIRE(level >= 0 && level < h->nlevels, ORBS_E_INVALID, "bad level");
    ORBS_CUDA(cudaSetDevice(h->device));
    const LevelPlan &L = h->plan.lv[level];
    int n = 0;
    ORBS_CUDA(cudaMemcpyAsync(&n, cand_count_ptr(h) + frame * h->plan.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    *n_out = n;
    if (!xys || n == 0) return ORBS_OK;
    ORBS_REQUIRE(n <= cap, ORBS_E_CAPACITY, "candidate output capacity too small");
    std::vector<uint2> tmp(n);
    ORBS_CUDA(cudaMemcpyAsync(tmp.data(), h->d_cand.as<uint2>() + (size_t)frame * h->plan.cand_frame_entries + L.cand_off,
                              (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost, h->stream));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; i++) {
        xys[3 * i] = (int)(tmp[i].x & 0xffff); xys[3 * i + 1] = (int)(tmp[i].x >> 16); xys[3 * i + 2] = (int)(tmp[i].y & 0xff);
    }
    return ORBS_OK;
}

int orbx_set_stream(orbx_handle *h, void *stream)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)stream; h->own_stream = false;
    return ORBS_OK;
}

void *orbx_stream(orbx_handle *h) { return h ? (void *)h->stream : nullptr; }
int orbx_synchronize(orbx_handle *h)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    return ORBS_OK;
}
long long orbx_kernel_launches(const orbx_handle *h) { return h ? h->launches : 0; }

int orbx_set_profiling(orbx_handle *h, int enabled)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    h->timer.collect();
    h->timer.enabled = enabled != 0;
    if (enabled) h->timer.reset();
    return ORBS_OK;
}

int orbx_get_kernel_times(orbx_handle *h, double *total_ms, long long *counts, int n)
{
    ORBS_REQUIRE(h && total_ms && counts && n > 0, ORBS_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    h->timer.collect();
    for (int i = 0; i < n && i < KernelTimer::kMaxKernels; i++) { total_ms[i] = h->timer.total_ms[i]; counts[i] = h->timer.count[i]; }
    return ORBS_OK;
}

}  // extern "C"
