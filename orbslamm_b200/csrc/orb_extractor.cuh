// orb_extractor.cuh -- plan structures shared by the extractor kernels and their host driver.
#pragma once
#include "common.cuh"

namespace orbs {

constexpr int kEdge = 19;          // EDGE_THRESHOLD, ORBextractor.cc:73
constexpr int kHalfPatch = 15;     // HALF_PATCH_SIZE
constexpr int kPatch = 31;         // PATCH_SIZE
constexpr int kMinBorder = kEdge - 3;

struct LevelPlan {
    int w, h, pitch;          // level size and row pitch (bytes) in the pyramid buffer
    int nfeat;                // mnFeaturesPerLevel[l]
    int n_ini;                // quad-tree roots, round(width/height)
    int node_cap;             // max(nfeat + 3, 4 * n_ini)
    int cand_cap;             // strict upper bound of FAST candidates after cell-local NMS
    int kp_off;               // offset of this level inside the per-frame level-keypoint slab
    int cell_first, cell_count;
    int bw, bh;               // maxBorder - minBorder (the region DistributeOctTree works on)
    int patch;                // (int)(PATCH_SIZE * scale)
    float scale;              // mvScaleFactor[l]
    float hx;                 // (float)bw / n_ini
    unsigned long long pyr_off;   // byte offset inside one frame's pyramid slab (levels >= 1)
    unsigned long long cand_off;  // entry offset inside one frame's candidate slab
    int rs_x_off, rs_y_off;   // offsets (entries) of this level's resize tables
    int tile_first, tile_count; // blur tiles
    int fast_box_w, fast_box_h; // TMA box of this level's FAST cells: widest aligned row span (multiple of 16) x tallest cell
};

struct ExtractPlan {
    int nlevels;
    int width, height;
    int ini_th, min_th;
    int kp_slab;                    // per-frame capacity of the final keypoint arrays
    int lvl_slab;                   // per-frame capacity of the per-level keypoint slab (sum node_cap)
    int total_cells, total_tiles;
    unsigned long long pyr_frame_bytes;   // one frame's pyramid slab (levels 1..)
    unsigned long long cand_frame_entries;
    int umax[kHalfPatch + 1];
    LevelPlan lv[ORBS_MAX_LEVELS];
};

struct CellDesc {           // one FAST cell = one sub-image the reference hands to cv::FAST
    short level;
    short x0, y0;           // sub-image origin in level pixels (iniX, iniY)
    short cw, ch;           // sub-image size (maxX-iniX, maxY-iniY)
    short ci, cj;           // cell row / column
    short addx, addy;       // j*wCell, i*hCell
};

struct TileDesc { short level, x0, y0; };   // blur tiles (kBlurTW x kBlurTH outputs)

constexpr int kBlurTW = 64;
constexpr int kBlurTH = 32;

}  // namespace orbs
