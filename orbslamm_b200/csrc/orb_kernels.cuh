// orb_kernels.cuh -- sm_100a kernels of the ORB extractor (included by orb_extractor.cu).
#pragma once
//
// Pipeline (one launch each per batch of frames):
//   k_resize_level     ORBextractor::ComputePyramid       ORBextractor.cc:1107-1132  (cv::resize INTER_LINEAR)
//   k_fast_cells       ComputeKeyPointsOctTree cell loop   ORBextractor.cc:765-829    (cv::FAST 9_16 + NMS, th fallback)
//   k_blur7            GaussianBlur 7x7 sigma 2            ORBextractor.cc:1085-1086
//   k_octree           DistributeOctTree / DivideNode      ORBextractor.cc:481-763
//   k_orient_describe  IC_Angle + computeOrbDescriptor     ORBextractor.cc:77-147, 1076-1103
//
// All arithmetic is integer / fixed point except the fp32 angle and rotation, which use
// explicit round-to-nearest intrinsics so that no FMA contraction can change a bit.
#include <cuda.h>
#include "orb_extractor.cuh"

namespace orbs {

// 256 tests x (x0, y0, x1, y1) int8: lane l of a descriptor warp owns tests 8l .. 8l+7 = 32 contiguous bytes, fetched as two
// 16-byte loads into registers (a __constant__ or byte-wise shared copy serialises: lane-varying addresses / 8-way bank conflicts)
__device__ __align__(16) const int8_t g_brief_pattern[1024] = {
#include "../../include/orb_brief_pattern.inc"
};

// ---------------------------------------------------------------------------------------------
// Pyramid: level l from level l-1, OpenCV fixed-point bilinear (11-bit coefficients).
// tab_x / tab_y entries: .x = source index, .y = (a1 << 16) | a0   (a0 + a1 = 2048)
//
// kWords (rows 4-byte aligned and the four source columns of a thread within 8 bytes of the aligned word left of the first one --
// the host checks both, any scale factor <= ~2): each source row arrives as THREE aligned 32-bit loads; the byte pair
// (p[sx], p[sx + 1]) of an output is one funnel shift of two of those words and the horizontal interpolation one dp2a against
// the packed (a0, a1) of the table entry -- 6 loads and ~25 instructions per output pixel instead of 16 byte loads with 64-bit
// address arithmetic each (~57).  Words past the row's last one are clamped to it: the bytes they would have held only ever
// meet a1 = 0 (sx = sw - 1 has fx = 0 in the table, as in cv::resize).  !kWords: byte loads (unaligned caller-owned level 0).
template <bool kWords>
__global__ void __launch_bounds__(256)
k_resize_level(const uint8_t *__restrict__ src, int sw, int sh, int spitch, size_t sframe,
               uint8_t *__restrict__ dst, int dw, int dh, int dpitch, size_t dframe,
               const int2 *__restrict__ tab_x, const int2 *__restrict__ tab_y)
{
    // 4 adjacent output pixels per thread: one pair of row pointers, vector table loads, one 32-bit store
    const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= dw || y >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sframe;
    const int2 ty = __ldg(&tab_y[y]);
    const int sy0 = ty.x, sy1 = min(sy0 + 1, sh - 1);
    const int b0 = ty.y & 0xffff, b1 = ty.y >> 16;
    const uint8_t *r0p = s + (size_t)sy0 * spitch, *r1p = s + (size_t)sy1 * spitch;
    const int4 t01 = __ldg(reinterpret_cast<const int4 *>(tab_x + x)), t23 = __ldg(reinterpret_cast<const int4 *>(tab_x + x + 2));
    const int sxs[4] = {t01.x, t01.z, t23.x, t23.z}, cf[4] = {t01.y, t01.w, t23.y, t23.w};
    unsigned out = 0u;
    if (kWords) {
        // entries beyond dw are table padding (0, 0) or the next table: they only select among registers already loaded, masked below
        const int base = sxs[0] & ~3, lastw = (sw - 1) & ~3;
        const int o0 = base, o1 = min(base + 4, lastw), o2 = min(base + 8, lastw);
        const unsigned u0 = __ldg(reinterpret_cast<const unsigned *>(r0p + o0)), u1 = __ldg(reinterpret_cast<const unsigned *>(r0p + o1)),
                       u2 = __ldg(reinterpret_cast<const unsigned *>(r0p + o2));
        const unsigned v0 = __ldg(reinterpret_cast<const unsigned *>(r1p + o0)), v1 = __ldg(reinterpret_cast<const unsigned *>(r1p + o1)),
                       v2 = __ldg(reinterpret_cast<const unsigned *>(r1p + o2));
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = sxs[k] - base;                                     // 0 .. 7 for a valid output, 0 .. 3 for the first one
            const bool up = k > 0 && o >= 4;
            const unsigned sh8 = (unsigned)(o & 3) * 8u;
            const unsigned p0 = __funnelshift_r(up ? u1 : u0, up ? u2 : u1, sh8);   // byte 0 = row0[sx], byte 1 = row0[sx + 1]
            const unsigned p1 = __funnelshift_r(up ? v1 : v0, up ? v2 : v1, sh8);
            const int r0 = (int)__dp2a_lo((unsigned)cf[k], p0, 0u);            // row0[sx] * a0 + row0[sx + 1] * a1
            const int r1 = (int)__dp2a_lo((unsigned)cf[k], p1, 0u);
            const int v = min((((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2, 255);   // every term is >= 0
            out |= (unsigned)v << (8 * k);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // entries beyond dw are table padding (0, 0) or the next table: harmless reads, masked below
            const int sx0 = min(sxs[k], sw - 1), sx1 = min(sx0 + 1, sw - 1);
            const int a0 = cf[k] & 0xffff, a1 = cf[k] >> 16;
            const int r0 = r0p[sx0] * a0 + r0p[sx1] * a1;
            const int r1 = r1p[sx0] * a0 + r1p[sx1] * a1;
            int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            out |= (unsigned)v << (8 * k);
        }
    }
    uint8_t *q = dst + (size_t)blockIdx.z * dframe + (size_t)y * dpitch + x;
    if (x + 3 < dw) *reinterpret_cast<unsigned *>(q) = out;
    else for (int k = 0; k < 4 && x + k < dw; k++) q[k] = (uint8_t)(out >> (8 * k));
}

// ---------------------------------------------------------------------------------------------
// FAST-9/16 on packed pairs of pixels.
//
// For a pixel v and ring p[0..15]:  score = max( max_arc min_arc(p - v), max_arc min_arc(v - p) ) - 1
// over the 16 cyclic arcs of 9 pixels; the pixel is a corner at threshold t iff score >= t, and the
// score is cv::FAST's response.  Two horizontally adjacent pixels live in the two 16-bit lanes of a
// register: A[i] = 256 + v - p[i] per lane (no inter-lane borrow); the sliding min / max of 9 is two levels of the native
// 3-input u16x2 min / max (VIMNMX3.U16x2): m3[k] = min3(A[k..k+2]), m9[k] = min3(m3[k], m3[k+3], m3[k+6]).
//
// One CTA per FAST cell (= one cv::FAST call of the reference, ORBextractor.cc:808-815):
//   0. the cell's sub-image (from the 16-byte aligned column left of it) is fetched by ONE TMA box load
//      (cp.async.bulk.tensor.3d -> UTMALDG) on a per-level (x, y, frame) tensor map, completing on an mbarrier; no thread touches
//      global pixels.  The box must start on a 16-byte boundary in global memory (an unaligned x coordinate raises "illegal
//      instruction", tools/tma_probe/), hence the aligned start + byte shift.  Fallbacks: one descriptor-free bulk copy per row
//      (cp.async.bulk -> UBLKCP) when no tensor map could be encoded, plain loads when the caller's level-0 buffer is unaligned.
//   1. bytes are widened to one pixel per 16-bit lane (pix16), so that ring pixel pairs are 32-bit words (even dx) or one PRMT
//      of two words (odd dx);
//   2. every thread scores 4 adjacent pixels per step from 21 64-bit shared loads;
//   3. cell-local strict 3x3 non-maximum suppression on the 16-bit score plane, the iniTh -> minTh fallback of the cell, and a
//      warp-aggregated append of the survivors.
// Candidates are appended unordered; the reference's candidate order (cell row, cell col, y, x) is carried in the record
// and only matters for response ties inside a quad-tree node.
// record: .x = x | y << 16 (relative to minBorder, as vToDistributeKeys), .y = score | ci << 8 | cj << 18
constexpr int kRawRows = 66;                 // cell sub-images are <= 66 x 66
constexpr int kRawPitchMax = 96;             // bytes per staged row: 16-byte aligned start (<= 15 bytes early) + cw + 1, rounded to 16
constexpr int kPixPitch16 = 96;              // u16 per pix16 row (48 words: consecutive rows start 16 banks apart)
constexpr int kScPitch16 = 68;               // u16 per score row (34 words: every row starts 8-byte aligned for the 64-bit loads of the NMS pass)
constexpr int kScRows = 62;                  // detection region is <= 60 x 60 (+ zero frame)

struct LevelTmaps { CUtensorMap m[ORBS_MAX_LEVELS]; };      // one (x, y, frame) u8 tensor map per pyramid level

__device__ __forceinline__ unsigned hi_lo(unsigned a, unsigned b) { return __byte_perm(a, b, 0x5432); }   // (a.hi, b.lo)

// A[k] = 256 + v - p[k] per lane -> score per lane in [0, 254]
__device__ __forceinline__ unsigned fast_score_lanes(const unsigned (&A)[16])
{
    unsigned n3[16], x3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        n3[k] = __vimin3_u16x2(A[k], A[(k + 1) & 15], A[(k + 2) & 15]);
        x3[k] = __vimax3_u16x2(A[k], A[(k + 1) & 15], A[(k + 2) & 15]);
    }
    unsigned bmn = 0u, bmx = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned na = __vimin3_u16x2(n3[k], n3[(k + 3) & 15], n3[(k + 6) & 15]);
        const unsigned nb = __vimin3_u16x2(n3[k + 1], n3[(k + 4) & 15], n3[(k + 7) & 15]);
        bmn = __vimax3_u16x2(bmn, na, nb);               // max over arcs of min (256 + v - p): darker ring
        const unsigned xa = __vimax3_u16x2(x3[k], x3[(k + 3) & 15], x3[(k + 6) & 15]);
        const unsigned xb = __vimax3_u16x2(x3[k + 1], x3[(k + 4) & 15], x3[(k + 7) & 15]);
        bmx = __vimin3_u16x2(bmx, xa, xb);               // min over arcs of max (256 + v - p): brighter ring
    }
    // lane score = max(bmn - 256, 256 - bmx, 1) - 1
    const unsigned c256 = 0x01000100u;
    const unsigned d = __vmaxu2(bmn, c256) - c256;
    const unsigned e = c256 - __vminu2(bmx, c256);
    return __vimax3_u16x2(d, e, 0x00010001u) - 0x00010001u;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// (row, col) of the linear index tid + 128 * it over a [rows x n] grid, advanced without divisions
struct RowCol {
    int row, col, drow, dcol, n;
    __device__ __forceinline__ RowCol(int tid, int n_) : n(n_) { row = tid / n_; col = tid - row * n_; drow = 128 / n_; dcol = 128 - drow * n_; }
    __device__ __forceinline__ void next() { row += drow; col += dcol; if (col >= n) { col -= n; row++; } }
};

__global__ void __launch_bounds__(128, 8)
k_fast_cells(const __grid_constant__ ExtractPlan plan, const __grid_constant__ LevelTmaps tmaps, const unsigned tmap_levels, const int frame_base,
             const unsigned tma_levels, const CellDesc *__restrict__ cells,
             const uint8_t *__restrict__ img0, int pitch0, size_t frame0,
             const uint8_t *__restrict__ pyr, uint2 *__restrict__ cand, int *__restrict__ cand_count,
             int *__restrict__ err_flag)
{
    __shared__ __align__(128) uint8_t raw[kRawRows * kRawPitchMax];
    __shared__ __align__(16) uint16_t pix16[kRawRows * kPixPitch16];      // sub-image pixel (x, y) at pix16[y][x + 1]
    __shared__ __align__(16) uint16_t sc16[kScRows * kScPitch16];         // score of detection pixel (d, gy) at sc16[gy + 1][d + 2]
    __shared__ __align__(8) unsigned long long mbar;
    const CellDesc cell = cells[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelPlan &L = plan.lv[cell.level];
    const int cw = cell.cw, ch = cell.ch;
    const int dw = cw - 6, dh = ch - 6;
    if (dw <= 0 || dh <= 0) return;
    const int tid = threadIdx.x;
    const bool use_tma = (tma_levels >> cell.level) & 1u;
    const uint8_t *img;
    int pitch;
    if (cell.level == 0) { img = img0 + (size_t)frame * frame0; pitch = pitch0; }
    else { img = pyr + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off; pitch = L.pitch; }
    // staged row r: raw[r * rb + shift + j] = level pixel (x0 - 1 + j, y0 + r), j = 0 .. cw
    const int xa = (cell.x0 - 1) & ~15, shift = (cell.x0 - 1) - xa;
    const bool use_map = (tmap_levels >> cell.level) & 1u;
    const int rb = use_map ? L.fast_box_w : (shift + cw + 1 + 15) & ~15;     // staged row pitch

    // ---- 0. fetch the rows
    if (use_map) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(L.fast_box_w * L.fast_box_h) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(raw)), "l"(reinterpret_cast<unsigned long long>(&tmaps.m[cell.level])), "r"(xa), "r"((int)cell.y0),
                         "r"(frame_base + frame), "r"(smem_u32(&mbar)) : "memory");
        }
        __syncthreads();
    } else if (use_tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid < 32) {
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(ch * rb) : "memory");
            __syncwarp();
            const uint8_t *src = img + (size_t)cell.y0 * pitch + xa;
            for (int r = tid; r < ch; r += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(raw + r * rb)), "l"(src + (size_t)r * pitch), "r"(rb), "r"(smem_u32(&mbar)) : "memory");
        }
    } else {
        // level 0 handed over with a base / pitch that is not 16-byte aligned: plain byte loads
        for (int t = tid; t < ch * rb; t += 128) {
            const int y = t / rb, j = t - y * rb;
            const int x = xa + j;
            raw[t] = x < L.w ? img[(size_t)(cell.y0 + y) * pitch + x] : (uint8_t)0;
        }
    }
    // zero frame of the score plane: rows 0 and dh + 1, the left word (indices 0, 1) and, right of the scores, every word up to the last one the NMS pass
    // reads (word 2 * ngroups + 1; the first of them is rewritten by the scores if it holds a real pixel)
    {
        unsigned *s32 = reinterpret_cast<unsigned *>(sc16);
        constexpr int SW = kScPitch16 / 2;                                  // 34 words per row
        for (int t = tid; t < SW; t += 128) { s32[t] = 0u; s32[(dh + 1) * SW + t] = 0u; }
        const int wr = (dw + 2) >> 1, wl = 2 * ((dw + 3) >> 2) + 1;         // first word holding an index >= dw + 2 (or dw + 1 if dw is odd), last word read
        for (int y = tid; y < dh; y += 128) { s32[(y + 1) * SW] = 0u; for (int w = wr; w <= wl; w++) s32[(y + 1) * SW + w] = 0u; }
    }
    if (use_map || use_tma) {
        unsigned done = 0;
        for (int spin = 0; !done && spin < (1 << 16); spin++)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        if (!__syncthreads_and((int)done)) { if (tid == 0) *err_flag = 3; return; }      // never hang the device
    } else {
        __syncthreads();
    }

    // ---- 1. widen: bytes shift + 4q .. shift + 4q + 3 of staged row y -> pix16[y][4q .. 4q+3]
    {
        const int wq = (cw + 1 + 3) >> 2;                                   // groups of 4 holding sub-image bytes j = 0 .. cw
        const int rw = rb >> 2, w0 = shift >> 2, sh = 8 * (shift & 3);
        const unsigned *raw32 = reinterpret_cast<const unsigned *>(raw);
        for (RowCol rc(tid, wq); rc.row < ch; rc.next()) {
            const int y = rc.row, q = rc.col;
            const int wi = y * rw + w0 + q;
            const unsigned lo = raw32[wi], hi = raw32[min(wi + 1, y * rw + rw - 1)];
            const unsigned v = __funnelshift_r(lo, hi, sh);
            uint2 o;
            o.x = __byte_perm(v, 0u, 0x4140);                               // (b0, b1) zero-extended
            o.y = __byte_perm(v, 0u, 0x4342);                               // (b2, b3)
            *reinterpret_cast<uint2 *>(pix16 + y * kPixPitch16 + 4 * q) = o;
        }
    }
    __syncthreads();

    // ---- 2. scores: one step = 4 adjacent detection pixels d = 4g .. 4g+3 of row gy (centres at pix16 index 4 + 4g ..)
    const int ngroups = (dw + 3) >> 2;
    for (RowCol rc(tid, ngroups); rc.row < dh; rc.next()) {
        const int gy = rc.row, g = rc.col;
        // r[dy][0..5]: words W-2 .. W+3 of row (3 + gy + dy - 3), W = 2 + 2g (the first centre pair)
        unsigned r[7][6];
#pragma unroll
        for (int dy = 0; dy < 7; dy++) {
            const uint2 *row = reinterpret_cast<const uint2 *>(pix16 + (gy + dy) * kPixPitch16 + 4 * g);
            const uint2 a = row[0], b2 = row[1], c = row[2];
            r[dy][0] = a.x; r[dy][1] = a.y; r[dy][2] = b2.x; r[dy][3] = b2.y; r[dy][4] = c.x; r[dy][5] = c.y;
        }
        // odd-offset pairs
        unsigned t3m[3], t3p[3];            // rows dy = -3 / +3: (1,2), (2,3), (3,4)
#pragma unroll
        for (int i = 0; i < 3; i++) { t3m[i] = hi_lo(r[0][1 + i], r[0][2 + i]); t3p[i] = hi_lo(r[6][1 + i], r[6][2 + i]); }
        unsigned s1m[4], s0[4], s1p[4];     // rows dy = -1, 0, +1: (0,1), (1,2), (3,4), (4,5)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int lo = i < 2 ? i : i + 1;
            s1m[i] = hi_lo(r[2][lo], r[2][lo + 1]); s0[i] = hi_lo(r[3][lo], r[3][lo + 1]); s1p[i] = hi_lo(r[4][lo], r[4][lo + 1]);
        }
        unsigned sc[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int c = 2 + p;
            const unsigned V = r[3][c] | 0x01000100u;
            unsigned A[16];
            // ring in cyclic order (dx, dy): (0,3) (1,3) (2,2) (3,1) (3,0) (3,-1) (2,-2) (1,-3) (0,-3) (-1,-3) (-2,-2) (-3,-1) (-3,0) (-3,1) (-2,2) (-1,3)
            A[0] = V - r[6][c];        A[1] = V - t3p[1 + p];     A[2] = V - r[5][c + 1];    A[3] = V - s1p[2 + p];
            A[4] = V - s0[2 + p];      A[5] = V - s1m[2 + p];     A[6] = V - r[1][c + 1];    A[7] = V - t3m[1 + p];
            A[8] = V - r[0][c];        A[9] = V - t3m[p];         A[10] = V - r[1][c - 1];   A[11] = V - s1m[p];
            A[12] = V - s0[p];         A[13] = V - s1p[p];        A[14] = V - r[5][c - 1];   A[15] = V - t3p[p];
            sc[p] = fast_score_lanes(A);
        }
        // pixels beyond the detection region score 0
        const int rem = dw - 4 * g;
        if (rem < 4) { if (rem < 3) sc[1] = 0u; else sc[1] &= 0xffffu; if (rem < 2) sc[0] &= 0xffffu; }
        unsigned *so = reinterpret_cast<unsigned *>(sc16 + (gy + 1) * kScPitch16 + 4 * g + 2);      // 4-byte aligned only
        so[0] = sc[0]; so[1] = sc[1];
    }
    __syncthreads();

    // ---- 3. cell-local strict 3x3 maxima; threshold of this cell: iniTh if cv::FAST(iniTh, nms) would return at least one
    // keypoint, else minTh (ORBextractor.cc:808-816; the test is on the NMS survivors, so a plateau of equal scores >= iniTh
    // that suppresses itself still triggers the fallback).  One step = the same 4 adjacent pixels as in the scoring pass: six 64-bit loads
    // (three rows of words 2g .. 2g+3), the column maxima of the rows above / below shared by the two centre words; keep bits stay in a register.
    unsigned keepbits = 0u, strongbits = 0u;      // bit 4*it + k: pixel 4g + k is an NMS survivor with score >= minTh / >= iniTh
    {
        constexpr int SW2 = kScPitch16 / 4;                                 // row pitch in uint2
        const unsigned minth2 = (unsigned)plan.min_th * 0x00010001u, inith2 = (unsigned)plan.ini_th * 0x00010001u;
        const unsigned bias2 = 0x80008000u - minth2;                        // lane + bias has bit 15 set iff lane >= minTh (scores <= 254: no carry between lanes)
        int it = 0;
        for (RowCol rc(tid, ngroups); rc.row < dh; rc.next(), it++) {
            const uint2 *up = reinterpret_cast<const uint2 *>(sc16 + rc.row * kScPitch16 + 4 * rc.col);      // score rows gy, gy + 1 (centre), gy + 2
            const uint2 m0 = up[SW2], m1 = up[SW2 + 1];
            const unsigned c1 = m0.y, c2 = m1.x;                            // pixels (4g, 4g+1) and (4g+2, 4g+3)
            if (((__vmaxu2(c1, c2) + bias2) & 0x80008000u) == 0u) continue;
            const uint2 u0 = up[0], u1 = up[1], d0 = up[2 * SW2], d1 = up[2 * SW2 + 1];
            const unsigned v0 = __vmaxu2(u0.x, d0.x), v1 = __vmaxu2(u0.y, d0.y), v2 = __vmaxu2(u1.x, d1.x), v3 = __vmaxu2(u1.y, d1.y);
            const unsigned v12 = hi_lo(v1, v2), m12 = hi_lo(c1, c2);
            const unsigned nb1 = __vimax3_u16x2(__vimax3_u16x2(hi_lo(v0, v1), v1, v12), hi_lo(m0.x, c1), m12);
            const unsigned nb2 = __vimax3_u16x2(__vimax3_u16x2(v12, v2, hi_lo(v2, v3)), m12, hi_lo(c2, m1.y));
            // per lane: keep iff centre > all eight neighbours and centre >= th  <=>  max(neighbours + 1, th, centre) == centre.  The lane-wise difference
            // (never negative, <= 255) plus 0x7fff sets bit 15 exactly where the lane differs.
            const unsigned n1 = nb1 + 0x00010001u, n2 = nb2 + 0x00010001u;
            const unsigned k1 = ~((__vimax3_u16x2(n1, minth2, c1) - c1) + 0x7fff7fffu) & 0x80008000u, k2 = ~((__vimax3_u16x2(n2, minth2, c2) - c2) + 0x7fff7fffu) & 0x80008000u;
            const unsigned s1 = ~((__vimax3_u16x2(n1, inith2, c1) - c1) + 0x7fff7fffu) & 0x80008000u, s2 = ~((__vimax3_u16x2(n2, inith2, c2) - c2) + 0x7fff7fffu) & 0x80008000u;
            const unsigned K = (k1 >> 15) | (k2 >> 13), S = (s1 >> 15) | (s2 >> 13);       // bits 0, 16 (pixels 4g, 4g+1) and 2, 18 (4g+2, 4g+3)
            keepbits |= ((K & 5u) | ((K >> 15) & 10u)) << (4 * it);
            strongbits |= ((S & 5u) | ((S >> 15) & 10u)) << (4 * it);
        }
    }
    // threshold of the cell: iniTh if any survivor reaches it, else the minTh retry
    const bool strong = __syncthreads_or(strongbits != 0u);
    const int th = strong ? plan.ini_th : plan.min_th;
    (void)th;
    unsigned mine = strong ? strongbits : keepbits;

    // ---- append survivors: one atomic per warp, then every thread writes its own (usually 0..2) records
    const int lane = tid & 31;
    const int cnt = __popc(mine);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = atomicAdd(cand_count + frame * plan.nlevels + cell.level, total);
    int pos = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
    uint2 *out = cand + (size_t)frame * plan.cand_frame_entries + L.cand_off;
    while (mine) {
        const int b = __ffs(mine) - 1;
        mine &= mine - 1;
        const int t = tid + 128 * (b >> 2), e = b & 3;
        const int gy = t / ngroups, d = 4 * (t - gy * ngroups) + e;
        const unsigned sc = sc16[(gy + 1) * kScPitch16 + d + 2];
        if (pos < L.cand_cap) {
            const unsigned kx = (unsigned)(d + 3 + cell.addx), ky = (unsigned)(gy + 3 + cell.addy);
            out[pos] = make_uint2(kx | (ky << 16), sc | ((unsigned)cell.ci << 8) | ((unsigned)cell.cj << 18));
        } else {
            *err_flag = 1;
        }
        pos++;
    }
}

// ---------------------------------------------------------------------------------------------
// 7x7 Gaussian, sigma 2, BORDER_REFLECT_101, OpenCV's 8.8 fixed-point kernel {18,34,48,56,48,34,18}:
// both passes exact, single rounding (v + 2^15) >> 16.
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// Tile = 64 x 32 outputs.  Input rows (reflected at the top / bottom of the level) are fetched by the TMA engine, one aligned
// bulk copy per row; the reflected columns of the left / right level border are patched in shared memory.  Horizontal pass:
// 4 outputs per thread with dp4a on funnel-shifted words; vertical pass: 4 outputs per thread with dp2a on vertically packed
// u16 pairs, one 32-bit store.
constexpr int kBlurRP = 112;                  // staged row pitch: 16-byte left margin + up to 96 loaded bytes
constexpr int kBlurIH = kBlurTH + 6;

__global__ void __launch_bounds__(256)
k_blur7(const __grid_constant__ ExtractPlan plan, const __grid_constant__ LevelTmaps tmaps, const unsigned tmap_levels, const int frame_base,
        const unsigned tma_levels, const TileDesc *__restrict__ tiles,
        const uint8_t *__restrict__ img0, int pitch0, size_t frame0,
        const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur)
{
    __shared__ __align__(128) uint8_t raw[kBlurIH * kBlurRP];     // raw[r][16 + x - xa] = level pixel (x, reflect(y0 - 3 + r))
    __shared__ __align__(16) uint16_t mid[kBlurIH * kBlurTW];
    __shared__ __align__(8) unsigned long long mbar;
    const TileDesc t = tiles[blockIdx.x];
    const int frame = blockIdx.y, tid = threadIdx.x;
    const LevelPlan &L = plan.lv[t.level];
    const uint8_t *img;
    int pitch;
    if (t.level == 0) { img = img0 + (size_t)frame * frame0; pitch = pitch0; }
    else { img = pyr + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off; pitch = L.pitch; }
    const bool use_tma = (tma_levels >> t.level) & 1u;
    const int xa = max(t.x0 - 16, 0);                             // first loaded column (16-byte aligned: x0 % 64 == 0)
    const int need = t.x0 + kBlurTW + 3 - xa;                     // columns xa .. x0 + 66
    // never past the row pitch (bulk copies: whole 16-byte units, the pitch is a multiple of 16 on that path)
    const int rb = use_tma ? min((need + 15) & ~15, (pitch - xa) & ~15) : min(need, pitch - xa);
    // with a tensor map: ONE box load of 112 x 38 bytes starting 16 bytes left of xa and 3 rows above the tile; rows / columns
    // outside the level arrive zero-filled and are patched below with their BORDER_REFLECT_101 sources, which lie inside the box
    const bool use_map = (tmap_levels >> t.level) & 1u;
    if (use_map) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(kBlurIH * kBlurRP) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(raw)), "l"(reinterpret_cast<unsigned long long>(&tmaps.m[t.level])), "r"(xa - 16), "r"((int)t.y0 - 3),
                         "r"(frame_base + frame), "r"(smem_u32(&mbar)) : "memory");
        }
        __syncthreads();
        unsigned done = 0;
        for (int spin = 0; !done && spin < (1 << 16); spin++)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        if (!__syncthreads_and((int)done)) return;                // cannot happen; never hang the device
        // reflected rows: y = -1..-3 above the level (first tile row), y = h..h+2 below it
        if (t.y0 == 0 || t.y0 + kBlurTH + 3 > L.h) {
            if (tid < 6 * (kBlurRP / 4)) {
                const int k = tid / (kBlurRP / 4), wq = tid - k * (kBlurRP / 4);
                const int sy = k < 3 ? -1 - k : L.h + (k - 3);                     // row outside the level
                const int r = sy - (t.y0 - 3), rs = reflect101(sy, L.h) - (t.y0 - 3);   // its box row and the box row of its source
                if (r >= 0 && r < kBlurIH && rs >= 0 && rs < kBlurIH)
                    reinterpret_cast<unsigned *>(raw + r * kBlurRP)[wq] = reinterpret_cast<const unsigned *>(raw + rs * kBlurRP)[wq];
            }
            __syncthreads();
        }
    } else if (use_tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(kBlurIH * rb) : "memory");
        __syncthreads();                                          // barrier initialised and armed before any copy is issued
        if (tid < kBlurIH) {
            const int sy = reflect101(t.y0 + tid - 3, L.h);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(raw + tid * kBlurRP + 16)), "l"(img + (size_t)sy * pitch + xa), "r"(rb), "r"(smem_u32(&mbar)) : "memory");
        }
        unsigned done = 0;
        for (int spin = 0; !done && spin < (1 << 16); spin++)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        if (!__syncthreads_and((int)done)) return;                // cannot happen; never hang the device
    } else {
        for (int i = tid; i < kBlurIH * rb; i += 256) {
            const int r = i / rb, j = i - r * rb;
            raw[r * kBlurRP + 16 + j] = img[(size_t)reflect101(t.y0 + r - 3, L.h) * pitch + xa + j];
        }
        __syncthreads();
    }
    // reflected columns (BORDER_REFLECT_101): x = -1..-3 at the left border, x = w..w+2 at the right border
    if (tid < kBlurIH * 3) {
        const int r = tid / 3, k = tid - r * 3 + 1;
        uint8_t *row = raw + r * kBlurRP + 16 - xa;               // row[x] = pixel x
        if (t.x0 == 0) row[-k] = row[reflect101(-k, L.w)];
        const int xr = L.w - 1 + k;
        if (xr <= t.x0 + kBlurTW + 2 && xr >= t.x0 - 3) row[xr] = row[reflect101(xr, L.w)];
    }
    __syncthreads();
    // horizontal pass: outputs x0 + 4g .. x0 + 4g + 3 of row y need bytes s .. s + 9, s = 16 + x0 - xa - 3 + 4g  (s % 4 == 1)
    constexpr unsigned c0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), c1 = 48u | (34u << 8) | (18u << 16);
    const int s0 = 16 + t.x0 - xa - 3;
    for (int i = tid; i < kBlurIH * (kBlurTW / 4); i += 256) {
        const int y = i >> 4, g = i & 15;
        const unsigned *row = reinterpret_cast<const unsigned *>(raw + y * kBlurRP + ((s0 + 4 * g) & ~3));
        const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
        unsigned o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned lo = k < 3 ? __funnelshift_r(w0, w1, 8 * (k + 1)) : w1;
            const unsigned hi = k < 3 ? __funnelshift_r(w1, w2, 8 * (k + 1)) : w2;
            o[k] = __dp4a(lo, c0, __dp4a(hi, c1, 0u));
        }
        *reinterpret_cast<uint2 *>(mid + y * kBlurTW + 4 * g) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
    }
    __syncthreads();
    // vertical pass: 4 adjacent outputs per thread; rows are packed in vertical pairs for dp2a (two taps per instruction)
    uint8_t *dst = blur + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off;   // blurred slab has the same layout
    constexpr unsigned k01 = 18u | (34u << 8), k23 = 48u | (56u << 8), k45 = 48u | (34u << 8), k6 = 18u;
    for (int i = tid; i < kBlurTH * (kBlurTW / 4); i += 256) {
        const int y = i >> 4, g = i & 15;
        const int ox = t.x0 + 4 * g, oy = t.y0 + y;
        if (oy >= L.h || ox >= L.w) continue;
        const uint2 *p = reinterpret_cast<const uint2 *>(mid + y * kBlurTW + 4 * g);
        constexpr int S = kBlurTW / 4;                            // row stride in uint2
        uint2 m[7];
#pragma unroll
        for (int r = 0; r < 7; r++) m[r] = p[r * S];
        unsigned out = 0u;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            unsigned w[7];
#pragma unroll
            for (int r = 0; r < 7; r++) w[r] = h ? m[r].y : m[r].x;
            unsigned lo = __dp2a_lo(__byte_perm(w[0], w[1], 0x5410), k01, 32768u);
            unsigned hi = __dp2a_lo(__byte_perm(w[0], w[1], 0x7632), k01, 32768u);
            lo = __dp2a_lo(__byte_perm(w[2], w[3], 0x5410), k23, lo); hi = __dp2a_lo(__byte_perm(w[2], w[3], 0x7632), k23, hi);
            lo = __dp2a_lo(__byte_perm(w[4], w[5], 0x5410), k45, lo); hi = __dp2a_lo(__byte_perm(w[4], w[5], 0x7632), k45, hi);
            lo = __dp2a_lo(w[6] & 0xffffu, k6, lo);                hi = __dp2a_lo(w[6] >> 16, k6, hi);
            out |= ((lo >> 16) | ((hi >> 16) << 8)) << (16 * h);
        }
        uint8_t *q = dst + (size_t)oy * L.pitch + ox;
        if (ox + 3 < L.w) *reinterpret_cast<unsigned *>(q) = out;
        else { for (int k = 0; k < 4 && ox + k < L.w; k++) q[k] = (uint8_t)(out >> (8 * k)); }
    }
}

// ---------------------------------------------------------------------------------------------
// Quad-tree distribution.  One CTA per (frame, level).  The std::list<ExtractorNode> of the reference is
// represented by its order: node arrays are indexed by list position and rebuilt every round.
//   round, phase 1: every node with > 1 key is split, in list order            (ORBextractor.cc:594-653)
//   round, phase 2: nodes are split in order (count desc, creation desc) until  (ORBextractor.cc:655-733)
//                   the list holds >= N nodes
//   children are pushed to the FRONT in creation order; untouched nodes keep their relative order.
struct OctShared {
    short4 *bnd[2];    // ulx, uly, urx, bry
    int *cnt[2];
    int *seq[2];
    int *ccnt;         // [cap*4] child key counts
    int *rank;         // [cap] processing rank of a node to split, -1 otherwise
    int *by_rank;      // [cap] by rank: nonempty children -> exclusive scan (child base)
    int *cum;          // [cap] by rank: inclusive scan of (children - 1)
    int *keep;         // [cap] by position: 1 if node survives the round -> exclusive scan
    int *scratch;      // [blockDim] scan scratch
};

__device__ int block_exclusive_scan(int *data, int n, int *scratch)
{
    // in-place exclusive scan of data[0..n), returns the total.  All threads must call.
    const int nt = blockDim.x, tid = threadIdx.x;
    const int chunk = (n + nt - 1) / nt;
    const int b = min(tid * chunk, n), e = min(b + chunk, n);
    int s = 0;
    for (int i = b; i < e; i++) s += data[i];
    scratch[tid] = s;
    __syncthreads();
    // scan of per-thread sums by warp 0 (nt <= 1024)
    if (tid < 32) {
        int carry = 0;
        for (int base = 0; base < nt; base += 32) {
            int v = scratch[base + tid];
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, inc, d); if (tid >= d) inc += o; }
            scratch[base + tid] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) scratch[nt] = carry;
    }
    __syncthreads();
    int run = scratch[tid];
    for (int i = b; i < e; i++) { int v = data[i]; data[i] = run; run += v; }
    const int total = scratch[nt];
    __syncthreads();
    return total;
}

__device__ __forceinline__ int quadrant_of(short4 b, int x, int y)
{
    const int hx = (b.z - b.x + 1) >> 1, hy = (b.w - b.y + 1) >> 1;   // ceil(d / 2)
    return (x < b.x + hx ? 0 : 1) | (y < b.y + hy ? 0 : 2);
}

__global__ void __launch_bounds__(512)
k_octree(const __grid_constant__ ExtractPlan plan, const uint2 *__restrict__ cand,
         const int *__restrict__ cand_count, unsigned *__restrict__ knode,
         uint2 *__restrict__ lvl_kp, int *__restrict__ lvl_count, int *__restrict__ err_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelPlan &L = plan.lv[level];
    const int cap = L.node_cap;
    const int N = L.nfeat;
    const int tid = threadIdx.x, nt = blockDim.x;
    int *out_count = lvl_count + frame * plan.nlevels + level;
    const int nC = min(cand_count[frame * plan.nlevels + level], L.cand_cap);
    if (nC == 0 || L.cell_count == 0) { if (tid == 0) *out_count = 0; return; }

    OctShared S;
    {
        unsigned char *p = smem_raw;
        S.bnd[0] = (short4 *)p; p += sizeof(short4) * cap;
        S.bnd[1] = (short4 *)p; p += sizeof(short4) * cap;
        S.cnt[0] = (int *)p; p += 4 * cap;  S.cnt[1] = (int *)p; p += 4 * cap;
        S.seq[0] = (int *)p; p += 4 * cap;  S.seq[1] = (int *)p; p += 4 * cap;
        S.ccnt = (int *)p; p += 16 * cap;
        S.rank = (int *)p; p += 4 * cap;
        S.by_rank = (int *)p; p += 4 * cap;
        S.cum = (int *)p; p += 4 * cap;
        S.keep = (int *)p; p += 4 * cap;
        S.scratch = (int *)p;
    }
    __shared__ int s_cut, s_nexp;

    const uint2 *K = cand + (size_t)frame * plan.cand_frame_entries + L.cand_off;
    unsigned *kn = knode + (size_t)frame * plan.cand_frame_entries + L.cand_off;

    // ---- roots (ORBextractor.cc:542-584)
    const int nIni = L.n_ini;
    const float hX = L.hx;
    int cur = 0;
    for (int i = tid; i < nIni; i += nt) {
        S.bnd[0][i] = make_short4((short)(int)__fmul_rn(hX, (float)i), 0, (short)(int)__fmul_rn(hX, (float)(i + 1)), (short)L.bh);
        S.cnt[0][i] = 0;
        S.seq[0][i] = i;
    }
    __syncthreads();
    for (int i = tid; i < nC; i += nt) {
        const int x = K[i].x & 0xffff;
        int r = (int)__fdiv_rn((float)x, hX);
        r = min(max(r, 0), nIni - 1);
        atomicAdd(&S.cnt[0][r], 1);
        kn[i] = (unsigned)r;
    }
    __syncthreads();
    for (int i = tid; i < nIni; i += nt) S.keep[i] = S.cnt[0][i] > 0;
    __syncthreads();
    int n = block_exclusive_scan(S.keep, nIni, S.scratch);
    for (int i = tid; i < nIni; i += nt)
        if (S.cnt[0][i] > 0) { const int np = S.keep[i]; S.bnd[1][np] = S.bnd[0][i]; S.cnt[1][np] = S.cnt[0][i]; S.seq[1][np] = S.seq[0][i]; }
    for (int i = tid; i < nC; i += nt) kn[i] = (unsigned)S.keep[kn[i]];
    __syncthreads();
    cur = 1;
    int seq_counter = nIni;
    int phase = 1;

    // ---- split rounds
    while (true) {
        const int prev_size = n;
        short4 *bnd = S.bnd[cur]; int *cnt = S.cnt[cur]; int *seq = S.seq[cur];
        short4 *nbnd = S.bnd[cur ^ 1]; int *ncnt = S.cnt[cur ^ 1]; int *nseq = S.seq[cur ^ 1];
        for (int i = tid; i < n * 4; i += nt) S.ccnt[i] = 0;
        if (tid == 0) { s_cut = 0x7fffffff; s_nexp = 0; }
        __syncthreads();
        // key pass 1: which child does every key of an expandable node fall into
        for (int i = tid; i < nC; i += nt) {
            const unsigned p = kn[i] & 0xffffffu;
            if (cnt[p] > 1) {
                const unsigned xy = K[i].x;
                const int q = quadrant_of(bnd[p], xy & 0xffff, xy >> 16);
                atomicAdd(&S.ccnt[p * 4 + q], 1);
                kn[i] = p | ((unsigned)q << 24);
            }
        }
        // processing order of the expandable nodes
        int m;
        if (phase == 1) {
            for (int p = tid; p < n; p += nt) S.rank[p] = cnt[p] > 1;
            __syncthreads();
            m = block_exclusive_scan(S.rank, n, S.scratch);
            for (int p = tid; p < n; p += nt) if (!(cnt[p] > 1)) S.rank[p] = -1;
        } else {
            __syncthreads();
            for (int p = tid; p < n; p += nt) {
                int r = -1;
                if (cnt[p] > 1) {
                    r = 0;
                    const int c = cnt[p], s = seq[p];
                    for (int o = 0; o < n; o++) {
                        const int co = cnt[o];
                        if (co > 1 && (co > c || (co == c && seq[o] > s))) r++;
                    }
                }
                S.rank[p] = r;
            }
            __syncthreads();
            // m = number of expandable nodes
            for (int p = tid; p < n; p += nt) S.keep[p] = cnt[p] > 1;
            __syncthreads();
            m = block_exclusive_scan(S.keep, n, S.scratch);
        }
        __syncthreads();
        // by-rank arrays
        for (int p = tid; p < n; p += nt) {
            const int r = S.rank[p];
            if (r >= 0) {
                const int nch = (S.ccnt[p * 4] > 0) + (S.ccnt[p * 4 + 1] > 0) + (S.ccnt[p * 4 + 2] > 0) + (S.ccnt[p * 4 + 3] > 0);
                S.by_rank[r] = nch;
                S.cum[r] = nch - 1;
            }
        }
        __syncthreads();
        const int total_children = block_exclusive_scan(S.by_rank, m, S.scratch);   // child base by rank
        // inclusive cumulative growth: cum[r] = sum_{t <= r} (nch_t - 1)
        {
            // exclusive scan then add own value: keep a copy of own values in keep[]
            for (int r = tid; r < m; r += nt) S.keep[r] = S.cum[r];
            __syncthreads();
            block_exclusive_scan(S.cum, m, S.scratch);
            for (int r = tid; r < m; r += nt) S.cum[r] += S.keep[r];
            __syncthreads();
        }
        int J = m;                                               // number of nodes actually split
        if (phase == 2) {
            for (int r = tid; r < m; r += nt) if (n + S.cum[r] >= N) atomicMin(&s_cut, r + 1);
            __syncthreads();
            J = min(s_cut, m);
        }
        const int C = (J < m) ? S.by_rank[J] : total_children;   // children created this round
#ifdef ORBS_OCT_DEBUG
        if (level == 0 && frame == 0 && tid == 0) {
            printf("round phase=%d n=%d m=%d J=%d C=%d\n", phase, n, m, J, C);
            if (phase == 2) for (int p = 0; p < n; p++) if (S.rank[p] >= 0) printf("  p=%d rank=%d cnt=%d seq=%d\n", p, S.rank[p], cnt[p], seq[p]);
        }
#endif
        // survivors
        for (int p = tid; p < n; p += nt) { const int r = S.rank[p]; S.keep[p] = !(r >= 0 && r < J); }
        __syncthreads();
        const int Kc = block_exclusive_scan(S.keep, n, S.scratch);
        const int n_new = C + Kc;
        if (n_new > cap) { if (tid == 0) { *err_flag = 2; *out_count = 0; } return; }
        // build the new list
        for (int p = tid; p < n; p += nt) {
            const int r = S.rank[p];
            if (r >= 0 && r < J) {
                const short4 b = bnd[p];
                const int hx = (b.z - b.x + 1) >> 1, hy = (b.w - b.y + 1) >> 1;
                int rr = S.by_rank[r];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = S.ccnt[p * 4 + q];
                    if (c > 0) {
                        const int np = C - 1 - rr;
                        short4 nb;
                        nb.x = (q & 1) ? (short)(b.x + hx) : b.x;
                        nb.z = (q & 1) ? b.z : (short)(b.x + hx);
                        nb.y = (q & 2) ? (short)(b.y + hy) : b.y;
                        nb.w = (q & 2) ? b.w : (short)(b.y + hy);
                        nbnd[np] = nb; ncnt[np] = c; nseq[np] = seq_counter + rr;
                        if (c > 1) atomicAdd(&s_nexp, 1);
                        rr++;
                    }
                }
            } else {
                const int np = C + S.keep[p];
                nbnd[np] = bnd[p]; ncnt[np] = cnt[p]; nseq[np] = seq[p];
            }
        }
        // key pass 2: new list position of every key
        for (int i = tid; i < nC; i += nt) {
            const unsigned v = kn[i];
            const unsigned p = v & 0xffffffu;
            const int r = S.rank[p];
            unsigned np;
            if (r >= 0 && r < J) {
                const int q = v >> 24;
                int before = 0;
                for (int qq = 0; qq < q; qq++) before += S.ccnt[p * 4 + qq] > 0;
                np = (unsigned)(C - 1 - (S.by_rank[r] + before));
            } else {
                np = (unsigned)(C + S.keep[p]);
            }
            kn[i] = np;
        }
        __syncthreads();
        const int n_exp = s_nexp;
        seq_counter += C;
        n = n_new;
        cur ^= 1;
        __syncthreads();
        if (n >= N || n == prev_size) break;
        if (phase == 1 && n + 3 * n_exp > N) phase = 2;
    }

    // ---- best key per node: max response, first in the reference's candidate order on ties
    unsigned long long *best = reinterpret_cast<unsigned long long *>(S.ccnt);   // 8 B per node, ccnt has 16
    for (int p = tid; p < n; p += nt) best[p] = 0ull;
    __syncthreads();
    constexpr unsigned long long kOrdMask = (1ull << 52) - 1;
    for (int i = tid; i < nC; i += nt) {
        const uint2 k = K[i];
        const unsigned long long ord = ((unsigned long long)((k.y >> 8) & 0x3ff) << 42) | ((unsigned long long)((k.y >> 18) & 0x3ff) << 32) |
                                       ((unsigned long long)(k.x >> 16) << 16) | (k.x & 0xffff);
        const unsigned long long pri = ((unsigned long long)(k.y & 0xff) << 52) | (~ord & kOrdMask);
        atomicMax(&best[kn[i] & 0xffffffu], pri);
    }
    __syncthreads();
    uint2 *out = lvl_kp + (size_t)frame * plan.lvl_slab + L.kp_off;
    for (int p = tid; p < n; p += nt) {
        const unsigned long long pri = best[p];
        const unsigned long long ord = ~pri & kOrdMask;
        out[p] = make_uint2((unsigned)(ord & 0xffffffffu), (unsigned)(pri >> 52));   // x | y << 16, score
    }
    if (tid == 0) *out_count = n;
}

// ---------------------------------------------------------------------------------------------
// Orientation (IC_Angle on the un-blurred level) + steered BRIEF (on the blurred level), one warp per
// keypoint; also concatenates the levels and scales coordinates (ORBextractor.cc:1076-1103).
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    // cv::fastAtan2 scalar path, fp32, no contraction
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(256)
k_orient_describe(const __grid_constant__ ExtractPlan plan,
                  const uint8_t *__restrict__ img0, int pitch0, size_t frame0,
                  const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur,
                  const uint2 *__restrict__ lvl_kp, const int *__restrict__ lvl_count,
                  float2 *__restrict__ kp_xy, float *__restrict__ kp_angle, float *__restrict__ kp_response,
                  int *__restrict__ kp_octave, float *__restrict__ kp_size, uint8_t *__restrict__ desc,
                  int *__restrict__ counts)
{
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // output slot
    // locate the level of slot o
    int level = -1, pos = 0, total = 0;
    for (int l = 0; l < plan.nlevels; l++) {
        const int c = lvl_count[frame * plan.nlevels + l];
        if (level < 0 && o < total + c) { level = l; pos = o - total; }
        total += c;
    }
    if (o == 0 && lane == 0) counts[frame] = total;
    if (level < 0) return;
    const LevelPlan &L = plan.lv[level];
    const uint2 k = lvl_kp[(size_t)frame * plan.lvl_slab + L.kp_off + pos];
    const int x = (int)(k.x & 0xffff) + kMinBorder, y = (int)(k.x >> 16) + kMinBorder;   // level pixels
    const uint8_t *img;
    int pitch;
    if (level == 0) { img = img0 + (size_t)frame * frame0; pitch = pitch0; }
    else { img = pyr + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off; pitch = L.pitch; }

    // IC_Angle: lane = column u + 15
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - kHalfPatch;
        const uint8_t *c = img + (size_t)y * pitch + x + u;
        const int au = abs(u);
        // all 31 row loads are issued before the first use (the loop is latency-bound otherwise)
        int vals[2 * kHalfPatch + 1];
#pragma unroll
        for (int v = -kHalfPatch; v <= kHalfPatch; v++)
            vals[v + kHalfPatch] = (au <= plan.umax[v < 0 ? -v : v]) ? (int)__ldg(c + (ptrdiff_t)v * pitch) : 0;
        int sum = 0;
#pragma unroll
        for (int v = -kHalfPatch; v <= kHalfPatch; v++) { sum += vals[v + kHalfPatch]; m01 += v * vals[v + kHalfPatch]; }
        m10 = u * sum;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, d);
        m01 += __shfl_xor_sync(0xffffffffu, m01, d);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // steered BRIEF: lane = descriptor byte
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    const float a = (float)cos((double)ang), b = (float)sin((double)ang);
    const uint8_t *bl = blur + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off;
    const uint8_t *center = bl + (size_t)y * L.pitch + x;
    unsigned pw[8];
    {
        const uint4 *gp = reinterpret_cast<const uint4 *>(g_brief_pattern) + 2 * lane;
        const uint4 q0 = __ldg(gp), q1 = __ldg(gp + 1);
        pw[0] = q0.x; pw[1] = q0.y; pw[2] = q0.z; pw[3] = q0.w; pw[4] = q1.x; pw[5] = q1.y; pw[6] = q1.z; pw[7] = q1.w;
    }
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        int tv[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = (float)(int)(int8_t)(pw[t] >> (16 * e)), py = (float)(int)(int8_t)(pw[t] >> (16 * e + 8));
            const int ry = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int rx = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            tv[e] = center[(ptrdiff_t)ry * L.pitch + rx];
        }
        val |= (tv[0] < tv[1]) << t;
    }
    const size_t oi = (size_t)frame * plan.kp_slab + o;
    desc[oi * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        float fx = (float)x, fy = (float)y;
        if (level != 0) { fx = __fmul_rn(fx, L.scale); fy = __fmul_rn(fy, L.scale); }
        kp_xy[oi] = make_float2(fx, fy);
        kp_angle[oi] = angle;
        kp_response[oi] = (float)k.y;
        kp_octave[oi] = level;
        kp_size[oi] = (float)L.patch;
    }
}

// packed host layout (row stride = any) -> 64-byte pitched level-0 buffer; one aligned 16-byte store per thread.
// The upload itself is ONE flat DMA per chunk (a strided 2-D copy of 1241-byte rows runs at a fraction of PCIe speed).
__global__ void __launch_bounds__(256)
k_repitch(const uint8_t *__restrict__ src, int stride, size_t sframe, uint8_t *__restrict__ dst, int pitch, size_t dframe, int width, int height)
{
    const int segs = pitch >> 4;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= segs * height) return;
    const int y = t / segs, sg = t - y * segs;
    const int x = sg << 4;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (x < width) {
        const uint8_t *p = src + (size_t)blockIdx.y * sframe + (size_t)y * stride + x;
        const uintptr_t a0 = (uintptr_t)p & ~(uintptr_t)3;
        const unsigned sh = 8u * (unsigned)((uintptr_t)p & 3);
        const unsigned *w = reinterpret_cast<const unsigned *>(a0);
        const unsigned w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = sh ? w[4] : 0u;
        o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh); o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
    }
    *reinterpret_cast<uint4 *>(dst + (size_t)blockIdx.y * dframe + (size_t)y * pitch + x) = o;
}

// reflect-101 bordered copy of one level (public mvImagePyramid, ORBextractor.cc:1122-1128)
__global__ void k_border_copy(const uint8_t *__restrict__ src, int w, int h, int pitch, uint8_t *__restrict__ dst, int border)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int W = w + 2 * border;
    if (x >= W) return;
    dst[(size_t)y * W + x] = src[(size_t)reflect101(y - border, h) * pitch + reflect101(x - border, w)];
}

}  // namespace orbs
