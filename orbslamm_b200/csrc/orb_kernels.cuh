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
#include "orb_extractor.cuh"

namespace orbs {

__constant__ int8_t c_brief_pattern[1024] = {
#include "../../include/orb_brief_pattern.inc"
};

// ---------------------------------------------------------------------------------------------
// Pyramid: level l from level l-1, OpenCV fixed-point bilinear (11-bit coefficients).
// tab_x / tab_y entries: .x = source index, .y = (a1 << 16) | a0   (a0 + a1 = 2048)
__global__ void __launch_bounds__(256)
k_resize_level(const uint8_t *__restrict__ src, int sw, int sh, int spitch, size_t sframe,
               uint8_t *__restrict__ dst, int dw, int dh, int dpitch, size_t dframe,
               const int2 *__restrict__ tab_x, const int2 *__restrict__ tab_y)
{
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= dw || y >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sframe;
    const int2 tx = __ldg(&tab_x[x]);
    const int2 ty = __ldg(&tab_y[y]);
    const int sx0 = tx.x, sx1 = min(sx0 + 1, sw - 1);
    const int sy0 = ty.x, sy1 = min(sy0 + 1, sh - 1);
    const int a0 = tx.y & 0xffff, a1 = tx.y >> 16;
    const int b0 = ty.y & 0xffff, b1 = ty.y >> 16;
    const uint8_t *r0p = s + (size_t)sy0 * spitch, *r1p = s + (size_t)sy1 * spitch;
    const int r0 = r0p[sx0] * a0 + r0p[sx1] * a1;
    const int r1 = r1p[sx0] * a0 + r1p[sx1] * a1;
    int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    dst[(size_t)blockIdx.z * dframe + (size_t)y * dpitch + x] = (uint8_t)v;
}

// ---------------------------------------------------------------------------------------------
// FAST-9/16 corner score on packed pairs of pixels.
//
// For a pixel v and ring p[0..15]:  score = max( max_arc min_arc(p - v), max_arc min_arc(v - p) ) - 1
// over the 16 cyclic arcs of 9 pixels; the pixel is a corner at threshold t iff score >= t, and the
// score is cv::FAST's response.  Two horizontally adjacent pixels are processed in the two 16-bit
// lanes of a register: A[i] = 256 + v - p[i] per lane (no inter-lane borrow), then sliding min / max
// of 9 with native u16x2 min/max.
constexpr int kPixPitch = 80;    // bytes per smem pixel row (cell sub-images are <= 66 wide)
constexpr int kPixRows = 66;
constexpr int kScPitch = 64;     // detection region is <= 60 wide (+ 2 zero border)
constexpr int kScRows = 62;

__device__ __forceinline__ unsigned pack_pair(unsigned wa, unsigned wb, int o)
{
    // bytes o and o+1 of the 8-byte string (wa, wb) into the two 16-bit lanes, zero extended
    const unsigned sel = o | (o << 4) | ((o + 1) << 8) | ((o + 1) << 12);
    return __byte_perm(wa, wb, sel) & 0x00ff00ffu;
}

// w[r][0..2]: the 12 bytes of pixel row (y - 3 + r) covering columns c-4 .. c+7 (c = first pixel of the
// 4-pixel group).  PAIR 0 = pixels c, c+1; PAIR 1 = pixels c+2, c+3.  Returns both lanes' scores
// clamped to [0, 254] as lo | hi << 16.
template <int PAIR>
__device__ __forceinline__ unsigned fast_score_pair(const unsigned (&w)[7][3])
{
    // ring in cyclic order: (dx, dy)
    constexpr int RDX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int RDY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    unsigned A[16];
    {
        constexpr int ic = 4 + 2 * PAIR;             // byte index of the pair's first centre pixel
        const unsigned V = pack_pair(w[3][ic / 4], w[3][ic / 4 + (ic / 4 < 2 ? 1 : 0)], ic % 4) | 0x01000100u;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int idx = 4 + 2 * PAIR + RDX[k];   // 1 .. 9
            const int r = 3 + RDY[k];
            const int wi = idx / 4;
            const unsigned P = pack_pair(w[r][wi], w[r][wi < 2 ? wi + 1 : wi], idx % 4);
            A[k] = V - P;                            // 256 + v - p per lane, in [1, 511]
        }
    }
    unsigned mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn2[k] = __vminu2(A[k], A[(k + 1) & 15]); mx2[k] = __vmaxu2(A[k], A[(k + 1) & 15]); }
    unsigned mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn4[k] = __vminu2(mn2[k], mn2[(k + 2) & 15]); mx4[k] = __vmaxu2(mx2[k], mx2[(k + 2) & 15]); }
    unsigned best_mn = 0u, best_mx = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const unsigned mn8 = __vminu2(mn4[k], mn4[(k + 4) & 15]);
        const unsigned mx8 = __vmaxu2(mx4[k], mx4[(k + 4) & 15]);
        const unsigned mn9 = __vminu2(mn8, A[(k + 8) & 15]);
        const unsigned mx9 = __vmaxu2(mx8, A[(k + 8) & 15]);
        best_mn = __vmaxu2(best_mn, mn9);            // max over arcs of min (256 + v - p): darker ring
        best_mx = __vminu2(best_mx, mx9);            // min over arcs of max (256 + v - p): brighter ring
    }
    // lane score = max(best_mn - 256, 256 - best_mx) - 1, clamped at 0
    const int lo = max(max((int)(best_mn & 0xffff) - 256, 256 - (int)(best_mx & 0xffff)) - 1, 0);
    const int hi = max(max((int)(best_mn >> 16) - 256, 256 - (int)(best_mx >> 16)) - 1, 0);
    return (unsigned)lo | ((unsigned)hi << 16);
}

// One CTA per FAST cell (= one cv::FAST call of the reference).  Candidates are appended
// unordered; the reference's candidate order (cell row, cell col, y, x) is carried in the record
// and only matters for response ties inside a quad-tree node.
// record: .x = x | y << 16 (relative to minBorder, as vToDistributeKeys), .y = score | ci << 8 | cj << 18
constexpr int kPixWords = kPixPitch / 4;   // 20 words per staged row

__global__ void __launch_bounds__(128)
k_fast_cells(const __grid_constant__ ExtractPlan plan, const CellDesc *__restrict__ cells,
             const uint8_t *__restrict__ img0, int pitch0, size_t frame0,
             const uint8_t *__restrict__ pyr, uint2 *__restrict__ cand, int *__restrict__ cand_count,
             int *__restrict__ err_flag)
{
    __shared__ __align__(16) uint8_t pix[kPixRows * kPixPitch];
    __shared__ __align__(16) uint8_t sc[kScRows * kScPitch];      // scores, 1-px zero frame around the detection region
    __shared__ __align__(16) uint8_t lm[kScRows * kScPitch];      // 1 = strict 3x3 local maximum inside the cell
    const CellDesc cell = cells[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelPlan &L = plan.lv[cell.level];
    const uint8_t *img;
    int pitch;
    if (cell.level == 0) { img = img0 + (size_t)frame * frame0; pitch = pitch0; }
    else { img = pyr + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off; pitch = L.pitch; }
    const int cw = cell.cw, ch = cell.ch;
    const int dw = cw - 6, dh = ch - 6;
    if (dw <= 0 || dh <= 0) return;
    const int tid = threadIdx.x;

    // ---- stage the sub-image with aligned 32-bit loads: pixel (x, y) of the sub-image lives at pix[y][x + 1].
    // smem word w of row y holds sub-image bytes 4w-1 .. 4w+2, i.e. level bytes x0 + 4w - 1 ..; the two aligned global
    // words around that address are funnel-shifted together.  Loads never pass the last byte of the level row.
    {
        const int nwords = (cw + 1 + 3) >> 2;                       // words that contain at least one sub-image byte
        const uint8_t *row_end_word = nullptr;
        for (int t = tid; t < ch * kPixWords; t += 128) {
            const int y = t / kPixWords, w = t - y * kPixWords;
            unsigned v = 0u;
            if (w <= nwords) {
                const uint8_t *rowp = img + (size_t)(cell.y0 + y) * pitch;
                const uint8_t *last = rowp + L.w - 1;               // last valid byte of this level row
                const uint8_t *p = rowp + cell.x0 + 4 * w - 1;      // first byte wanted
                const uintptr_t a0 = (uintptr_t)p & ~(uintptr_t)3;
                const uintptr_t amax = (uintptr_t)last & ~(uintptr_t)3;
                const unsigned lo = *reinterpret_cast<const unsigned *>(a0 <= amax ? a0 : amax);
                const unsigned hi = *reinterpret_cast<const unsigned *>(a0 + 4 <= amax ? a0 + 4 : amax);
                v = __funnelshift_r(lo, hi, 8 * (unsigned)((uintptr_t)p & 3));
            }
            reinterpret_cast<unsigned *>(pix)[t] = v;
            (void)row_end_word;
        }
        for (int t = tid; t < (dh + 2) * (kScPitch / 4); t += 128) { reinterpret_cast<unsigned *>(sc)[t] = 0u; reinterpret_cast<unsigned *>(lm)[t] = 0u; }
    }
    __syncthreads();

    // ---- scores of the detection region [3, cw-3) x [3, ch-3): groups of 4 pixels
    const int ngroups = (dw + 3) >> 2;
    for (int t = tid; t < ngroups * dh; t += 128) {
        const int gy = t / ngroups, gx = t - gy * ngroups;
        const int y = 3 + gy;                 // sub-image row of the centre
        const int c = 4 + 4 * gx;             // smem column of the first centre pixel (x = c - 1)
        unsigned w[7][3];
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const unsigned *row = reinterpret_cast<const unsigned *>(pix + (y - 3 + r) * kPixPitch + c - 4);
            w[r][0] = row[0]; w[r][1] = row[1]; w[r][2] = row[2];
        }
        const unsigned s01 = fast_score_pair<0>(w);
        const unsigned s23 = fast_score_pair<1>(w);
        // detection pixel index dx = 4*gx + k  ->  sc[gy + 1][dx + 1]; pixels beyond the region stay 0
        const int rem = dw - 4 * gx;
        unsigned packed = (s01 & 0xff) | ((rem > 1 ? (s01 >> 16) & 0xff : 0u) << 8) | ((rem > 2 ? s23 & 0xff : 0u) << 16) |
                          ((rem > 3 ? (s23 >> 16) & 0xff : 0u) << 24);
        uint8_t *o = sc + (gy + 1) * kScPitch + 4 * gx + 1;       // byte address = 4*gx + 1 (mod 4 == 1): write bytes
        o[0] = (uint8_t)packed; o[1] = (uint8_t)(packed >> 8); o[2] = (uint8_t)(packed >> 16); o[3] = (uint8_t)(packed >> 24);
    }
    __syncthreads();

    // ---- cell-local strict 3x3 maxima, once; threshold of this cell: iniTh if cv::FAST(iniTh, nms) would return at
    // least one keypoint, else minTh (ORBextractor.cc:808-816; the test is on the NMS survivors, so a plateau of equal
    // scores >= iniTh that suppresses itself still triggers the fallback)
    int any = 0;
    const int xq = tid & 63, yq = tid >> 6;                        // 64 columns x 2 rows per sweep
    for (int y = yq; y < dh; y += 2) {
        if (xq < dw) {
            const uint8_t *p = sc + (y + 1) * kScPitch + xq + 1;
            const int s = p[0];
            if (s >= plan.min_th) {
                const bool keep = s > p[-1] && s > p[1] && s > p[-kScPitch - 1] && s > p[-kScPitch] && s > p[-kScPitch + 1] &&
                                  s > p[kScPitch - 1] && s > p[kScPitch] && s > p[kScPitch + 1];
                if (keep) { lm[(y + 1) * kScPitch + xq + 1] = 1; any |= s >= plan.ini_th; }
            }
        }
    }
    const int th = __syncthreads_or(any) ? plan.ini_th : plan.min_th;

    // ---- append survivors
    int *counter = cand_count + frame * plan.nlevels + cell.level;
    uint2 *out = cand + (size_t)frame * plan.cand_frame_entries + L.cand_off;
    for (int y0 = 0; y0 < dh; y0 += 2) {
        const int y = y0 + yq;
        bool keep = false;
        int s = 0;
        if (y < dh && xq < dw) { s = sc[(y + 1) * kScPitch + xq + 1]; keep = lm[(y + 1) * kScPitch + xq + 1] && s >= th; }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            const int lane = tid & 31;
            int pos = 0;
            if (lane == 0) pos = atomicAdd(counter, __popc(ballot));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(ballot & ((1u << lane) - 1));
            if (keep) {
                if (pos < L.cand_cap) {
                    const unsigned kx = (unsigned)(xq + 3 + cell.addx), ky = (unsigned)(y + 3 + cell.addy);
                    out[pos] = make_uint2(kx | (ky << 16), (unsigned)s | ((unsigned)cell.ci << 8) | ((unsigned)cell.cj << 18));
                } else {
                    *err_flag = 1;
                }
            }
        }
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

__global__ void __launch_bounds__(256)
k_blur7(const __grid_constant__ ExtractPlan plan, const TileDesc *__restrict__ tiles,
        const uint8_t *__restrict__ img0, int pitch0, size_t frame0,
        const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur)
{
    constexpr int IW = kBlurTW + 6, IH = kBlurTH + 6, IP = 80;    // 70 x 38 input, 80-byte smem rows (word aligned)
    __shared__ __align__(16) uint8_t tin[IH * IP];
    __shared__ __align__(16) uint16_t mid[IH * kBlurTW];
    const TileDesc t = tiles[blockIdx.x];
    const int frame = blockIdx.y, tid = threadIdx.x;
    const LevelPlan &L = plan.lv[t.level];
    const uint8_t *img;
    int pitch;
    if (t.level == 0) { img = img0 + (size_t)frame * frame0; pitch = pitch0; }
    else { img = pyr + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off; pitch = L.pitch; }
    // stage: thread = (column, row phase); the reflected source column is computed once
    {
        const int x = tid % 80, ph = tid / 80;                    // 3 row phases, 240 active threads
        if (ph < 3) {
            const int sx = x < IW ? reflect101(t.x0 + x - 3, L.w) : 0;
            for (int y = ph; y < IH; y += 3) {
                const int sy = reflect101(t.y0 + y - 3, L.h);
                tin[y * IP + x] = x < IW ? img[(size_t)sy * pitch + sx] : (uint8_t)0;
            }
        }
    }
    __syncthreads();
    // horizontal pass, 4 outputs per thread: two dp4a per output on funnel-shifted words
    constexpr unsigned c0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), c1 = 48u | (34u << 8) | (18u << 16);
    for (int i = tid; i < IH * (kBlurTW / 4); i += 256) {
        const int y = i >> 4, g = i & 15;
        const unsigned *row = reinterpret_cast<const unsigned *>(tin + y * IP) + g;
        const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
        unsigned o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned lo = __funnelshift_r(w0, w1, 8 * k), hi = __funnelshift_r(w1, w2, 8 * k);
            o[k] = __dp4a(lo, c0, __dp4a(hi, c1, 0u));
        }
        uint2 pk = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
        *reinterpret_cast<uint2 *>(mid + y * kBlurTW + 4 * g) = pk;
    }
    __syncthreads();
    // vertical pass, 2 adjacent outputs per thread (one 32-bit load of two u16 per tap)
    uint8_t *dst = blur + (size_t)frame * plan.pyr_frame_bytes + L.pyr_off;   // blurred slab has the same layout
    for (int i = tid; i < kBlurTH * (kBlurTW / 2); i += 256) {
        const int y = i >> 5, xp = (i & 31) * 2;
        const unsigned *p = reinterpret_cast<const unsigned *>(mid + y * kBlurTW + xp);
        constexpr int S = kBlurTW / 2;                            // row stride in 32-bit words
        const unsigned m0 = p[0], m1 = p[S], m2 = p[2 * S], m3 = p[3 * S], m4 = p[4 * S], m5 = p[5 * S], m6 = p[6 * S];
        const unsigned lo = 18u * ((m0 & 0xffff) + (m6 & 0xffff)) + 34u * ((m1 & 0xffff) + (m5 & 0xffff)) + 48u * ((m2 & 0xffff) + (m4 & 0xffff)) + 56u * (m3 & 0xffff);
        const unsigned hi = 18u * ((m0 >> 16) + (m6 >> 16)) + 34u * ((m1 >> 16) + (m5 >> 16)) + 48u * ((m2 >> 16) + (m4 >> 16)) + 56u * (m3 >> 16);
        const int ox = t.x0 + xp, oy = t.y0 + y;
        if (oy < L.h) {
            if (ox < L.w) dst[(size_t)oy * L.pitch + ox] = (uint8_t)((lo + 32768u) >> 16);
            if (ox + 1 < L.w) dst[(size_t)oy * L.pitch + ox + 1] = (uint8_t)((hi + 32768u) >> 16);
        }
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
    __shared__ int8_t pat[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) pat[i] = c_brief_pattern[i];
    __syncthreads();
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
#pragma unroll 1
        for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
            if (au <= plan.umax[abs(v)]) {
                const int val = c[(ptrdiff_t)v * pitch];
                m10 += u * val;
                m01 += v * val;
            }
        }
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
    const int8_t *pp = pat + lane * 32;
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        int tv[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = (float)pp[4 * t + 2 * e], py = (float)pp[4 * t + 2 * e + 1];
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

// reflect-101 bordered copy of one level (public mvImagePyramid, ORBextractor.cc:1122-1128)
__global__ void k_border_copy(const uint8_t *__restrict__ src, int w, int h, int pitch, uint8_t *__restrict__ dst, int border)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int W = w + 2 * border;
    if (x >= W) return;
    dst[(size_t)y * W + x] = src[(size_t)reflect101(y - border, h) * pitch + reflect101(x - border, w)];
}

}  // namespace orbs
