/*
 * oracle/orb_oracle.c -- CPU restatement of the reference ORB front end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Follows /root/reference/SingleRobotScenario/src/ORBextractor.cc (cited per
 * function as ORBextractor.cc:LINE).  The OpenCV primitives the reference calls
 * (cv::resize INTER_LINEAR u8, cv::GaussianBlur 7x7 sigma 2 u8, cv::FAST
 * TYPE_9_16 + NMS, cv::fastAtan2, cvRound) are NOT in /root/reference (OpenCV is
 * an un-vendored system dependency, version only bounded >=2.4.3 / 3.0,
 * CMakeLists.txt:31-37).  They are restated here from OpenCV's published
 * fixed-point algorithms and pinned against cv2 4.13.0 by
 * tests/test_oracle_opencv_pin.py (bit-exact on random images).  Parity of the
 * whole extractor is therefore pinned to "OpenCV 4.13.0 semantics"; the
 * reference repository holds no golden vectors for this path.
 *
 * Defined tie-break: the reference sorts (count, ExtractorNode*) pairs
 * (ORBextractor.cc:684), i.e. ties are broken by heap address.  This oracle
 * breaks ties by node creation order (later-created node first, which is what
 * a bump allocator handing out ascending addresses produces).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORB_PATCH_SIZE 31
#define ORB_HALF_PATCH 15
#define ORB_EDGE 19
#define ORB_MAX_LEVELS 16

static const int8_t brief_pattern[1024] = {
#include "../include/orb_brief_pattern.inc"
};

/* cvRound: round-half-to-even, like _mm_cvtsd_si32 / lrint in default mode */
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_round_f(float v) { return (int)lrintf(v); }

/* ------------------------------------------------------------------ */
/* cv::resize(..., INTER_LINEAR) for CV_8UC1: 11-bit fixed point coefficients
 * (INTER_RESIZE_COEF_BITS), horizontal pass in int32, vertical pass
 * ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.   Called from
 * ORBextractor.cc:1120. */
static void resize_coeffs(int ssize, int dsize, int *idx, short *coef)
{
    double inv_scale = (double)dsize / ssize;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= ssize - 1) { fx = 0; sx = ssize - 1; }
        idx[d] = sx;
        coef[2 * d] = (short)cv_round_f((1.f - fx) * 2048.f);
        coef[2 * d + 1] = (short)cv_round_f(fx * 2048.f);
    }
}

void oracle_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                             uint8_t *dst, int dw, int dh, int dstride)
{
    int *xi = (int *)malloc(sizeof(int) * dw), *yi = (int *)malloc(sizeof(int) * dh);
    short *xa = (short *)malloc(sizeof(short) * 2 * dw), *ya = (short *)malloc(sizeof(short) * 2 * dh);
    int *r0 = (int *)malloc(sizeof(int) * dw), *r1 = (int *)malloc(sizeof(int) * dw);
    resize_coeffs(sw, dw, xi, xa);
    resize_coeffs(sh, dh, yi, ya);
    for (int y = 0; y < dh; y++) {
        int sy0 = yi[y], sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const uint8_t *s0 = src + (size_t)sy0 * sstride, *s1 = src + (size_t)sy1 * sstride;
        for (int x = 0; x < dw; x++) {
            int sx0 = xi[x], sx1 = sx0 + 1 < sw ? sx0 + 1 : sw - 1;
            r0[x] = s0[sx0] * xa[2 * x] + s0[sx1] * xa[2 * x + 1];
            r1[x] = s1[sx0] * xa[2 * x] + s1[sx1] * xa[2 * x + 1];
        }
        int b0 = ya[2 * y], b1 = ya[2 * y + 1];
        for (int x = 0; x < dw; x++) {
            int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)y * dstride + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xi); free(yi); free(xa); free(ya); free(r0); free(r1);
}

/* BORDER_REFLECT_101 index */
static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on CV_8U: OpenCV's
 * bit-exact fixed-point path; the 8.8 kernel is {18,34,48,56,48,34,18}/256,
 * both passes exact, one rounding (v + 2^15) >> 16.  ORBextractor.cc:1086. */
static const int gauss7_q8[7] = {18, 34, 48, 56, 48, 34, 18};

void oracle_gaussian_blur7_u8(const uint8_t *src, int w, int h, int sstride,
                              uint8_t *dst, int dstride)
{
    uint16_t *hz = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int k = 0; k < 7; k++)
                s += gauss7_q8[k] * src[(size_t)y * sstride + reflect101(x + k - 3, w)];
            hz[(size_t)y * w + x] = (uint16_t)s;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int k = 0; k < 7; k++)
                s += (uint32_t)gauss7_q8[k] * hz[(size_t)reflect101(y + k - 3, h) * w + x];
            dst[(size_t)y * dstride + x] = (uint8_t)((s + 32768u) >> 16);
        }
    free(hz);
}

/* ------------------------------------------------------------------ */
/* cv::FAST TYPE_9_16.  A pixel is a corner at threshold t iff 9 contiguous
 * ring pixels are all brighter than v+t or all darker than v-t.  Its
 * response is the largest t for which it is still a corner:
 *   score = max(max_arc min_arc(p-v), max_arc min_arc(v-p)) - 1.
 * NMS keeps a corner iff its score is strictly greater than the scores of
 * all 8 neighbours that are corners too (non-corners count as 0).  Detection
 * runs on rows/cols [3, size-3).  Output order is row-major. */
static const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static int fast_score(const uint8_t *p, int stride)
{
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = (int)p[ring_dy[k] * stride + ring_dx[k]] - v;
    for (int k = 0; k < 9; k++) d[16 + k] = d[k];
    int best = 0;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) {
            if (d[s + k] < mn) mn = d[s + k];
            if (d[s + k] > mx) mx = d[s + k];
        }
        if (mn > best) best = mn;      /* all brighter by at least mn */
        if (-mx > best) best = -mx;    /* all darker by at least -mx */
    }
    return best - 1;                   /* corner at t  <=>  score >= t */
}

/* returns number of keypoints written (x, y, score triplets) */
int oracle_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold,
                    int nms, int *out_xys, int cap)
{
    if (w < 7 || h < 7) return 0;
    int *sc = (int *)calloc((size_t)w * h, sizeof(int));
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = fast_score(img + (size_t)y * stride + x, stride);
            sc[(size_t)y * w + x] = s >= threshold ? s : 0;
        }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sc[(size_t)y * w + x];
            /* threshold 0 corners with score 0 cannot be told from non-corners
             * here; the reference only uses thresholds >= 1 */
            if (s <= 0) continue;
            if (nms) {
                int keep = 1;
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        if (sc[(size_t)(y + dy) * w + x + dx] >= s) { keep = 0; break; }
                    }
                if (!keep) continue;
            }
            if (n < cap) { out_xys[3 * n] = x; out_xys[3 * n + 1] = y; out_xys[3 * n + 2] = s; }
            n++;
        }
    free(sc);
    return n;
}

/* cv::fastAtan2(y, x) scalar path: degree-7 odd polynomial, fp32, no FMA
 * contraction (pinned against cv2.fastAtan2).  ORBextractor.cc:103. */
float oracle_fast_atan2(float y, float x)
{
    const float s = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------ */
/* Extractor parameters and tables: ORBextractor.cc:410-470 */
typedef struct {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor;             /* member is double, ORBextractor.h:98 */
    float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS];
    float sigma2[ORB_MAX_LEVELS], inv_sigma2[ORB_MAX_LEVELS];
    int features_per_level[ORB_MAX_LEVELS];
    int umax[ORB_HALF_PATCH + 1];
} oracle_orb_params;

int oracle_orb_params_init(oracle_orb_params *P, int nfeatures, float scale_factor,
                           int nlevels, int ini_th, int min_th)
{
    if (nlevels < 1 || nlevels > ORB_MAX_LEVELS) return -1;
    P->nfeatures = nfeatures; P->nlevels = nlevels; P->ini_th = ini_th; P->min_th = min_th;
    P->scale_factor = scale_factor;
    P->scale[0] = 1.0f; P->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        P->scale[i] = (float)(P->scale[i - 1] * P->scale_factor);
        P->sigma2[i] = P->scale[i] * P->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        P->inv_scale[i] = 1.0f / P->scale[i];
        P->inv_sigma2[i] = 1.0f / P->sigma2[i];
    }
    float factor = (float)(1.0f / P->scale_factor);
    float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        P->features_per_level[l] = cv_round_f(desired);
        sum += P->features_per_level[l];
        desired *= factor;
    }
    P->features_per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;

    int v, v0;
    int vmax = (int)floorf(ORB_HALF_PATCH * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(ORB_HALF_PATCH * sqrtf(2.f) / 2);
    const double hp2 = ORB_HALF_PATCH * ORB_HALF_PATCH;
    for (v = 0; v <= vmax; ++v) P->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = ORB_HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (P->umax[v0] == P->umax[v0 + 1]) ++v0;
        P->umax[v] = v0;
        ++v0;
    }
    return 0;
}

void oracle_level_size(const oracle_orb_params *P, int w, int h, int level, int *lw, int *lh)
{
    float s = P->inv_scale[level];                       /* ORBextractor.cc:1111-1112 */
    *lw = cv_round_f((float)w * s);
    *lh = cv_round_f((float)h * s);
}

/* ------------------------------------------------------------------ */
/* Per-cell FAST with threshold fallback: ORBextractor.cc:765-829.
 * img = level image (no border), coordinates in level pixels.  Output
 * candidates are relative to (minBorderX, minBorderY) like vToDistributeKeys.
 * Each cell is an independent FAST call on a sub-image, so NMS never sees
 * neighbours across a cell seam. */
typedef struct { float x, y, response; } oracle_cand;

int oracle_detect_cells(const uint8_t *img, int w, int h, int stride, int ini_th, int min_th,
                        oracle_cand *out, int cap)
{
    const int minBX = ORB_EDGE - 3, minBY = minBX;
    const int maxBX = w - ORB_EDGE + 3, maxBY = h - ORB_EDGE + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const float W = 30;
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols < 1 || nRows < 1) return 0;    /* reference would divide by zero */
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    int n = 0;
    int *tmp = (int *)malloc(sizeof(int) * 3 * (size_t)(wCell + 6) * (hCell + 6));
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
            const uint8_t *sub = img + (size_t)y0 * stride + x0;
            int cap_cell = (wCell + 6) * (hCell + 6);
            int k = oracle_fast9_16(sub, cw, ch, stride, ini_th, 1, tmp, cap_cell);
            if (k == 0) k = oracle_fast9_16(sub, cw, ch, stride, min_th, 1, tmp, cap_cell);
            for (int q = 0; q < k; q++) {
                if (n < cap) {
                    out[n].x = (float)(tmp[3 * q] + j * wCell);
                    out[n].y = (float)(tmp[3 * q + 1] + i * hCell);
                    out[n].response = (float)tmp[3 * q + 2];
                }
                n++;
            }
        }
    }
    free(tmp);
    return n;
}

/* ------------------------------------------------------------------ */
/* DistributeOctTree / DivideNode: ORBextractor.cc:481-763.  The std::list is
 * restated as an index-linked list over an arena; a node's arena index is its
 * creation sequence number (used for the documented tie-break). */
typedef struct {
    int ulx, uly, urx, ury, blx, bly, brx, bry;
    int *keys; int nkeys;
    int prev, next;          /* list links, -1 = none */
    int no_more;
} onode;

typedef struct {
    onode *a; int n, cap;
    int head, tail, size;
} olist;

static int ol_new(olist *L)
{
    if (L->n == L->cap) { L->cap = L->cap ? L->cap * 2 : 256; L->a = (onode *)realloc(L->a, sizeof(onode) * L->cap); }
    memset(&L->a[L->n], 0, sizeof(onode));
    L->a[L->n].prev = L->a[L->n].next = -1;
    return L->n++;
}
static void ol_push_front(olist *L, int i)
{
    L->a[i].prev = -1; L->a[i].next = L->head;
    if (L->head >= 0) L->a[L->head].prev = i; else L->tail = i;
    L->head = i; L->size++;
}
static void ol_push_back(olist *L, int i)
{
    L->a[i].next = -1; L->a[i].prev = L->tail;
    if (L->tail >= 0) L->a[L->tail].next = i; else L->head = i;
    L->tail = i; L->size++;
}
static int ol_erase(olist *L, int i)       /* returns next */
{
    int p = L->a[i].prev, n = L->a[i].next;
    if (p >= 0) L->a[p].next = n; else L->head = n;
    if (n >= 0) L->a[n].prev = p; else L->tail = p;
    L->size--;
    free(L->a[i].keys); L->a[i].keys = NULL;
    return n;
}

/* DivideNode, ORBextractor.cc:481-537; children are created in the arena
 * (detached), returns their indices in c[4] */
static void divide_node(olist *L, int ni, const oracle_cand *K, int c[4])
{
    for (int q = 0; q < 4; q++) c[q] = ol_new(L);
    onode *n = &L->a[ni];
    onode *n1 = &L->a[c[0]], *n2 = &L->a[c[1]], *n3 = &L->a[c[2]], *n4 = &L->a[c[3]];
    const int halfX = (int)ceilf((float)(n->urx - n->ulx) / 2);
    const int halfY = (int)ceilf((float)(n->bry - n->uly) / 2);
    n1->ulx = n->ulx; n1->uly = n->uly;
    n1->urx = n->ulx + halfX; n1->ury = n->uly;
    n1->blx = n->ulx; n1->bly = n->uly + halfY;
    n1->brx = n->ulx + halfX; n1->bry = n->uly + halfY;
    n2->ulx = n1->urx; n2->uly = n1->ury;
    n2->urx = n->urx; n2->ury = n->ury;
    n2->blx = n1->brx; n2->bly = n1->bry;
    n2->brx = n->urx; n2->bry = n->uly + halfY;
    n3->ulx = n1->blx; n3->uly = n1->bly;
    n3->urx = n1->brx; n3->ury = n1->bry;
    n3->blx = n->blx; n3->bly = n->bly;
    n3->brx = n1->brx; n3->bry = n->bly;
    n4->ulx = n3->urx; n4->uly = n3->ury;
    n4->urx = n2->brx; n4->ury = n2->bry;
    n4->blx = n3->brx; n4->bly = n3->bry;
    n4->brx = n->brx; n4->bry = n->bry;
    for (int q = 0; q < 4; q++) { L->a[c[q]].keys = (int *)malloc(sizeof(int) * (n->nkeys ? n->nkeys : 1)); }
    for (int i = 0; i < n->nkeys; i++) {
        const oracle_cand *kp = &K[n->keys[i]];
        onode *t;
        if (kp->x < n1->urx) t = (kp->y < n1->bry) ? n1 : n3;
        else t = (kp->y < n1->bry) ? n2 : n4;
        t->keys[t->nkeys++] = n->keys[i];
    }
    for (int q = 0; q < 4; q++) if (L->a[c[q]].nkeys == 1) L->a[c[q]].no_more = 1;
}

typedef struct { int count, node; } size_node;
static int cmp_size_node(const void *pa, const void *pb)
{
    const size_node *a = (const size_node *)pa, *b = (const size_node *)pb;
    if (a->count != b->count) return a->count < b->count ? -1 : 1;
    return a->node < b->node ? -1 : (a->node > b->node);   /* creation order stands in for the address */
}

/* returns number of selected keypoints; sel[] receives indices into K in
 * list order (ORBextractor.cc:741-761) */
int oracle_distribute_octree(const oracle_cand *K, int nK, int minX, int maxX, int minY, int maxY,
                             int N, int *sel, int cap)
{
    if (maxY - minY <= 0) return -1;
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return -1;                 /* reference indexes an empty vector here */
    const float hX = (float)(maxX - minX) / nIni;
    olist L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    int *roots = (int *)malloc(sizeof(int) * nIni);
    for (int i = 0; i < nIni; i++) {
        int ni = ol_new(&L);
        onode *n = &L.a[ni];
        n->ulx = (int)(hX * (float)i); n->uly = 0;
        n->urx = (int)(hX * (float)(i + 1)); n->ury = 0;
        n->blx = n->ulx; n->bly = maxY - minY;
        n->brx = n->urx; n->bry = maxY - minY;
        n->keys = (int *)malloc(sizeof(int) * (nK ? nK : 1));
        ol_push_back(&L, ni);
        roots[i] = ni;
    }
    for (int i = 0; i < nK; i++) {
        int r = (int)(K[i].x / hX);
        if (r < 0 || r >= nIni) { r = r < 0 ? 0 : nIni - 1; }   /* unreachable for extractor inputs */
        onode *n = &L.a[roots[r]];
        n->keys[n->nkeys++] = i;
    }
    free(roots);
    for (int it = L.head; it >= 0;) {
        if (L.a[it].nkeys == 1) { L.a[it].no_more = 1; it = L.a[it].next; }
        else if (L.a[it].nkeys == 0) it = ol_erase(&L, it);
        else it = L.a[it].next;
    }
    int finish = 0;
    size_node *vs = NULL, *vprev = NULL; int nvs = 0, capvs = 0, nprev = 0, capprev = 0;
#define VS_PUSH(cnt, nd) do { if (nvs == capvs) { capvs = capvs ? capvs * 2 : 256; vs = (size_node *)realloc(vs, sizeof(size_node) * capvs); } \
                              vs[nvs].count = (cnt); vs[nvs].node = (nd); nvs++; } while (0)
    while (!finish) {
        int prevSize = L.size;
        int nToExpand = 0;
        nvs = 0;
        for (int it = L.head; it >= 0;) {
            if (L.a[it].no_more) { it = L.a[it].next; continue; }
            int c[4];
            divide_node(&L, it, K, c);
            for (int q = 0; q < 4; q++) {
                if (L.a[c[q]].nkeys > 0) {
                    ol_push_front(&L, c[q]);
                    if (L.a[c[q]].nkeys > 1) { nToExpand++; VS_PUSH(L.a[c[q]].nkeys, c[q]); }
                } else { free(L.a[c[q]].keys); L.a[c[q]].keys = NULL; }
            }
            it = ol_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) finish = 1;
        else if (L.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                if (nvs > capprev) { capprev = nvs; vprev = (size_node *)realloc(vprev, sizeof(size_node) * capprev); }
                memcpy(vprev, vs, sizeof(size_node) * nvs); nprev = nvs; nvs = 0;
                qsort(vprev, nprev, sizeof(size_node), cmp_size_node);
                for (int j = nprev - 1; j >= 0; j--) {
                    int c[4];
                    divide_node(&L, vprev[j].node, K, c);
                    for (int q = 0; q < 4; q++) {
                        if (L.a[c[q]].nkeys > 0) {
                            ol_push_front(&L, c[q]);
                            if (L.a[c[q]].nkeys > 1) VS_PUSH(L.a[c[q]].nkeys, c[q]);
                        } else { free(L.a[c[q]].keys); L.a[c[q]].keys = NULL; }
                    }
                    ol_erase(&L, vprev[j].node);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
#undef VS_PUSH
    int n = 0;
    for (int it = L.head; it >= 0; it = L.a[it].next) {
        onode *nd = &L.a[it];
        int best = nd->keys[0]; float mr = K[best].response;
        for (int k = 1; k < nd->nkeys; k++)
            if (K[nd->keys[k]].response > mr) { best = nd->keys[k]; mr = K[best].response; }
        if (n < cap) sel[n] = best;
        n++;
    }
    for (int i = 0; i < L.n; i++) free(L.a[i].keys);
    free(L.a); free(vs); free(vprev);
    return n;
}

/* ------------------------------------------------------------------ */
/* IC_Angle moments: ORBextractor.cc:77-104 (image = un-blurred level) */
void oracle_ic_moments(const uint8_t *img, int stride, int x, int y, const int *umax, int *m01, int *m10)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (size_t)y * stride + x;
    for (int u = -ORB_HALF_PATCH; u <= ORB_HALF_PATCH; ++u) m_10 += u * center[u];
    for (int v = 1; v <= ORB_HALF_PATCH; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * stride], val_minus = center[u - v * stride];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    *m01 = m_01; *m10 = m_10;
}

/* computeOrbDescriptor: ORBextractor.cc:108-147 (image = blurred level).
 * cos/sin are evaluated in double on the float angle and cast to float;
 * the rotate expression is fp32 WITHOUT FMA contraction (built with
 * -ffp-contract=off). */
void oracle_orb_descriptor(const uint8_t *img, int stride, int x, int y, float angle_deg, uint8_t *desc)
{
    const float factorPI = (float)(M_PI / 180.f);
    float angle = angle_deg * factorPI;
    float a = (float)cos((double)angle), b = (float)sin((double)angle);
    const uint8_t *center = img + (size_t)y * stride + x;
    const int8_t *pat = brief_pattern;
#define GETV(idx) center[cv_round_f(pat[2 * (idx)] * b + pat[2 * (idx) + 1] * a) * stride + \
                         cv_round_f(pat[2 * (idx)] * a - pat[2 * (idx) + 1] * b)]
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t0 = GETV(2 * k), t1 = GETV(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
#undef GETV
}

/* ------------------------------------------------------------------ */
/* Whole extractor (operator(), ORBextractor.cc:1043-1105) on restated
 * primitives.  Outputs SoA keypoints: x,y (scaled to level 0), angle, response,
 * octave, size; desc N x 32.  Returns N (or <0 on unsupported shape).
 * pyr_out (optional): receives nlevels image pointers (malloc'd, caller frees
 * with oracle_free) of the un-bordered levels for inspection. */
int oracle_orb_extract(const oracle_orb_params *P, const uint8_t *image, int w, int h, int stride,
                       float *kx, float *ky, float *kangle, float *kresp, int *koct, float *ksize,
                       uint8_t *desc, int cap, int *level_counts)
{
    if (w <= 0 || h <= 0) return 0;          /* empty image -> silent return, :1046 */
    const int L = P->nlevels;
    uint8_t *lev[ORB_MAX_LEVELS]; int lw[ORB_MAX_LEVELS], lh[ORB_MAX_LEVELS];
    for (int l = 0; l < L; l++) {
        oracle_level_size(P, w, h, l, &lw[l], &lh[l]);
        if (lw[l] < 1 || lh[l] < 1) { for (int q = 0; q < l; q++) free(lev[q]); return -2; }
        lev[l] = (uint8_t *)malloc((size_t)lw[l] * lh[l]);
        if (l == 0) for (int y = 0; y < h; y++) memcpy(lev[0] + (size_t)y * w, image + (size_t)y * stride, w);
        else oracle_resize_linear_u8(lev[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], lev[l], lw[l], lh[l], lw[l]);
    }
    int n = 0, status = 0;
    for (int l = 0; l < L && status == 0; l++) {
        int ccap = (lw[l] / 2 + 1) * (lh[l] / 2 + 1) + 16;
        oracle_cand *cand = (oracle_cand *)malloc(sizeof(oracle_cand) * ccap);
        int nc = 0;
        const int minB = ORB_EDGE - 3;
        const int maxBX = lw[l] - ORB_EDGE + 3, maxBY = lh[l] - ORB_EDGE + 3;
        int nl = 0;
        if (maxBX - minB >= 30 && maxBY - minB >= 30) {
            nc = oracle_detect_cells(lev[l], lw[l], lh[l], lw[l], P->ini_th, P->min_th, cand, ccap);
            int scap = nc + 8;
            int *sel = (int *)malloc(sizeof(int) * scap);
            nl = oracle_distribute_octree(cand, nc, minB, maxBX, minB, maxBY, P->features_per_level[l], sel, scap);
            if (nl < 0) status = -3;
            else {
                uint8_t *blur = NULL;
                if (nl > 0) { blur = (uint8_t *)malloc((size_t)lw[l] * lh[l]); oracle_gaussian_blur7_u8(lev[l], lw[l], lh[l], lw[l], blur, lw[l]); }
                const int scaledPatch = (int)(ORB_PATCH_SIZE * P->scale[l]);
                for (int i = 0; i < nl; i++) {
                    float px = cand[sel[i]].x + minB, py = cand[sel[i]].y + minB;
                    int ix = cv_round_f(px), iy = cv_round_f(py);
                    int m01, m10;
                    oracle_ic_moments(lev[l], lw[l], ix, iy, P->umax, &m01, &m10);
                    float ang = oracle_fast_atan2((float)m01, (float)m10);
                    if (n < cap) {
                        oracle_orb_descriptor(blur, lw[l], ix, iy, ang, desc + (size_t)n * 32);
                        float sc = P->scale[l];
                        kx[n] = l ? px * sc : px; ky[n] = l ? py * sc : py;
                        kangle[n] = ang; kresp[n] = cand[sel[i]].response; koct[n] = l; ksize[n] = (float)scaledPatch;
                    }
                    n++;
                }
                free(blur);
            }
            free(sel);
        }
        if (level_counts) level_counts[l] = nl;
        free(cand);
    }
    for (int l = 0; l < L; l++) free(lev[l]);
    return status ? status : n;
}
