// optimizer.cu -- host driver and C-ABI of the optimizer (orbo_*): PoseOptimization (device-resident LM, one CTA
// per frame) and Local / Global bundle adjustment (host-driven Levenberg-Marquardt over CUDA kernels: per-point
// Schur complement, dense Cholesky of the reduced pose system, back-substitution).
//
// Replaces S/src/Optimizer.cc:68-260, 262-474, 476-801 and the g2o stack behind it.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>
#include <chrono>
#include <climits>
#include <dlfcn.h>
#include <nccl.h>
#include "common.cuh"
#include "pose_opt.cuh"
#include "sim3_opt.cuh"
#include "pose_graph.cuh"
#include "ba_kernels.cuh"

using namespace orbs;

// NCCL is bound at run time (dlopen by soname) so that single-GPU users need no NCCL at all
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        CommAbort = (decltype(CommAbort))dlsym(lib, "ncclCommAbort");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && GetErrorString;
    }
};
NcclApi g_nccl;
}  // namespace

#define ORBS_NCCL(call)                                                                        \
    do {                                                                                       \
        ncclResult_t _r = (call);                                                              \
        if (_r != ncclSuccess) { set_last_error(std::string("NCCL error: ") + g_nccl.GetErrorString(_r) + " in " #call); return ORBS_E_CUDA; } \
    } while (0)

struct orbo_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    long long launches = 0;
    std::mutex mu;
    StagePool pool;          // pose optimisation staging
    StagePool ba_pool;       // bundle adjustment buffers
    DevBuf ba_tasks;         // panel / update task lists of the tiled Cholesky (sized by the symbolic factorisation)
    DevBuf ba_packed;        // sharded solve: the structurally nonzero tiles of the reduced system, contiguous, for the all-reduce
    PinnedBuf h_scalars;
    KernelTimer timer;       // BA kernels, ids = BaK
    ncclComm_t comm = nullptr;   // set by orbo_comm_init: orbo_bundle_adjust becomes a collective over map-point shards
    int nranks = 1, rank = 0;
    long long ba_skyline[3] = {0, 0, 0};
    double ba_timing[4] = {0, 0, 0, 0};   // last BA call: LM-loop seconds, total seconds, setup (layout + H2D) seconds, Schur bytes
};

enum BaK { BK_ERRORS = 0, BK_BUILD_POINTS, BK_BUILD_POSES, BK_SCHUR, BK_POTRF, BK_TRSM, BK_SYRK, BK_TRS, BK_BACKSUB, BK_UPDATE, BK_MEMSET, BK_COUNT };

extern "C" {

int orbo_create(orbo_handle **out, int device)
{
    ORBS_REQUIRE(out, ORBS_E_INVALID, "orbo_create: null out pointer");
    *out = nullptr;
    ORBS_CUDA(cudaSetDevice(device));
    orbo_handle *h = new orbo_handle();
    h->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    if (int rc = h->h_scalars.reserve(256)) { orbo_destroy(h); return rc; }
    cudaFuncSetAttribute(k_chol_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPanelSmem);
    cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpdateSmem);
    *out = h;
    return ORBS_OK;
}

int orbo_destroy(orbo_handle *h)
{
    if (!h) return ORBS_OK;
    cudaSetDevice(h->device);
    if (h->stream && h->own_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    else cudaDeviceSynchronize();
    if (h->comm) { if (g_nccl.CommAbort) g_nccl.CommAbort(h->comm); else g_nccl.CommDestroy(h->comm); }   // abort: never block on a peer at teardown
    h->pool.release(); h->ba_pool.release(); h->ba_tasks.release(); h->ba_packed.release(); h->h_scalars.release(); h->timer.release();
    delete h;
    return ORBS_OK;
}

void *orbo_stream(orbo_handle *h) { return h ? (void *)h->stream : nullptr; }

int orbo_set_stream(orbo_handle *h, void *stream)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)stream; h->own_stream = false;
    return ORBS_OK;
}

int orbo_synchronize(orbo_handle *h)
{
    ORBS_REQUIRE(h, ORBS_E_INVALID, "null handle");
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    return ORBS_OK;
}

long long orbo_kernel_launches(const orbo_handle *h) { return h ? h->launches : 0; }

int orbo_comm_unique_id(uint8_t *id128)
{
    ORBS_REQUIRE(id128, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(g_nccl.load(), ORBS_E_CUDA, "libnccl.so.2 not found");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ORBS_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return ORBS_OK;
}

int orbo_comm_init(orbo_handle *h, int nranks, int rank, const uint8_t *id128)
{
    ORBS_REQUIRE(h && id128 && nranks >= 1 && rank >= 0 && rank < nranks, ORBS_E_INVALID, "bad argument");
    ORBS_REQUIRE(g_nccl.load(), ORBS_E_CUDA, "libnccl.so.2 not found");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ORBS_NCCL(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->nranks = nranks; h->rank = rank;
    return ORBS_OK;
}

int orbo_pose_optimization(orbo_handle *h, int n_frames, float *Tcw, const float *K4, const float *Xw, const float *obs,
                           const float *inv_sigma2, const int32_t *counts, int slab, uint8_t *outlier, int32_t *n_inliers,
                           int memspace)
{
    ORBS_REQUIRE(h && Tcw && K4 && Xw && obs && inv_sigma2 && counts && outlier && n_inliers, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && slab > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t ne = (size_t)n_frames * slab;
    PoseArgs A;
    A.slab = slab;
    A.fx = K4[0]; A.fy = K4[1]; A.cx = K4[2]; A.cy = K4[3];          // float members of Frame, promoted to double by the edge
    A.Tcw = S.inout(Tcw, (size_t)n_frames * 16);
    A.Xw = S.in(Xw, ne * 3); A.obs = S.in(obs, ne * 2); A.w = S.in(inv_sigma2, ne);
    A.counts = S.in(counts, n_frames);
    A.outlier = S.inout(outlier, ne, false); A.n_inliers = S.inout(n_inliers, n_frames, false);
    A.err = S.scratch<double>(ne * 2);
    if (S.rc) return S.rc;
    k_pose_optimization<<<n_frames, kPoseThreads, 0, h->stream>>>(A);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbo_pose_optimization_matched(orbo_handle *h, int n_frames, float *Tcw, const float *K4, const float *f_xy, const int32_t *f_octave,
                                   const int32_t *f_counts, int f_slab, const int32_t *feat_match, const float *q_Xw, const int32_t *q_counts,
                                   int q_slab, const float *inv_level_sigma2, int nlevels, uint8_t *f_outlier, int32_t *n_inliers,
                                   int32_t *n_edges, int memspace)
{
    ORBS_REQUIRE(h && Tcw && K4 && f_xy && f_octave && f_counts && feat_match && q_Xw && q_counts && inv_level_sigma2 && f_outlier && n_inliers,
                 ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && f_slab > 0 && q_slab > 0 && nlevels > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t nf = (size_t)n_frames * f_slab, nq = (size_t)n_frames * q_slab;
    PoseGatherArgs G;
    G.f_slab = f_slab; G.q_slab = q_slab; G.nlevels = nlevels;
    G.f_xy = (const float2 *)S.in(f_xy, nf * 2); G.f_octave = S.in(f_octave, nf); G.f_counts = S.in(f_counts, n_frames);
    G.feat_match = S.in(feat_match, nf); G.q_Xw = S.in(q_Xw, nq * 3); G.q_counts = S.in(q_counts, n_frames);
    G.inv_level_sigma2 = S.in(inv_level_sigma2, nlevels);
    G.Xw = S.scratch<float>(nf * 3); G.obs = S.scratch<float>(nf * 2); G.w = S.scratch<float>(nf); G.edge_feat = S.scratch<int>(nf);
    int32_t *d_nedges = n_edges ? S.inout(n_edges, n_frames, false) : S.scratch<int32_t>(n_frames);
    G.counts = d_nedges;
    PoseArgs A;
    A.slab = f_slab;
    A.fx = K4[0]; A.fy = K4[1]; A.cx = K4[2]; A.cy = K4[3];
    A.Tcw = S.inout(Tcw, (size_t)n_frames * 16);
    A.Xw = G.Xw; A.obs = G.obs; A.w = G.w; A.counts = d_nedges;
    uint8_t *d_eout = S.scratch<uint8_t>(nf);
    A.outlier = d_eout; A.n_inliers = S.inout(n_inliers, n_frames, false);
    A.err = S.scratch<double>(nf * 2);
    uint8_t *d_fout = S.inout(f_outlier, nf, false);
    if (S.rc) return S.rc;
    ORBS_CUDA(cudaMemsetAsync(d_fout, 0, nf, h->stream));
    k_pose_gather<<<n_frames, 256, 0, h->stream>>>(G);
    k_pose_optimization<<<n_frames, kPoseThreads, 0, h->stream>>>(A);
    k_pose_scatter<<<dim3((f_slab + 255) / 256, n_frames), 256, 0, h->stream>>>(f_slab, d_nedges, G.edge_feat, d_eout, d_fout);
    h->launches += 3;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbo_optimize_sim3(orbo_handle *h, int n_pairs, double *sim3, const uint8_t *valid, const float *P1c, const float *P2c, const float *obs1,
                       const float *obs2, const float *inv_sigma2_1, const float *inv_sigma2_2, const float *K1, const float *K2, const int32_t *counts,
                       int slab, float th2, int fix_scale, uint8_t *inlier, int32_t *n_inliers, int32_t *lm_stats, int memspace)
{
    ORBS_REQUIRE(h && sim3 && valid && P1c && P2c && obs1 && obs2 && inv_sigma2_1 && inv_sigma2_2 && K1 && K2 && counts && inlier && n_inliers,
                 ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_pairs > 0 && slab > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t ne = (size_t)n_pairs * slab;
    Sim3Args A;
    A.slab = slab;
    A.sim3 = S.inout(sim3, (size_t)n_pairs * 8);
    A.valid = S.in(valid, ne); A.P1c = S.in(P1c, ne * 3); A.P2c = S.in(P2c, ne * 3); A.obs1 = S.in(obs1, ne * 2); A.obs2 = S.in(obs2, ne * 2);
    A.w1 = S.in(inv_sigma2_1, ne); A.w2 = S.in(inv_sigma2_2, ne); A.K1 = S.in(K1, (size_t)n_pairs * 4); A.K2 = S.in(K2, (size_t)n_pairs * 4);
    A.counts = S.in(counts, n_pairs);
    A.th2 = th2; A.fix_scale = fix_scale ? 1 : 0;
    A.inlier = S.inout(inlier, ne, false); A.n_inliers = S.inout(n_inliers, n_pairs, false);
    A.stats = lm_stats ? S.inout(lm_stats, (size_t)n_pairs * 2, false) : nullptr;
    A.err = S.scratch<double>(ne * 4); A.active = S.scratch<uint8_t>(ne);
    if (S.rc) return S.rc;
    if (memspace == ORBS_MEM_HOST) ORBS_CUDA(cudaMemsetAsync(A.inlier, 0, ne, h->stream));
    k_optimize_sim3<<<n_pairs, kSim3Threads, 0, h->stream>>>(A);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

int orbo_sim3_prepare(orbo_handle *h, int N, const float *X3Dc, const int32_t *octave, const float *level_sigma2, int nlevels, const float *K4,
                      int32_t *max_err, float *p2d, int memspace)
{
    ORBS_REQUIRE(h && X3Dc && octave && level_sigma2 && K4 && max_err && p2d, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(N > 0 && nlevels > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const float *dX = S.in(X3Dc, (size_t)N * 3);
    const int32_t *doct = S.in(octave, N);
    float *dls = S.scratch<float>(nlevels);                              // level_sigma2 is a host array in both memory spaces
    int32_t *dme = S.inout(max_err, N, false);
    float *dp = S.inout(p2d, (size_t)N * 2, false);
    if (S.rc) return S.rc;
    ORBS_CUDA(cudaMemcpyAsync(dls, level_sigma2, sizeof(float) * nlevels, cudaMemcpyHostToDevice, h->stream));
    k_sim3_prepare<<<(N + 255) / 256, 256, 0, h->stream>>>(N, dX, doct, dls, nlevels, K4[0], K4[1], K4[2], K4[3], dme, dp);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    if (memspace == ORBS_MEM_DEVICE) ORBS_CUDA(cudaStreamSynchronize(h->stream));
    return S.finish();
}

int orbo_sim3_check_inliers(orbo_handle *h, int n_hyp, const float *T12, const float *T21, int N, const float *X3Dc1, const float *X3Dc2, const float *P1im1,
                            const float *P2im2, const int32_t *max_err1, const int32_t *max_err2, const float *K1, const float *K2, uint8_t *inliers,
                            int32_t *n_inliers, int memspace)
{
    ORBS_REQUIRE(h && T12 && T21 && X3Dc1 && X3Dc2 && P1im1 && P2im2 && max_err1 && max_err2 && K1 && K2 && inliers && n_inliers, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_hyp > 0 && n_hyp <= 65535 && N > 0, ORBS_E_INVALID, "1..65535 hypotheses, at least one correspondence");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    Sim3CheckArgs A;
    A.n_hyp = n_hyp; A.N = N;
    A.T12 = S.in(T12, (size_t)n_hyp * 16); A.T21 = S.in(T21, (size_t)n_hyp * 16);
    A.X1 = S.in(X3Dc1, (size_t)N * 3); A.X2 = S.in(X3Dc2, (size_t)N * 3); A.P1im1 = S.in(P1im1, (size_t)N * 2); A.P2im2 = S.in(P2im2, (size_t)N * 2);
    A.max_err1 = S.in(max_err1, N); A.max_err2 = S.in(max_err2, N);
    for (int k = 0; k < 4; k++) { A.K1[k] = K1[k]; A.K2[k] = K2[k]; }
    A.inliers = S.inout(inliers, (size_t)n_hyp * N, false); A.n_inliers = S.inout(n_inliers, n_hyp, false);
    if (S.rc) return S.rc;
    ORBS_CUDA(cudaMemsetAsync(A.n_inliers, 0, sizeof(int) * n_hyp, h->stream));
    k_sim3_check_inliers<<<dim3((N + 255) / 256, n_hyp), 256, 0, h->stream>>>(A);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Bundle adjustment driver
namespace {

struct BaHost {
    orbo_handle *h;
    cudaStream_t st;
    BaDev B;
    int ntiles = 0;
    int err_blocks = 0, pt_blocks = 0, pose_blocks = 0, diag_blocks = 0, xp_blocks = 0;
    double lambda = 0, ni = 2;
    int nbad = 0;
    const volatile int *stop = nullptr;
    int lm_iterations = 0, lm_trials = 0, chol_failures = 0;
    double *Linv = nullptr;       // [ntiles][64*64] inverses of the diagonal Cholesky tiles
    int *ready = nullptr;         // [2*ntiles] dataflow flags of the triangular solves
    double *h_scal = nullptr;     // pinned [16]
    int *h_flag = nullptr;        // pinned
    CholPlan plan = {};           // sparse tile structure + task lists of the reduced system (device arrays)
    int nt_max = 0;
    int *d_plan_i = nullptr;      // rows_start, rows, cols_start, cols  [2 (nt_max + 1) + nt_max (nt_max - 1)]
    int4 *d_tasks = nullptr;      // panel tasks then update tasks
    int *d_rowbase = nullptr; uint8_t *d_rowpad = nullptr;
    std::vector<int> panel_lv, update_lv;   // per level: first task index (size nlevels + 1)
    int nlevels = 0;
    long long l_tiles = 0;        // structurally nonzero tiles of L (incl. diagonal)
    size_t n_panel = 0;

    // Nested-dissection order of the tile groups by recursive bisection of the natural (temporal) order: the separator of
    // [lo, hi) is the run of groups right of the middle that the left half reaches; left and right halves are then
    // independent and are eliminated in parallel, the separator after both.
    static void nd_order(const std::vector<uint8_t> &adj, int ng, int lo, int hi, std::vector<int> &out)
    {
        const int n = hi - lo;
        if (n <= 3) { for (int g = lo; g < hi; g++) out.push_back(g); return; }
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

    // adj[ng x ng]: covisibility of the tile groups (kPosesPerTile consecutive free keyframes each)
    int build_schedule(const std::vector<uint8_t> &adj, int ng)
    {
        const int nt = ng, nA = B.nA;
        std::vector<int> order; order.reserve(nt);
        nd_order(adj, ng, 0, ng, order);
        std::vector<int> pos(ng);
        for (int t = 0; t < nt; t++) pos[order[t]] = t;
        // symbolic factorisation on the permuted tile pattern
        std::vector<uint8_t> pat((size_t)nt * nt, 0);
        for (int a = 0; a < ng; a++)
            for (int b2 = 0; b2 < ng; b2++) if (a != b2 && adj[(size_t)a * ng + b2]) { const int i = std::max(pos[a], pos[b2]), j = std::min(pos[a], pos[b2]); pat[(size_t)i * nt + j] = 1; }
        std::vector<int> rows_start(nt + 1, 0), rows, cols_start(nt + 1, 0), cols, level(nt, 0);
        for (int k = 0; k < nt; k++) {
            const size_t r0 = rows.size();
            for (int i = k + 1; i < nt; i++) if (pat[(size_t)i * nt + k]) rows.push_back(i);
            rows_start[k + 1] = (int)rows.size();
            for (size_t x = r0; x < rows.size(); x++)
                for (size_t y = r0; y < x; y++) pat[(size_t)rows[x] * nt + rows[y]] = 1;
        }
        nlevels = 0;
        for (int i = 0; i < nt; i++) {
            int lv = 0;
            for (int k = 0; k < i; k++) if (pat[(size_t)i * nt + k]) { cols.push_back(k); lv = std::max(lv, level[k] + 1); }
            cols_start[i + 1] = (int)cols.size();
            level[i] = lv;
            nlevels = std::max(nlevels, lv + 1);
        }
        l_tiles = nt + (long long)rows.size();
        // task lists by level
        std::vector<std::vector<int>> by_level(nlevels);
        for (int k = 0; k < nt; k++) by_level[level[k]].push_back(k);
        std::vector<int4> panel, update;
        panel_lv.assign(nlevels + 1, 0); update_lv.assign(nlevels + 1, 0);
        std::vector<int> hits((size_t)nt * nt, 0);
        for (int l = 0; l < nlevels; l++) {
            const size_t u0 = update.size();
            for (int k : by_level[l]) {
                panel.push_back(make_int4(k, k, 0, 0));
                for (int x = rows_start[k]; x < rows_start[k + 1]; x++) panel.push_back(make_int4(k, rows[x], 0, 0));
                for (int x = rows_start[k]; x < rows_start[k + 1]; x++)
                    for (int y = rows_start[k]; y <= x; y++) { update.push_back(make_int4(k, rows[x], rows[y], 0)); hits[(size_t)rows[x] * nt + rows[y]]++; }
            }
            for (size_t t = u0; t < update.size(); t++) {
                int &hcount = hits[(size_t)update[t].y * nt + update[t].z];
                if (hcount > 1) update[t].w = 1;                  // several columns of this level update the tile: atomics
            }
            for (size_t t = u0; t < update.size(); t++) hits[(size_t)update[t].y * nt + update[t].z] = 0;
            panel_lv[l + 1] = (int)panel.size(); update_lv[l + 1] = (int)update.size();
        }
        n_panel = panel.size();
        if (panel.size() + update.size() > ((size_t)1 << 26)) { set_last_error("reduced pose system too large / too dense for the tiled Cholesky (more than 2^26 tile tasks)"); return ORBS_E_INVALID; }
        if (int rc = h->ba_tasks.reserve((panel.size() + update.size() + 1) * sizeof(int4))) return rc;
        d_tasks = h->ba_tasks.as<int4>();
        // row map
        std::vector<int> rowbase(std::max(nA, 1));
        std::vector<uint8_t> rowpad((size_t)nt * NB, 1);
        for (int ip = 0; ip < nA; ip++) {
            rowbase[ip] = NB * pos[ip / kPosesPerTile] + 6 * (ip % kPosesPerTile);
            for (int a2 = 0; a2 < 6; a2++) rowpad[rowbase[ip] + a2] = 0;
        }
        int *d_rows_start = d_plan_i, *d_cols_start = d_plan_i + (nt_max + 1), *d_rows = d_plan_i + 2 * (nt_max + 1);
        int *d_cols = d_rows + (size_t)nt_max * (nt_max - 1) / 2 + 1;
        ORBS_CUDA(cudaMemcpyAsync(d_rows_start, rows_start.data(), (nt + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_cols_start, cols_start.data(), (nt + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        if (!rows.empty()) {
            ORBS_CUDA(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            ORBS_CUDA(cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        }
        ORBS_CUDA(cudaMemcpyAsync(d_tasks, panel.data(), panel.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        if (!update.empty()) ORBS_CUDA(cudaMemcpyAsync(d_tasks + panel.size(), update.data(), update.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        if (nA > 0) ORBS_CUDA(cudaMemcpyAsync(d_rowbase, rowbase.data(), nA * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_rowpad, rowpad.data(), rowpad.size(), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaStreamSynchronize(st));                  // the host vectors die here
        plan.rows_start = d_rows_start; plan.rows = d_rows; plan.cols_start = d_cols_start; plan.cols = d_cols;
        plan.panel = d_tasks; plan.update = d_tasks + panel.size();
        B.rowbase = d_rowbase; B.row_pad = d_rowpad;
        return ORBS_OK;
    }

    bool multi() const { return h->nranks > 1; }
    int allreduce(void *buf, size_t n, ncclDataType_t t, ncclRedOp_t op)
    {
        if (!multi()) return ORBS_OK;
        ORBS_NCCL(g_nccl.AllReduce(buf, buf, n, t, op, h->comm, st));
        return ORBS_OK;
    }
    // the stop flag must lead to the same decision on every rank: reduce it (max) when sharded
    int *d_stop = nullptr; int *h_stop = nullptr;
    bool terminate()
    {
        int v = (stop && *stop) ? 1 : 0;
        if (multi()) {
            *h_stop = v;
            if (cudaMemcpyAsync(d_stop, h_stop, sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) return true;
            if (allreduce(d_stop, 1, ncclInt32, ncclMax)) return true;
            if (cudaMemcpyAsync(h_stop, d_stop, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return true;
            if (cudaStreamSynchronize(st) != cudaSuccess) return true;
            v = *h_stop;
        }
        return v != 0;
    }
    KernelTimer &T() { return h->timer; }
    void count(int n = 1) { h->launches += n; }

    int read_scalars()
    {
        // sharded: chi2 and the two computeScale parts are sums over ranks, the diagonal maximum a max
        if (int rc = allreduce(B.scalars, 3, ncclDouble, ncclSum)) return rc;
        if (int rc = allreduce(B.scalars + 3, 1, ncclDouble, ncclMax)) return rc;
        ORBS_CUDA(cudaMemcpyAsync(h_scal, B.scalars, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaMemcpyAsync(h_flag, B.flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        return ORBS_OK;
    }

    // computeActiveErrors + activeRobustChi2 -> scalars[0]
    void errors()
    {
        T().begin(BK_ERRORS, st);
        k_ba_errors<<<err_blocks, 256, 0, st>>>(B);
        k_reduce_partials<<<1, 256, 0, st>>>(B.partial, err_blocks, B.scalars, 0, 0);
        T().end(st);
        count(2);
    }

    int build_system()
    {
        T().begin(BK_BUILD_POINTS, st);
        k_ba_build_points<<<pt_blocks, 256, 0, st>>>(B);
        T().end(st);
        T().begin(BK_BUILD_POSES, st);
        if (pose_blocks) k_ba_build_poses<<<pose_blocks, 256, 0, st>>>(B);
        T().end(st);
        count(2);
        // sharded: every rank saw only its shard's observations of a keyframe -> sum the pose blocks (42 doubles per keyframe);
        // afterwards Hpp / b_p are complete on every rank and enter the reduced system through the lead rank only
        if (multi() && B.nA > 0) {
            if (int rc = allreduce(B.Hpp, (size_t)B.nA * 36, ncclDouble, ncclSum)) return rc;
            if (int rc = allreduce(B.bp, (size_t)B.nA * 6, ncclDouble, ncclSum)) return rc;
        }
        return ORBS_OK;
    }

    // setLambda + Schur + factor + solve + back-substitution; scalars[1] (+ scalars[2]) = computeScale
    int solve()
    {
        const int ld = B.ld;
        ORBS_CUDA(cudaMemsetAsync(B.flags, 0, 4 * sizeof(int), st));
        if (B.n > 0) {
            T().begin(BK_MEMSET, st);
            ORBS_CUDA(cudaMemsetAsync(B.S, 0, (size_t)ld * ld * sizeof(double), st));
            T().end(st);
            const int t = std::max(B.nA * 36, ld);
            T().begin(BK_SCHUR, st);
            k_ba_schur_init<<<(t + 255) / 256, 256, 0, st>>>(B, lambda, h->rank == 0 ? 1 : 0);
            k_ba_schur<<<pt_blocks, 256, 0, st>>>(B, lambda);
            T().end(st);
            // the one exchange step of the sharded solve: sum the partial reduced systems over the map-point shards
            if (multi()) {
                // all-reduce the structurally nonzero tiles only (148 of 1275 at 500 keyframes: 4.8 MB instead of the 82 MB ld x ld array)
                if (int rc = h->ba_packed.reserve(n_panel * 4096 * sizeof(double))) return rc;
                double *packed = h->ba_packed.as<double>();
                k_tiles_pack<<<(unsigned)n_panel, 256, 0, st>>>(B.S, ld, plan.panel, packed, 1);
                if (int rc = allreduce(packed, n_panel * 4096, ncclDouble, ncclSum)) return rc;
                k_tiles_pack<<<(unsigned)n_panel, 256, 0, st>>>(B.S, ld, plan.panel, packed, 0);
                if (int rc = allreduce(B.bs, (size_t)ld, ncclDouble, ncclSum)) return rc;
                count(2);
            }
            count(2);
            T().begin(BK_POTRF, st);
            for (int l = 0; l < nlevels; l++) {
                k_chol_panel<<<panel_lv[l + 1] - panel_lv[l], 256, kPanelSmem, st>>>(B.S, ld, plan.panel + panel_lv[l], Linv, B.flags);
                count(1);
                if (update_lv[l + 1] > update_lv[l]) {
                    k_chol_update<<<update_lv[l + 1] - update_lv[l], 256, kUpdateSmem, st>>>(B.S, ld, plan.update + update_lv[l]);
                    count(1);
                }
            }
            T().end(st);
            T().begin(BK_TRS, st);
            ORBS_CUDA(cudaMemsetAsync(ready, 0, 2 * (size_t)ntiles * sizeof(int), st));
            const int solve_ctas = std::min(ntiles, 128);          // all co-resident (one 256-thread CTA per SM at most)
            k_chol_solve<<<solve_ctas, 256, 0, st>>>(B.S, ld, ntiles, plan, Linv, B.bs, ready, 0);
            k_chol_solve<<<solve_ctas, 256, 0, st>>>(B.S, ld, ntiles, plan, Linv, B.bs, ready + ntiles, 1);
            count(2);
            T().end(st);
            k_ba_take_xp<<<xp_blocks, 256, 0, st>>>(B, lambda, h->rank == 0 ? 1 : 0);
            k_reduce_partials<<<1, 256, 0, st>>>(B.partial, xp_blocks, B.scalars, 2, 0);
            count(2);
        } else {
            ORBS_CUDA(cudaMemsetAsync(B.scalars + 2, 0, sizeof(double), st));
        }
        T().begin(BK_BACKSUB, st);
        k_ba_backsub<<<pt_blocks, 256, 0, st>>>(B, lambda);
        k_reduce_partials<<<1, 256, 0, st>>>(B.partial, pt_blocks, B.scalars, 1, 0);
        T().end(st);
        count(2);
        return ORBS_OK;
    }

    enum { LM_OK = 0, LM_TERMINATE = 1, LM_ERROR = 2 };

    // OptimizationAlgorithmLevenberg::solve, optimization_algorithm_levenberg.cpp:61-164
    // errors_current: the stored edge errors and chi2 belong to the current estimate (the last trial was accepted), so the
    // computeActiveErrors at the top of the next iteration (levenberg.cpp:69-72) would reproduce them bit for bit: skipped,
    // together with its host round trip.  After a rejected trial (estimate restored) they are recomputed, as in g2o.
    bool errors_current = false;
    double chi_current = 0;

    int lm_iteration(int iteration)
    {
        if (iteration == 0) errors_current = false;
        if (!errors_current) errors();
        if (build_system()) return LM_ERROR;
        if (iteration == 0) {
            k_ba_max_diag<<<diag_blocks, 256, 0, st>>>(B);
            k_reduce_partials<<<1, 256, 0, st>>>(B.partial, diag_blocks, B.scalars, 3, 1);
            count(2);
        }
        if (!errors_current || iteration == 0) { if (read_scalars()) return LM_ERROR; chi_current = h_scal[0]; }
        double currentChi = chi_current;
        const double iniChi = currentChi;
        if (iteration == 0) { lambda = 1e-5 * h_scal[3]; ni = 2; nbad = 0; }
        double rho = 0;
        int qmax = 0;
        const int upd_blocks = (B.K + B.P + 255) / 256;
        do {
            if (solve()) return LM_ERROR;
            T().begin(BK_UPDATE, st);
            k_ba_update<<<upd_blocks, 256, 0, st>>>(B);
            T().end(st);
            count(1);
            errors();
            if (read_scalars()) return LM_ERROR;
            const bool ok2 = *h_flag == 0;
            if (!ok2) chol_failures++;
            double tempChi = h_scal[0];
            if (!ok2) tempChi = 1.7976931348623157e308;
            rho = currentChi - tempChi;
            double scale = h_scal[1] + h_scal[2];
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                errors_current = true; chi_current = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                k_ba_restore<<<upd_blocks, 256, 0, st>>>(B);
                count(1);
                errors_current = false;
            }
            qmax++;
            lm_trials++;
        } while (rho < 0 && qmax < 10 && !terminate());
        lm_iterations++;
        if (qmax == 10 || rho == 0) return LM_TERMINATE;
        if ((iniChi - currentChi) * 1e3 < iniChi) nbad++; else nbad = 0;
        if (nbad >= 3) return LM_TERMINATE;
        return LM_OK;
    }

    // SparseOptimizer::optimize, sparse_optimizer.cpp:354-419
    int optimize(int iterations, bool any_active)
    {
        if (!any_active) return ORBS_OK;
        for (int i = 0; i < iterations && !terminate(); i++) {
            const int r = lm_iteration(i);
            if (r == LM_ERROR) return ORBS_E_CUDA;
            if (r != LM_OK) break;
        }
        return ORBS_OK;
    }
};

}  // namespace

extern "C" int orbo_bundle_adjust(orbo_handle *h, int K, float *poses, const uint8_t *fixed, const double *intr, int P, float *points,
                                  int E, const int32_t *e_kf, const int32_t *e_pt, const float *e_uv, const float *e_inv_sigma2,
                                  int two_stage, int its0, int its1, int robust, const volatile int *stop_flag,
                                  double *e_chi2, uint8_t *e_depth_ok, uint8_t *e_outlier, int32_t *stats)
{
    ORBS_REQUIRE(h && poses && fixed && intr && points && e_kf && e_pt && e_uv && e_inv_sigma2, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(K > 0 && P > 0 && E > 0, ORBS_E_INVALID, "empty graph");
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (h->nranks == 1 && stop_flag && *stop_flag) { if (stats) stats[3] = 1; return 1; }   // Optimizer.cc:678-680 (sharded: decided collectively below)
    for (int e = 0; e < E; e++)
        ORBS_REQUIRE(e_kf[e] >= 0 && e_kf[e] < K && e_pt[e] >= 0 && e_pt[e] < P, ORBS_E_INVALID, "edge references a vertex out of range");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const auto t_begin = std::chrono::steady_clock::now();

    // ---- host-side graph layout: edges grouped by point (stable), second CSR by pose
    std::vector<int> pt_start(P + 1, 0), order(E), pose_start(K + 1, 0), pose_edges(E), kf_s(E), pt_s(E);
    for (int e = 0; e < E; e++) pt_start[e_pt[e] + 1]++;
    for (int p = 0; p < P; p++) pt_start[p + 1] += pt_start[p];
    { std::vector<int> fill(pt_start.begin(), pt_start.end() - 1); for (int e = 0; e < E; e++) order[fill[e_pt[e]]++] = e; }
    std::vector<double> obs_s(2 * (size_t)E), w_s(E);
    for (int j = 0; j < E; j++) {
        const int e = order[j];
        kf_s[j] = e_kf[e]; pt_s[j] = e_pt[e];
        obs_s[2 * j] = e_uv[2 * e]; obs_s[2 * j + 1] = e_uv[2 * e + 1]; w_s[j] = e_inv_sigma2[e];
        pose_start[e_kf[e] + 1]++;
    }
    for (int k = 0; k < K; k++) pose_start[k + 1] += pose_start[k];
    { std::vector<int> fill(pose_start.begin(), pose_start.end() - 1); for (int j = 0; j < E; j++) pose_edges[fill[kf_s[j]]++] = j; }
    std::vector<uint8_t> level(E, 0);
    std::vector<double> pts_d(3 * (size_t)P);
    for (size_t i = 0; i < pts_d.size(); i++) pts_d[i] = points[i];

    // ---- device buffers
    Stager S(&h->ba_pool, st, ORBS_MEM_HOST);
    BaHost D;
    D.h = h; D.st = st; D.stop = stop_flag;
    D.h_scal = h->h_scalars.as<double>(); D.h_flag = (int *)(D.h_scal + 16);
    BaDev &B = D.B;
    memset(&B, 0, sizeof B);
    B.K = K; B.P = P; B.E = E;
    const float *d_T = S.in(poses, (size_t)K * 16);
    const uint8_t *d_fixed = S.in(fixed, K);
    B.intr = S.in(intr, (size_t)K * 4);
    B.pt = const_cast<double *>(S.in(pts_d.data(), pts_d.size()));
    B.pt_bak = S.scratch<double>(pts_d.size());
    B.pose = S.scratch<Se3>(K); B.pose_bak = S.scratch<Se3>(K);
    B.pt_start = S.in(pt_start.data(), pt_start.size()); B.e_kf = S.in(kf_s.data(), E); B.e_point = S.in(pt_s.data(), E);
    B.e_obs = S.in(obs_s.data(), obs_s.size()); B.e_w = S.in(w_s.data(), E);
    B.e_level = S.scratch<uint8_t>(E); B.e_err = S.scratch<double>(2 * (size_t)E); B.e_W = S.scratch<double>(18 * (size_t)E);
    B.pose_start = S.in(pose_start.data(), pose_start.size()); B.pose_edges = S.in(pose_edges.data(), E);
    int *d_pose_idx = S.scratch<int>(K); uint8_t *d_pt_active = S.scratch<uint8_t>(P);
    B.pose_idx = d_pose_idx; B.pt_active = d_pt_active;
    B.Hpp = S.scratch<double>(36 * (size_t)K); B.bp = S.scratch<double>(6 * (size_t)K + NB);
    B.Hll = S.scratch<double>(9 * (size_t)P); B.bl = S.scratch<double>(3 * (size_t)P);
    B.x = S.scratch<double>(6 * (size_t)K + 3 * (size_t)P);
    const int ld_max = NB * ((K + kPosesPerTile - 1) / kPosesPerTile);      // 10 keyframes (60 rows + 4 padding rows) per 64-row tile
    B.S = S.scratch<double>((size_t)ld_max * ld_max); B.bs = S.scratch<double>(ld_max);
    D.Linv = S.scratch<double>((size_t)ld_max * NB); D.ready = S.scratch<int>(2 * (size_t)(ld_max / NB) + 2);
    D.nt_max = ld_max / NB;
    D.d_plan_i = S.scratch<int>(2 * (size_t)(D.nt_max + 1) + (size_t)D.nt_max * D.nt_max + 8);
    D.d_rowbase = S.scratch<int>(K + 1); D.d_rowpad = S.scratch<uint8_t>(ld_max + 16);
    int *d_adj = S.scratch<int>((size_t)D.nt_max * D.nt_max + 4);
    ORBS_REQUIRE(ld_max / NB <= 1024, ORBS_E_INVALID, "more than 10240 keyframes in one bundle adjustment");
    const int max_blocks = std::max({(E + 255) / 256, (P + 7) / 8, (K + P + 255) / 256, (6 * K + 255) / 256}) + 1;
    B.partial = S.scratch<double>(max_blocks); B.scalars = S.scratch<double>(8); B.flags = S.scratch<int>(4);
    double *d_chi2 = S.scratch<double>(E); uint8_t *d_depth = S.scratch<uint8_t>(E);
    D.d_stop = S.scratch<int>(4); D.h_stop = D.h_flag + 1;
    int *d_pose_act = S.scratch<int>(K + 1);
    float *d_Tout = S.scratch<float>((size_t)K * 16);
    if (S.rc) return S.rc;
    B.delta = (double)(float)sqrt(5.991); B.dsqr = B.delta * B.delta;              // thHuberMono, Optimizer.cc:591
    ORBS_CUDA(cudaMemsetAsync(B.e_level, 0, E, st));
    ORBS_CUDA(cudaMemsetAsync(B.e_err, 0, 2 * (size_t)E * sizeof(double), st));
    ORBS_CUDA(cudaMemsetAsync(B.scalars, 0, 8 * sizeof(double), st));
    k_ba_import_poses<<<(K + 255) / 256, 256, 0, st>>>(K, d_T, B.pose);
    ORBS_CUDA(cudaMemcpyAsync(d_Tout, d_T, (size_t)K * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    h->launches++;
    D.err_blocks = (E + 255) / 256; D.pt_blocks = (P + 7) / 8; D.pose_blocks = (K + 7) / 8;

    // initializeOptimization(level 0) + buildIndexMapping, sparse_optimizer.cpp:166-267
    std::vector<int> pose_idx(K);
    std::vector<uint8_t> pt_active(P);
    auto init_active = [&]() -> int {
        std::vector<int> pose_act(K + 1, 0);                   // [K] = any active edge at all
        std::fill(pt_active.begin(), pt_active.end(), 0);
        for (int j = 0; j < E; j++) if (!level[j]) { pose_act[kf_s[j]] = 1; pt_active[pt_s[j]] = 1; pose_act[K] = 1; }
        if (D.multi()) {                                       // a keyframe is active if any rank's shard observes it
            ORBS_CUDA(cudaMemcpyAsync(d_pose_act, pose_act.data(), (K + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
            if (int rc = D.allreduce(d_pose_act, K + 1, ncclInt32, ncclMax)) return rc;
            ORBS_CUDA(cudaMemcpyAsync(pose_act.data(), d_pose_act, (K + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
            ORBS_CUDA(cudaStreamSynchronize(st));
        }
        const bool any = pose_act[K] != 0;
        int nA = 0;
        for (int k = 0; k < K; k++) pose_idx[k] = (pose_act[k] && !fixed[k]) ? nA++ : -1;
        B.nA = nA; B.n = 6 * nA; D.ntiles = (nA + kPosesPerTile - 1) / kPosesPerTile; B.ld = NB * D.ntiles;
        D.diag_blocks = (nA + P + 255) / 256; D.xp_blocks = std::max(1, (B.n + 255) / 256);
        ORBS_CUDA(cudaMemcpyAsync(d_pose_idx, pose_idx.data(), K * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_pt_active, pt_active.data(), P, cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        // tile-level structure of the reduced system: free poses i, j are coupled iff an active point is seen by both (edges are
        // grouped by point); tiles = groups of kPosesPerTile consecutive hessian indices.  Sharded: union over the ranks' shards.
        const int ng = D.ntiles;
        std::vector<uint8_t> adj((size_t)ng * ng, 0);
        {
            int gs[64];
            for (int p = 0; p < P; p++) {
                int m = 0;
                for (int j = pt_start[p]; j < pt_start[p + 1]; j++) {
                    if (level[j] || pose_idx[kf_s[j]] < 0) continue;
                    const int g = pose_idx[kf_s[j]] / kPosesPerTile;
                    bool seen = false;
                    for (int q = 0; q < m; q++) if (gs[q] == g) { seen = true; break; }
                    if (!seen) {
                        for (int q = 0; q < m; q++) { adj[(size_t)g * ng + gs[q]] = 1; adj[(size_t)gs[q] * ng + g] = 1; }
                        if (m < 64) gs[m++] = g;
                        else for (int q2 = 0; q2 < ng; q2++) { adj[(size_t)g * ng + q2] = 1; adj[(size_t)q2 * ng + g] = 1; }   // > 64 distinct groups: couple to all
                    }
                }
            }
        }
        if (D.multi() && ng > 0) {
            std::vector<int> a32(adj.begin(), adj.end());
            ORBS_CUDA(cudaMemcpyAsync(d_adj, a32.data(), a32.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            if (int rc = D.allreduce(d_adj, a32.size(), ncclInt32, ncclMax)) return rc;
            ORBS_CUDA(cudaMemcpyAsync(a32.data(), d_adj, a32.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
            ORBS_CUDA(cudaStreamSynchronize(st));
            for (size_t q = 0; q < adj.size(); q++) adj[q] = (uint8_t)a32[q];
        }
        if (ng > 0 && D.build_schedule(adj, ng)) return -1;
        return any ? 1 : 0;
    };

    if (D.multi() && D.terminate()) { if (stats) stats[3] = 1; return 1; }
    int any = init_active();
    if (any < 0) return any;
    const auto t_loop = std::chrono::steady_clock::now();
    B.robust = two_stage ? 1 : (robust ? 1 : 0);
    int rc = D.optimize(its0, any == 1);
    if (rc) return rc;
    std::vector<double> chi2_s(E);
    std::vector<uint8_t> depth_s(E);
    auto edge_check = [&]() -> int {
        k_ba_edge_check<<<(E + 255) / 256, 256, 0, st>>>(B, d_chi2, d_depth);
        h->launches++;
        ORBS_CUDA(cudaMemcpyAsync(chi2_s.data(), d_chi2, E * sizeof(double), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaMemcpyAsync(depth_s.data(), d_depth, E, cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        return ORBS_OK;
    };
    if (two_stage && !D.terminate()) {
        if ((rc = edge_check())) return rc;
        for (int j = 0; j < E; j++) if (chi2_s[j] > 5.991 || !depth_s[j]) level[j] = 1;      // Optimizer.cc:691-705
        ORBS_CUDA(cudaMemcpyAsync(B.e_level, level.data(), E, cudaMemcpyHostToDevice, st));
        B.robust = 0;
        any = init_active();
        if (any < 0) return any;
        if ((rc = D.optimize(its1, any == 1))) return rc;
    }
    if ((rc = edge_check())) return rc;
    const auto t_loop_end = std::chrono::steady_clock::now();
    for (int j = 0; j < E; j++) {                                                         // Optimizer.cc:734-766
        const int e = order[j];
        if (e_chi2) e_chi2[e] = chi2_s[j];
        if (e_depth_ok) e_depth_ok[e] = depth_s[j];
        if (e_outlier) e_outlier[e] = (uint8_t)(chi2_s[j] > 5.991 || !depth_s[j]);
    }
    k_ba_export_poses<<<(K + 255) / 256, 256, 0, st>>>(K, B.pose, d_fixed, d_Tout);
    h->launches++;
    ORBS_CUDA(cudaMemcpyAsync(poses, d_Tout, (size_t)K * 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaMemcpyAsync(pts_d.data(), B.pt, pts_d.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < pts_d.size(); i++) points[i] = (float)pts_d[i];
    {
        const auto t_end = std::chrono::steady_clock::now();
        auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        h->ba_timing[0] = sec(t_loop, t_loop_end); h->ba_timing[1] = sec(t_begin, t_end); h->ba_timing[2] = sec(t_begin, t_loop);
        h->ba_timing[3] = (double)B.ld;
        h->ba_skyline[0] = D.ntiles; h->ba_skyline[1] = D.l_tiles; h->ba_skyline[2] = D.nlevels;
    }
    if (stats) { stats[0] = D.lm_iterations; stats[1] = D.lm_trials; stats[2] = D.chol_failures; stats[3] = 0; }
    return ORBS_OK;
}

extern "C" int orbo_last_ba_timing(orbo_handle *h, double *out4)
{
    ORBS_REQUIRE(h && out4, ORBS_E_INVALID, "null argument");
    for (int i = 0; i < 4; i++) out4[i] = h->ba_timing[i];
    return ORBS_OK;
}

extern "C" int orbo_last_ba_structure(orbo_handle *h, long long *out2)
{
    ORBS_REQUIRE(h && out2, ORBS_E_INVALID, "null argument");
    out2[0] = h->ba_skyline[0]; out2[1] = h->ba_skyline[1]; out2[2] = h->ba_skyline[2];
    return ORBS_OK;
}

extern "C" int orbo_set_profiling(orbo_handle *h, int enabled)
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

extern "C" int orbo_get_kernel_times(orbo_handle *h, double *total_ms, long long *counts, int n)
{
    ORBS_REQUIRE(h && total_ms && counts && n > 0, ORBS_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    ORBS_CUDA(cudaStreamSynchronize(h->stream));
    h->timer.collect();
    for (int i = 0; i < n && i < KernelTimer::kMaxKernels; i++) { total_ms[i] = h->timer.total_ms[i]; counts[i] = h->timer.count[i]; }
    return ORBS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Essential-graph (Sim3 pose graph) optimisation: host-driven Levenberg over the kernels of pose_graph.cuh, normal equations factored by
// the tiled sparse Cholesky of the BA (same kernels, own schedule: 9 vertices per 64-row tile)
namespace {

struct PgHost {
    orbo_handle *h; cudaStream_t st; PgDev P;
    DevBuf bufs[16]; int nbuf = 0;
    int nt = 0, nlevels = 0; size_t n_panel = 0;
    std::vector<int> panel_lv, update_lv;
    CholPlan plan = {};
    double *Linv = nullptr, *packed = nullptr; int *ready = nullptr, *flags = nullptr;
    int eblocks = 0;
    ~PgHost() { for (int i = 0; i < nbuf; i++) bufs[i].release(); }
    template <typename T> T *alloc(size_t n, int *rc) { if (*rc) return nullptr; if (nbuf >= 16) { *rc = ORBS_E_INVALID; return nullptr; } *rc = bufs[nbuf].reserve(std::max<size_t>(n, 1) * sizeof(T)); return bufs[nbuf++].as<T>(); }

    // tile groups of kPgPerTile consecutive free vertices; adjacency from the edges; nested dissection; symbolic factorisation; task lists
    int schedule(int K, const std::vector<int> &hidx, int E, const int32_t *e_i, const int32_t *e_j)
    {
        const int nA = P.nA;
        nt = (nA + kPgPerTile - 1) / kPgPerTile;
        const int ng = nt;
        std::vector<uint8_t> adj((size_t)ng * ng, 0);
        for (int e = 0; e < E; e++) {
            const int a = hidx[e_i[e]], b = hidx[e_j[e]];
            if (a < 0 || b < 0) continue;
            adj[(size_t)(a / kPgPerTile) * ng + b / kPgPerTile] = 1; adj[(size_t)(b / kPgPerTile) * ng + a / kPgPerTile] = 1;
        }
        std::vector<int> order; order.reserve(nt);
        BaHost::nd_order(adj, ng, 0, ng, order);
        std::vector<int> pos(ng);
        for (int t = 0; t < nt; t++) pos[order[t]] = t;
        std::vector<uint8_t> pat((size_t)nt * nt, 0);
        for (int a = 0; a < ng; a++)
            for (int b = 0; b < ng; b++) if (a != b && adj[(size_t)a * ng + b]) { const int i = std::max(pos[a], pos[b]), j = std::min(pos[a], pos[b]); pat[(size_t)i * nt + j] = 1; }
        std::vector<int> rows_start(nt + 1, 0), rows, cols_start(nt + 1, 0), cols, level(nt, 0);
        for (int k = 0; k < nt; k++) {
            const size_t r0 = rows.size();
            for (int i = k + 1; i < nt; i++) if (pat[(size_t)i * nt + k]) rows.push_back(i);
            rows_start[k + 1] = (int)rows.size();
            for (size_t x = r0; x < rows.size(); x++) for (size_t y = r0; y < x; y++) pat[(size_t)rows[x] * nt + rows[y]] = 1;
        }
        nlevels = 0;
        for (int i = 0; i < nt; i++) {
            int lv = 0;
            for (int k = 0; k < i; k++) if (pat[(size_t)i * nt + k]) { cols.push_back(k); lv = std::max(lv, level[k] + 1); }
            cols_start[i + 1] = (int)cols.size();
            level[i] = lv; nlevels = std::max(nlevels, lv + 1);
        }
        std::vector<std::vector<int>> by_level(nlevels);
        for (int k = 0; k < nt; k++) by_level[level[k]].push_back(k);
        std::vector<int4> panel, update;
        panel_lv.assign(nlevels + 1, 0); update_lv.assign(nlevels + 1, 0);
        std::vector<int> hits((size_t)nt * nt, 0);
        for (int l = 0; l < nlevels; l++) {
            const size_t u0 = update.size();
            for (int k : by_level[l]) {
                panel.push_back(make_int4(k, k, 0, 0));
                for (int x = rows_start[k]; x < rows_start[k + 1]; x++) panel.push_back(make_int4(k, rows[x], 0, 0));
                for (int x = rows_start[k]; x < rows_start[k + 1]; x++)
                    for (int y = rows_start[k]; y <= x; y++) { update.push_back(make_int4(k, rows[x], rows[y], 0)); hits[(size_t)rows[x] * nt + rows[y]]++; }
            }
            for (size_t t = u0; t < update.size(); t++) if (hits[(size_t)update[t].y * nt + update[t].z] > 1) update[t].w = 1;
            for (size_t t = u0; t < update.size(); t++) hits[(size_t)update[t].y * nt + update[t].z] = 0;
            panel_lv[l + 1] = (int)panel.size(); update_lv[l + 1] = (int)update.size();
        }
        n_panel = panel.size();
        ORBS_REQUIRE(panel.size() + update.size() <= ((size_t)1 << 26), ORBS_E_INVALID, "pose graph too large / too dense for the tiled Cholesky");
        std::vector<int> rowbase(std::max(nA, 1));
        std::vector<uint8_t> rowpad((size_t)nt * NB, 1);
        for (int ip = 0; ip < nA; ip++) {
            rowbase[ip] = NB * pos[ip / kPgPerTile] + 7 * (ip % kPgPerTile);
            for (int a = 0; a < 7; a++) rowpad[rowbase[ip] + a] = 0;
        }
        int rc = ORBS_OK;
        int *d_rs = alloc<int>(nt + 1, &rc), *d_cs = alloc<int>(nt + 1, &rc), *d_rows = alloc<int>(rows.size(), &rc), *d_cols = alloc<int>(cols.size(), &rc);
        int4 *d_tasks = alloc<int4>(panel.size() + update.size(), &rc);
        int *d_rowbase = alloc<int>(nA, &rc); uint8_t *d_rowpad = alloc<uint8_t>(rowpad.size(), &rc);
        if (rc) return rc;
        ORBS_CUDA(cudaMemcpyAsync(d_rs, rows_start.data(), (nt + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_cs, cols_start.data(), (nt + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        if (!rows.empty()) { ORBS_CUDA(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, st)); ORBS_CUDA(cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice, st)); }
        ORBS_CUDA(cudaMemcpyAsync(d_tasks, panel.data(), panel.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        if (!update.empty()) ORBS_CUDA(cudaMemcpyAsync(d_tasks + panel.size(), update.data(), update.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_rowbase, rowbase.data(), std::max(nA, 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_rowpad, rowpad.data(), rowpad.size(), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        plan.rows_start = d_rs; plan.rows = d_rows; plan.cols_start = d_cs; plan.cols = d_cols; plan.panel = d_tasks; plan.update = d_tasks + panel.size();
        P.rowbase = d_rowbase; P.row_pad = d_rowpad; P.ld = nt * NB;
        return ORBS_OK;
    }

    int chi2(double *out)
    {
        k_pg_errors<<<eblocks, 128, 0, st>>>(P);
        k_pg_sum<<<1, 32, 0, st>>>(P.partial, eblocks, P.scalars);
        h->launches += 2;
        ORBS_CUDA(cudaMemcpyAsync(out, P.scalars, sizeof(double), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        return ORBS_OK;
    }
    // S <- H (packed copy of the structurally nonzero tiles) + lambda I, factor, solve; ok = the factorisation succeeded
    int solve(double lambda, bool *ok)
    {
        const int ld = P.ld;
        ORBS_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
        k_tiles_pack<<<(unsigned)n_panel, 256, 0, st>>>(P.S, ld, plan.panel, packed, 0);
        k_pg_prepare<<<(ld + 255) / 256, 256, 0, st>>>(P, lambda);
        h->launches += 2;
        for (int l = 0; l < nlevels; l++) {
            k_chol_panel<<<panel_lv[l + 1] - panel_lv[l], 256, kPanelSmem, st>>>(P.S, ld, plan.panel + panel_lv[l], Linv, flags);
            h->launches++;
            if (update_lv[l + 1] > update_lv[l]) { k_chol_update<<<update_lv[l + 1] - update_lv[l], 256, kUpdateSmem, st>>>(P.S, ld, plan.update + update_lv[l]); h->launches++; }
        }
        ORBS_CUDA(cudaMemsetAsync(ready, 0, 2 * (size_t)nt * sizeof(int), st));
        const int ctas = std::min(nt, 128);
        k_chol_solve<<<ctas, 256, 0, st>>>(P.S, ld, nt, plan, Linv, P.v, ready, 0);
        k_chol_solve<<<ctas, 256, 0, st>>>(P.S, ld, nt, plan, Linv, P.v, ready + nt, 1);
        h->launches += 2;
        int f = 0;
        ORBS_CUDA(cudaMemcpyAsync(&f, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        *ok = f == 0;
        return ORBS_OK;
    }
};

}  // namespace

extern "C" int orbo_optimize_pose_graph(orbo_handle *h, int K, double *sim3, const uint8_t *fixed, int E, const int32_t *e_i, const int32_t *e_j,
                                        const double *e_meas, int fix_scale, int iterations, double lambda_init, int32_t *stats)
{
    ORBS_REQUIRE(h && sim3 && fixed && e_i && e_j && e_meas, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(K > 0 && E > 0 && iterations >= 0, ORBS_E_INVALID, "non-positive size");
    for (int e = 0; e < E; e++) ORBS_REQUIRE(e_i[e] >= 0 && e_i[e] < K && e_j[e] >= 0 && e_j[e] < K && e_i[e] != e_j[e], ORBS_E_INVALID, "edge endpoint out of range");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    if (stats) { stats[0] = stats[1] = stats[2] = 0; }
    PgHost G; G.h = h; G.st = h->stream;
    PgDev &P = G.P;
    std::vector<int> hidx(K);
    int nA = 0;
    for (int k = 0; k < K; k++) hidx[k] = fixed[k] ? -1 : nA++;
    if (nA == 0) return ORBS_OK;
    P.K = K; P.E = E; P.nA = nA; P.fix_scale = fix_scale ? 1 : 0;
    static_assert(sizeof(Sim3) == 8 * sizeof(double), "Sim3 must be 8 packed doubles (r xyzw, t, s)");
    int rc = ORBS_OK;
    if ((rc = G.schedule(K, hidx, E, e_i, e_j))) return rc;
    const int ld = P.ld, nt = G.nt;
    ORBS_REQUIRE((size_t)ld * ld * sizeof(double) <= ((size_t)24 << 30), ORBS_E_INVALID, "pose graph too large for one GPU");
    G.eblocks = (E + 127) / 128;
    P.verts = G.alloc<Sim3>(K, &rc); P.backup = G.alloc<Sim3>(K, &rc);
    int *d_hidx = G.alloc<int>(K, &rc), *d_ei = G.alloc<int>(E, &rc), *d_ej = G.alloc<int>(E, &rc);
    Sim3 *d_meas = G.alloc<Sim3>(E, &rc);
    P.err = G.alloc<double>((size_t)7 * E, &rc);
    // one buffer: S | b | v | partial | scalars | Linv | packed | ready | flags
    const size_t nS = (size_t)ld * ld, nLinv = (size_t)nt * NB * NB, nPacked = G.n_panel * 4096;
    double *big = G.alloc<double>(nS + 2 * (size_t)ld + G.eblocks + 8 + nLinv + nPacked + (size_t)nt + 8, &rc);
    if (rc) return rc;
    P.S = big; P.b = big + nS; P.v = P.b + ld; P.partial = P.v + ld; P.scalars = P.partial + G.eblocks; G.Linv = P.scalars + 8; G.packed = G.Linv + nLinv;
    G.ready = reinterpret_cast<int *>(G.packed + nPacked); G.flags = G.ready + 2 * nt;
    P.hidx = d_hidx; P.e_i = d_ei; P.e_j = d_ej; P.meas = d_meas;
    cudaStream_t st = h->stream;
    ORBS_CUDA(cudaMemcpyAsync(P.verts, sim3, sizeof(Sim3) * K, cudaMemcpyHostToDevice, st));
    ORBS_CUDA(cudaMemcpyAsync(d_hidx, hidx.data(), sizeof(int) * K, cudaMemcpyHostToDevice, st));
    ORBS_CUDA(cudaMemcpyAsync(d_ei, e_i, sizeof(int) * E, cudaMemcpyHostToDevice, st));
    ORBS_CUDA(cudaMemcpyAsync(d_ej, e_j, sizeof(int) * E, cudaMemcpyHostToDevice, st));
    ORBS_CUDA(cudaMemcpyAsync(d_meas, e_meas, sizeof(Sim3) * E, cudaMemcpyHostToDevice, st));
    std::vector<double> hb(ld), hx(ld);
    std::vector<uint8_t> rowpad(ld);
    ORBS_CUDA(cudaMemcpyAsync(rowpad.data(), P.row_pad, ld, cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaStreamSynchronize(st));
    double lambda = 0, ni = 2;
    int nbad = 0, its = 0, trials = 0, fails = 0;
    // OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-164) per iteration, SparseOptimizer::optimize around it
    for (int it = 0; it < iterations; it++) {
        double currentChi = 0, tempChi = 0;
        if ((rc = G.chi2(&currentChi))) return rc;
        const double iniChi = currentChi;
        ORBS_CUDA(cudaMemsetAsync(P.S, 0, nS * sizeof(double), st));
        ORBS_CUDA(cudaMemsetAsync(P.b, 0, (size_t)ld * sizeof(double), st));
        k_pg_build<<<(E + 63) / 64, 64, 0, st>>>(P);
        k_tiles_pack<<<(unsigned)G.n_panel, 256, 0, st>>>(P.S, ld, G.plan.panel, G.packed, 1);        // keep H: the factorisation overwrites S
        h->launches += 2;
        ORBS_CUDA(cudaMemcpyAsync(hb.data(), P.b, sizeof(double) * ld, cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        if (it == 0) {
            if (lambda_init > 0) lambda = lambda_init;                                                   // setUserLambdaInit
            else {                                                                                       // computeLambdaInit: 1e-5 * max diagonal
                std::vector<double> diag(ld);
                ORBS_CUDA(cudaMemcpy2DAsync(diag.data(), sizeof(double), P.S, ((size_t)ld + 1) * sizeof(double), sizeof(double), ld, cudaMemcpyDeviceToHost, st));
                ORBS_CUDA(cudaStreamSynchronize(st));
                double mx = 0;
                for (int t = 0; t < ld; t++) if (!rowpad[t]) mx = std::max(mx, std::fabs(diag[t]));
                lambda = 1e-5 * mx;
            }
            ni = 2; nbad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            bool ok2 = true;
            if ((rc = G.solve(lambda, &ok2))) return rc;
            if (!ok2) fails++;
            ORBS_CUDA(cudaMemcpyAsync(hx.data(), P.v, sizeof(double) * ld, cudaMemcpyDeviceToHost, st));
            k_pg_update<<<(K + 127) / 128, 128, 0, st>>>(P);
            h->launches++;
            if ((rc = G.chi2(&tempChi))) return rc;
            if (!ok2) tempChi = 1.7976931348623157e308;
            rho = currentChi - tempChi;
            double scale = 0.;
            for (int t = 0; t < ld; t++) if (!rowpad[t]) scale += hx[t] * (lambda * hx[t] + hb[t]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha); ni = 2; currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                k_pg_restore<<<(K + 127) / 128, 128, 0, st>>>(P);
                h->launches++;
            }
            qmax++; trials++;
        } while (rho < 0 && qmax < 10);
        its++;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nbad++; else nbad = 0;
        if (nbad >= 3) break;
    }
    ORBS_CUDA(cudaMemcpyAsync(sim3, P.verts, sizeof(Sim3) * K, cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaStreamSynchronize(st));
    if (stats) { stats[0] = its; stats[1] = trials; stats[2] = fails; }
    return ORBS_OK;
}
