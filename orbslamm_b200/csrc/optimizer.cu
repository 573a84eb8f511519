// optimizer.cu -- host driver and C-ABI of the optimizer (orbo_*): PoseOptimization and OptimizeSim3 (whole LM schedule on the device, one CTA per
// frame / keyframe pair), Local / Global bundle adjustment (device-resident Levenberg-Marquardt control block, the host only enqueues trial "slots":
// per-target Schur assembly, sparse tiled Cholesky of the reduced pose system in one persistent DMMA kernel, back-substitution; sharded over GPUs with
// one exchange of the reduced system per trial), the Sim3Solver RANSAC pieces and the essential-graph optimisation.
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
#include <cub/cub.cuh>
#include "ba_kernels.cuh"
#include "peer_reduce.cuh"

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
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
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
    DevBuf ba_tasks;         // task / dependency lists of the tiled solver (sized by the symbolic factorisation)
    DevBuf ba_sys;           // reduced system: [A tiles | rhs] (contiguous: the one all-reduce of the sharded solve), L, Linv, y, x
    DevBuf ba_flags;         // dataflow epoch flags + ticket counters of k_rs_solve
    DevBuf ba_cub;           // cub temp storage (pair sort, scans)
    DevBuf ba_items;         // per-work-item partial blocks of the Schur assembly
    DevBuf pg_bufs[24];      // essential-graph buffers, kept across calls (a cudaMalloc / cudaFree pair per buffer and call cost 5 - 200 ms in a busy process)
    PinnedBuf h_scalars;     // LmCtl copies + the mirrored stop flag
    PinnedBuf ba_host;       // pinned arena of a BA call: the point-sorted graph arrays are laid out here and go to the device by DMA at PCIe speed
                             // (from pageable std::vectors the 12 MB of a 500-keyframe graph took ~3 ms of a 13 ms call)
    static constexpr int kCtlCopies = 4;
    cudaEvent_t slot_done[kCtlCopies] = {nullptr, nullptr, nullptr, nullptr};
    int rs_epoch = 0;
    int sm_count = 148;
    KernelTimer timer;       // BA kernels, ids = BaK
    ncclComm_t comm = nullptr;   // set by orbo_comm_init: orbo_bundle_adjust becomes a collective over map-point shards
    int nranks = 1, rank = 0;
    // NVLink peer-memory exchange (peer_reduce.cuh): every rank's control block and packed system buffer mapped through cudaIpc handles
    bool peer_on = false;
    PeerDev peer = {};
    size_t peer_sys_bytes = 0;
    unsigned peer_epoch = 0, peer_sepoch = 0;
    long long ba_skyline[3] = {0, 0, 0};
    double ba_timing[4] = {0, 0, 0, 0};   // last BA call: LM-loop seconds, total seconds, setup (layout + H2D) seconds, Schur bytes
};

enum BaK { BK_ERRORS = 0, BK_BUILD_POINTS, BK_BUILD_POSES, BK_PREP, BK_SCHUR, BK_SOLVE, BK_BACKSUB, BK_UPDATE, BK_DECIDE, BK_EXCHANGE, BK_COUNT };


namespace {
// ---- NVLink peer-memory exchange: set-up (collective over the NCCL communicator) ----------------------------------------
// Exchanges one cudaIpc handle per rank (all-gather of 64 bytes through NCCL) and maps the peers' allocations.  Returns the mapped pointers in out[] (own
// pointer for own rank); false if anything fails on this rank.
bool peer_exchange(orbo_handle *h, void *mine, void **out)
{
    cudaIpcMemHandle_t hm;
    if (cudaIpcGetMemHandle(&hm, mine) != cudaSuccess) { cudaGetLastError(); return false; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    uint8_t *d = nullptr;
    if (cudaMalloc(&d, 64 * (size_t)(h->nranks + 1)) != cudaSuccess) { cudaGetLastError(); return false; }
    bool ok = cudaMemcpyAsync(d, &hm, 64, cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
    ok = ok && g_nccl.AllGather(d, d + 64, 64, ncclChar, h->comm, h->stream) == ncclSuccess;
    std::vector<cudaIpcMemHandle_t> all(h->nranks);
    ok = ok && cudaMemcpyAsync(all.data(), d + 64, 64 * (size_t)h->nranks, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(h->stream) == cudaSuccess;
    cudaFree(d);
    if (!ok) { cudaGetLastError(); return false; }
    for (int r = 0; r < h->nranks; r++) {
        if (r == h->rank) { out[r] = mine; continue; }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; p = nullptr; }
        out[r] = p;
    }
    return ok;
}

// every rank must take the same path: agree on `ok` (1 only if it is 1 everywhere)
bool peer_agree(orbo_handle *h, bool ok)
{
    int *d = nullptr, v = ok ? 1 : 0;
    if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return false;
    cudaMemcpyAsync(d, &v, sizeof(int), cudaMemcpyHostToDevice, h->stream);
    const bool sent = g_nccl.AllReduce(d, d, 1, ncclInt32, ncclMin, h->comm, h->stream) == ncclSuccess;
    cudaMemcpyAsync(&v, d, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(d);
    return sent && v == 1;
}

void peer_close(orbo_handle *h, bool sys_only)
{
    for (int r = 0; r < kMaxPeers; r++) {
        if (r != h->rank && h->peer.sys[r]) cudaIpcCloseMemHandle(h->peer.sys[r]);
        if (!sys_only && r != h->rank && h->peer.ctl[r]) cudaIpcCloseMemHandle(h->peer.ctl[r]);
    }
    if (h->peer.sys[h->rank]) cudaFree(h->peer.sys[h->rank]);
    if (!sys_only && h->peer.ctl[h->rank]) cudaFree(h->peer.ctl[h->rank]);
    for (int r = 0; r < kMaxPeers; r++) { h->peer.sys[r] = nullptr; if (!sys_only) h->peer.ctl[r] = nullptr; }
    h->peer_sys_bytes = 0;
    cudaGetLastError();
}

// control blocks: once per communicator.  ORBS_NO_PEER=1 keeps the NCCL path (A/B comparisons).
void peer_setup(orbo_handle *h)
{
    h->peer_on = false;
    h->peer = PeerDev{};
    h->peer.n = h->nranks; h->peer.rank = h->rank;
    h->peer_epoch = h->peer_sepoch = 0;
    const char *off = getenv("ORBS_NO_PEER");
    bool ok = h->nranks > 1 && h->nranks <= kMaxPeers && g_nccl.AllGather && !(off && off[0] == '1');
    PeerCtlDev *mine = nullptr;
    if (ok) ok = cudaMalloc(&mine, sizeof(PeerCtlDev)) == cudaSuccess && cudaMemset(mine, 0, sizeof(PeerCtlDev)) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    if (h->nranks > 1 && g_nccl.AllGather) {                        // the collective part runs on every rank or on none (nranks and the symbol are the same everywhere)
        void *ptrs[kMaxPeers] = {};
        const bool all_want = peer_agree(h, ok);
        if (all_want) {
            ok = peer_exchange(h, mine, ptrs);
            for (int r = 0; r < h->nranks; r++) h->peer.ctl[r] = (PeerCtlDev *)ptrs[r];
            ok = peer_agree(h, ok);
        } else ok = false;
    } else ok = false;
    if (!ok) { if (mine) { h->peer.ctl[h->rank] = mine; } peer_close(h, false); return; }
    h->peer_on = true;
}

// the packed system buffer in shareable memory: (re)allocated and re-mapped collectively when it has to grow (the size is the same on every rank)
int peer_reserve_sys(orbo_handle *h, size_t bytes)
{
    if (bytes <= h->peer_sys_bytes) return ORBS_OK;
    cudaStreamSynchronize(h->stream);
    peer_close(h, true);
    peer_agree(h, true);                                            // nobody frees while a peer still has the old buffer mapped ... and everybody has unmapped
    const size_t want = bytes + bytes / 4 + 4096;
    void *mine = nullptr;
    bool ok = cudaMalloc(&mine, want) == cudaSuccess;
    void *ptrs[kMaxPeers] = {};
    if (peer_agree(h, ok)) {
        ok = peer_exchange(h, mine, ptrs);
        for (int r = 0; r < h->nranks; r++) h->peer.sys[r] = (double *)ptrs[r];
        ok = peer_agree(h, ok);
    } else ok = false;
    if (!ok) {
        if (mine && !h->peer.sys[h->rank]) h->peer.sys[h->rank] = (double *)mine;
        peer_close(h, true);
        h->peer_on = false;                                         // every rank lands here together: NCCL from now on
        return ORBS_OK;
    }
    h->peer_sys_bytes = want;
    return ORBS_OK;
}
}  // namespace

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
    if (int rc = h->h_scalars.reserve(orbo_handle::kCtlCopies * sizeof(LmCtl) + 256)) { orbo_destroy(h); return rc; }
    for (auto &ev : h->slot_done) { e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); if (e != cudaSuccess) { orbo_destroy(h); return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); } }
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) h->sm_count = v; }
    // opt-in shared memory limits are per-device function attributes: set once, to the static sizes
    cudaFuncSetAttribute(k_rs_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsSmemBytes);
    *out = h;
    return ORBS_OK;
}

int orbo_destroy(orbo_handle *h)
{
    if (!h) return ORBS_OK;
    cudaSetDevice(h->device);
    if (h->stream && h->own_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    else cudaDeviceSynchronize();
    if (h->peer.ctl[h->rank] || h->peer.sys[h->rank]) peer_close(h, false);
    if (h->comm) { if (g_nccl.CommAbort) g_nccl.CommAbort(h->comm); else g_nccl.CommDestroy(h->comm); }   // abort: never block on a peer at teardown
    h->pool.release(); h->ba_pool.release(); h->ba_tasks.release(); h->ba_sys.release(); h->ba_flags.release(); h->ba_cub.release(); h->ba_items.release(); for (auto &b : h->pg_bufs) b.release();
    h->h_scalars.release(); h->ba_host.release(); h->timer.release();
    for (auto &ev : h->slot_done) if (ev) cudaEventDestroy(ev);
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
    if (h->peer_on || h->peer.ctl[h->rank]) { peer_close(h, false); h->peer_on = false; }
    ORBS_NCCL(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->nranks = nranks; h->rank = rank;
    peer_setup(h);                                                  // NVLink peer-memory exchange where the ranks can map each other's memory, else NCCL only
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

int orbo_sim3_compute(orbo_handle *h, int n_hyp, const float *X1, const float *X2, int fix_scale, float *T12, float *T21, float *Rts, int memspace)
{
    ORBS_REQUIRE(h && X1 && X2 && T12 && T21, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_hyp > 0, ORBS_E_INVALID, "at least one hypothesis");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const float *d1 = S.in(X1, (size_t)n_hyp * 9), *d2 = S.in(X2, (size_t)n_hyp * 9);
    float *o12 = S.inout(T12, (size_t)n_hyp * 16, false), *o21 = S.inout(T21, (size_t)n_hyp * 16, false);
    float *ort = Rts ? S.inout(Rts, (size_t)n_hyp * 13, false) : S.scratch<float>((size_t)n_hyp * 13);
    if (S.rc) return S.rc;
    k_sim3_compute<<<(n_hyp + 127) / 128, 128, 0, h->stream>>>(n_hyp, d1, d2, fix_scale ? 1 : 0, o12, o21, ort);
    h->launches++;
    ORBS_CUDA(cudaGetLastError());
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
// Bundle adjustment driver: the graph layout on the host, the structure of the reduced system (tile adjacency -> nested-dissection
// order -> symbolic factorisation -> task list; Schur pairs sorted by target block) once per call, then LM "slots" enqueued
// back to back.  Every slot is the same kernel sequence; what it does is decided by the device control block (LmCtl):
//     [state == BUILD]  linearise (Hll, b_l, Hpl | Hpp, b_p), first time: lambda init
//     [state != DONE]   per-point G, Z | zero + lambda | Schur blocks | (sharded: ONE all-reduce of the packed tiles + rhs) |
//                       factor + solves (one persistent kernel) | back-substitution | push + oplus | errors | reduce |
//                       (sharded: 5 scalars) | decide | pop if rejected
// The host never waits for a decision: it reads a pinned copy of the control block two slots later and stops enqueueing.
namespace {

// profiling aid (ORBS_BA_STAGES=1): wall-clock marks of the host-side stages of one orbo_bundle_adjust call, printed to stderr at the end of the call
struct StageTrace {
    bool on = false;
    std::vector<std::pair<const char *, std::chrono::steady_clock::time_point>> v;
    StageTrace() { on = getenv("ORBS_BA_STAGES") != nullptr; if (on) v.reserve(32); }
    void mark(const char *name) { if (on) v.emplace_back(name, std::chrono::steady_clock::now()); }
    void dump() const
    {
        if (!on || v.empty()) return;
        fprintf(stderr, "[orbo_bundle_adjust stages, ms]");
        for (size_t i = 1; i < v.size(); i++) fprintf(stderr, " %s %.3f", v[i].first, std::chrono::duration<double, std::milli>(v[i].second - v[i - 1].second).count());
        fprintf(stderr, " | total %.3f\n", std::chrono::duration<double, std::milli>(v.back().second - v.front().second).count());
    }
};

struct BaRun {
    orbo_handle *h = nullptr;
    StageTrace *trace = nullptr;
    cudaStream_t st = nullptr;
    Stager *S = nullptr;
    BaDev B;
    TilePlanHost plan;
    RsPlan rsplan = {};
    RsBuf rsbuf = {};
    const volatile uint8_t *stop = nullptr;
    int *h_stop = nullptr;                 // pinned, read by k_lm_reduce through the unified address space
    LmCtl *h_ctl = nullptr;                // pinned [kCtlCopies]
    double *diag6 = nullptr;               // [6 nA | nranks] first-iteration diagonal exchange
    int err_blocks = 0, pt_blocks = 0, diag_blocks = 0, xp_blocks = 0, upd_blocks = 0;
    int solve_ctas = 0, n_seg = 0, n_items = 0;
    std::vector<int> pose_idx;
    const uint8_t *fixed = nullptr;
    // device scratch that build_structure needs
    int *d_pose_idx = nullptr, *d_rowbase = nullptr, *d_slot_of = nullptr, *d_tile_pose = nullptr, *d_pose_act = nullptr;
    uint8_t *d_adj = nullptr;
    long long pair_cap = 0;
    int *d_pair_cnt = nullptr, *d_pair_start = nullptr, *d_head = nullptr, *d_scan = nullptr, *d_seg_start = nullptr, *d_nseg = nullptr;
    unsigned long long *d_key[2] = {nullptr, nullptr}, *d_val[2] = {nullptr, nullptr};
    void *d_cub = nullptr; size_t cub_bytes = 0;
    RsTask *d_tasks = nullptr; int2 *d_deps = nullptr; size_t tasks_cap = 0, deps_cap = 0;
    int nt_max = 0;

    bool multi() const { return h->nranks > 1; }
    int allreduce(void *buf, size_t n, ncclDataType_t t, ncclRedOp_t op)
    {
        if (!multi()) return ORBS_OK;
        ORBS_NCCL(g_nccl.AllReduce(buf, buf, n, t, op, h->comm, st));
        return ORBS_OK;
    }
    // the per-trial exchanges: NVLink peer memory (peer_reduce.cuh) when the ranks could map each other's buffers, NCCL otherwise
    int exchange_system()
    {
        if (!multi()) return ORBS_OK;
        const size_t n = (size_t)B.ns * TS2 + (size_t)B.nt * TS;
        if (!h->peer_on) return allreduce(B.A, n, ncclDouble, ncclSum);
        const unsigned ep = ++h->peer_epoch;
        // the sums go to the private copy the solver reads (rsbuf.A / rsbuf.b); the shared buffer keeps this rank's partial until the next trial
        double *out = const_cast<double *>(rsbuf.A);
        if (h->nranks <= 4) {
            const int grid = std::max(8, std::min(h->sm_count, (int)(n / 2 / 512) + 1));
            k_peer_all_reduce<<<grid, 256, 0, st>>>(h->peer, B.ctl, (long long)(n / 2), ep, out);
            count();
        } else {
            const int grid = std::max(8, std::min(h->sm_count, (int)(n / 2 / h->nranks / 256) + 1));
            k_peer_reduce_scatter<<<grid, 256, 0, st>>>(h->peer, B.ctl, (long long)(n / 2), ep);
            k_peer_all_gather<<<grid, 256, 0, st>>>(h->peer, B.ctl, (long long)(n / 2), ep, out);
            count(2);
        }
        ORBS_CUDA(cudaGetLastError());
        return ORBS_OK;
    }
    int exchange_scalars()
    {
        if (!multi()) return ORBS_OK;
        if (!h->peer_on) return allreduce(B.scalars, 5, ncclDouble, ncclSum);
        k_peer_scalars<<<1, 32, 0, st>>>(h->peer, B.ctl, B.scalars, ++h->peer_sepoch);
        count();
        ORBS_CUDA(cudaGetLastError());
        return ORBS_OK;
    }
    KernelTimer &T() { return h->timer; }
    void count(int n = 1) { h->launches += n; }

    // pose activity from the device (after a level change) -> hessian indices; returns 1 if any edge is active, < 0 on error
    int read_activity(bool *changed)
    {
        ORBS_CUDA(cudaMemsetAsync(d_pose_act, 0, (B.K + 1) * sizeof(int), st));
        k_ba_activity<<<(B.P + 255) / 256, 256, 0, st>>>(B, d_pose_act);
        count();
        if (int rc = allreduce(d_pose_act, B.K + 1, ncclInt32, ncclMax)) return -1 - 0 * rc;
        std::vector<int> act(B.K + 1);
        ORBS_CUDA(cudaMemcpyAsync(act.data(), d_pose_act, (B.K + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        std::vector<int> idx(B.K);
        int nA = 0;
        for (int k = 0; k < B.K; k++) idx[k] = (act[k] && !fixed[k]) ? nA++ : -1;
        *changed = idx != pose_idx;
        pose_idx.swap(idx);
        B.nA = nA; B.n = 6 * nA;
        return act[B.K] ? 1 : 0;
    }

    // structure of the reduced system for the current pose_idx / edge levels
    int build_structure()
    {
        const int nA = B.nA, ng = (nA + kPosesPerTile - 1) / kPosesPerTile;
        B.nt = ng;
        ORBS_CUDA(cudaMemcpyAsync(d_pose_idx, pose_idx.data(), B.K * sizeof(int), cudaMemcpyHostToDevice, st));
        if (ng == 0) { B.ns = 0; n_seg = 0; return ORBS_OK; }
        ORBS_CUDA(cudaMemsetAsync(d_adj, 0, (size_t)ng * ng, st));
        k_ba_tile_adj<<<(B.P + 255) / 256, 256, 0, st>>>(B, d_adj, ng);
        count();
        if (int rc = allreduce(d_adj, (size_t)ng * ng, ncclUint8, ncclMax)) return rc;
        std::vector<uint8_t> adj((size_t)ng * ng);
        ORBS_CUDA(cudaMemcpyAsync(adj.data(), d_adj, adj.size(), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        if (trace) trace->mark("tile_adjacency");
        plan.build(adj, ng);
        if (trace) trace->mark("tile_plan");
        B.ns = plan.ns;
        ORBS_REQUIRE(plan.ns < (1 << 24), ORBS_E_INVALID, "reduced pose system too large / too dense for the tiled Cholesky (more than 2^24 tiles)");
        std::vector<int> rowbase(std::max(nA, 1)), tile_pose((size_t)ng * kPosesPerTile, -1);
        for (int ip = 0; ip < nA; ip++) {
            const int t = plan.pos[ip / kPosesPerTile];
            rowbase[ip] = TS * t + 6 * (ip % kPosesPerTile);
            tile_pose[(size_t)t * kPosesPerTile + ip % kPosesPerTile] = ip;
        }
        if (plan.tasks.size() > tasks_cap || plan.deps.size() + 1 > deps_cap) {
            tasks_cap = plan.tasks.size() + 64; deps_cap = plan.deps.size() + 256;
            if (int rc = h->ba_tasks.reserve(tasks_cap * sizeof(RsTask) + deps_cap * sizeof(int2))) return rc;
        }
        d_tasks = h->ba_tasks.as<RsTask>(); d_deps = reinterpret_cast<int2 *>(d_tasks + tasks_cap);
        ORBS_CUDA(cudaMemcpyAsync(d_tasks, plan.tasks.data(), plan.tasks.size() * sizeof(RsTask), cudaMemcpyHostToDevice, st));
        if (!plan.deps.empty()) ORBS_CUDA(cudaMemcpyAsync(d_deps, plan.deps.data(), plan.deps.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_rowbase, rowbase.data(), std::max(nA, 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_slot_of, plan.slot_of.data(), plan.slot_of.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_tile_pose, tile_pose.data(), tile_pose.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        rsplan.tasks = d_tasks; rsplan.ntasks = (int)plan.tasks.size(); rsplan.deps = d_deps; rsplan.nt = ng; rsplan.ns = plan.ns;
        solve_ctas = std::min(rsplan.ntasks, h->sm_count);
        // reduced-system storage depends on ns: [A | bs] contiguous (one all-reduce), L, Linv, b/y/x vectors, flags
        {
            const size_t nA_t = (size_t)plan.ns * TS2, nv = (size_t)ng * TS;
            if (int rc = h->ba_sys.reserve((2 * nA_t + (size_t)ng * TS2 + 3 * nv + (size_t)plan.n_part * TS2 + 16) * sizeof(double))) return rc;
            double *base = h->ba_sys.as<double>();
            B.A = base; B.bs = base + nA_t;
            if (multi() && h->peer_on) {
                // [A | b] lives in the shareable buffer the peers have mapped (the first part of ba_sys stays unused)
                if (int rc = peer_reserve_sys(h, (nA_t + nv) * sizeof(double))) return rc;
                if (h->peer_on) { B.A = h->peer.sys[h->rank]; B.bs = B.A + nA_t; }
            }
            rsbuf.A = B.A; rsbuf.b = B.bs;
            if (multi() && h->peer_on) { rsbuf.A = base; rsbuf.b = base + nA_t; }       // peer exchange: the solver reads the summed copy in ba_sys
            rsbuf.L = base + nA_t + nv; rsbuf.Linv = rsbuf.L + nA_t; rsbuf.y = rsbuf.Linv + (size_t)ng * TS2; rsbuf.x = rsbuf.y + nv;
            rsbuf.part = rsbuf.x + nv;
            B.x = rsbuf.x;
            const size_t nflags = (size_t)plan.ns + 2 * (size_t)ng + (size_t)plan.n_part + 8;
            const bool fresh = nflags * sizeof(int) > h->ba_flags.bytes;
            if (int rc = h->ba_flags.reserve(nflags * sizeof(int))) return rc;
            if (fresh) { ORBS_CUDA(cudaMemsetAsync(h->ba_flags.p, 0, h->ba_flags.bytes, st)); }
            int *f = h->ba_flags.as<int>();
            rsbuf.counters = f; rsbuf.flags = f + 4; rsbuf.done_slot = f + 8; rsbuf.done_y = rsbuf.done_slot + plan.ns; rsbuf.done_x = rsbuf.done_y + ng;
            rsbuf.done_part = rsbuf.done_x + ng;
            B.flags = rsbuf.flags;
            ORBS_CUDA(cudaMemsetAsync(f, 0, 8 * sizeof(int), st));
        }
        xp_blocks = std::max(1, (B.n + 255) / 256); B.n_xp = xp_blocks;
        diag_blocks = (nA + B.P + 255) / 256; B.n_diag = diag_blocks;
        // Schur pairs, sorted by target block
        k_ba_pair_count<<<(B.P + 255) / 256, 256, 0, st>>>(B, d_pair_cnt);
        ORBS_CUDA(cudaMemsetAsync(d_pair_cnt + B.P, 0, sizeof(int), st));
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, d_pair_cnt, d_pair_start, B.P + 1, st);
        size_t need2 = 0;
        const int end_bit = 32 + 7 + std::max(1, 32 - __builtin_clz((unsigned)std::max(plan.ns, 1)));
        cub::DeviceRadixSort::SortPairs(nullptr, need2, d_key[0], d_key[1], d_val[0], d_val[1], (int)pair_cap, 0, std::min(end_bit, 64), st);
        size_t need3 = 0;
        cub::DeviceScan::InclusiveSum(nullptr, need3, d_head, d_scan, (int)pair_cap, st);
        need = std::max({need, need2, need3});
        if (int rc = h->ba_cub.reserve(need + 256)) return rc;
        d_cub = h->ba_cub.p; cub_bytes = h->ba_cub.bytes;
        cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_pair_cnt, d_pair_start, B.P + 1, st);
        ORBS_CUDA(cudaMemsetAsync(d_key[0], 0xff, (size_t)pair_cap * sizeof(unsigned long long), st));
        k_ba_pair_emit<<<(B.P + 256) / 256, 256, 0, st>>>(B, d_pair_start, d_key[0], d_val[0]);
        cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_key[0], d_key[1], d_val[0], d_val[1], (int)pair_cap, 0, std::min(end_bit, 64), st);
        k_ba_seg_heads<<<(unsigned)((pair_cap + 255) / 256), 256, 0, st>>>(d_key[1], d_pair_start + B.P, d_head, (int)pair_cap);
        cub::DeviceScan::InclusiveSum(d_cub, cub_bytes, d_head, d_scan, (int)pair_cap, st);
        k_ba_seg_starts<<<(unsigned)((pair_cap + 255) / 256), 256, 0, st>>>(d_head, d_scan, d_pair_start + B.P, d_seg_start, d_nseg, (int)pair_cap);
        count(8);
        ORBS_CUDA(cudaGetLastError());
        ORBS_CUDA(cudaMemcpyAsync(&n_seg, d_nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        if (trace) trace->mark("pair_sort");
        B.pair_key = d_key[1]; B.pair_val = d_val[1]; B.seg_start = d_seg_start; B.n_seg = d_nseg;
        // work items of the Schur assembly: segments cut into pieces of kItemPairs pairs (d_head / d_scan are free again)
        n_items = 0;
        if (n_seg > 0) {
            k_ba_seg_items<<<(n_seg + 256) / 256, 256, 0, st>>>(d_seg_start, d_nseg, d_head, n_seg + 1);
            cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_head, d_scan, n_seg + 1, st);
            count(2);
            ORBS_CUDA(cudaMemcpyAsync(&n_items, d_scan + n_seg, sizeof(int), cudaMemcpyDeviceToHost, st));
            ORBS_CUDA(cudaStreamSynchronize(st));
            if (int rc = h->ba_items.reserve((size_t)std::max(n_items, 1) * kSchurVals * sizeof(double))) return rc;
        }
        if (trace) trace->mark("schur_items");
        return ORBS_OK;
    }

    void errors()
    {
        T().begin(BK_ERRORS, st);
        k_ba_errors<<<err_blocks, 256, 0, st>>>(B);
        T().end(st);
        count();
    }

    // one LM slot (see the header comment); `first`: slot 0 of an optimize() call carries the lambda initialisation
    int enqueue_slot(bool first)
    {
        T().begin(BK_BUILD_POINTS, st);
        k_ba_build_points<<<pt_blocks, 256, 0, st>>>(B);
        T().end(st);
        T().begin(BK_BUILD_POSES, st);
        k_ba_build_poses<<<B.K, 128, 0, st>>>(B);
        T().end(st);
        count(2);
        if (first) {
            k_ba_max_diag<<<diag_blocks, 256, 0, st>>>(B, diag6);
            k_ba_max_diag_finish<<<1, 256, 0, st>>>(B, diag6 + B.n);
            count(2);
            if (int rc = allreduce(diag6, (size_t)B.n + h->nranks, ncclDouble, ncclSum)) return rc;
        }
        if (first) { k_lm_iter_begin<<<1, 1, 0, st>>>(B, diag6, diag6 + B.n); count(); }     // later iterations are prepared by the decide step of the previous one
        T().begin(BK_PREP, st);
        k_ba_point_prep<<<pt_blocks, 256, 0, st>>>(B);
        T().end(st);
        count(2);
        if (B.n > 0) {
            T().begin(BK_SCHUR, st);
            ORBS_CUDA(cudaMemsetAsync(B.A, 0, ((size_t)B.ns * TS2 + (size_t)B.nt * TS) * sizeof(double), st));
            k_ba_diag_init<<<(B.nt * TS + 255) / 256, 256, 0, st>>>(B);
            if (n_items > 0) {
                k_ba_schur_items<<<n_items, 128, 0, st>>>(B, d_scan, h->ba_items.as<double>());
                k_ba_schur_finish<<<(n_seg + 7) / 8, 256, 0, st>>>(B, d_scan, h->ba_items.as<double>());
            }
            T().end(st);
            count(3);
            // the one exchange step of the sharded solve: the structurally nonzero tiles of the reduced system and its right-hand side, summed over the map-point shards
            T().begin(BK_EXCHANGE, st);
            if (int rc = exchange_system()) return rc;
            T().end(st);
            T().begin(BK_SOLVE, st);
            k_rs_solve<<<solve_ctas, 256, kRsSmemBytes, st>>>(rsplan, rsbuf, B.ctl, ++h->rs_epoch);
            T().end(st);
            count();
        }
        T().begin(BK_BACKSUB, st);
        k_ba_backsub<<<std::max(pt_blocks, xp_blocks), 256, 0, st>>>(B);
        T().end(st);
        T().begin(BK_UPDATE, st);
        k_ba_update<<<upd_blocks, 256, 0, st>>>(B);
        T().end(st);
        count(2);
        errors();
        T().begin(BK_DECIDE, st);
        if (!multi()) {
            k_lm_reduce<<<1, 256, 0, st>>>(B, h_stop, 2);                   // sums + decision in one launch
        } else {
            k_lm_reduce<<<1, 256, 0, st>>>(B, h_stop, 1);
            T().end(st);
            T().begin(BK_EXCHANGE, st);
            if (int rc = exchange_scalars()) return rc;
            T().end(st);
            T().begin(BK_DECIDE, st);
            k_lm_decide<<<1, 1, 0, st>>>(B);
            count();
        }
        k_ba_restore<<<upd_blocks, 256, 0, st>>>(B);
        T().end(st);
        count(2);
        ORBS_CUDA(cudaGetLastError());
        return ORBS_OK;
    }

    void poll_stop() { if (stop && *stop) *h_stop = 1; }

    // SparseOptimizer::optimize(iterations), sparse_optimizer.cpp:354-419
    int optimize(int iterations, bool any_active, LmCtl *final_ctl)
    {
        LmCtl c0;
        ORBS_CUDA(cudaMemcpyAsync(&c0, B.ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        poll_stop();
        const bool run = any_active && iterations > 0 && !*h_stop;
        c0.state = run ? LM_BUILD : LM_DONE; c0.iteration = 0; c0.max_iterations = iterations; c0.qmax = 0; c0.nbad = 0; c0.first = 1; c0.last_rejected = 0;
        ORBS_CUDA(cudaMemcpyAsync(B.ctl, &c0, sizeof(LmCtl), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaStreamSynchronize(st));                       // c0 lives on this stack frame
        if (run) {
            // computeActiveErrors + activeRobustChi2 of the starting estimate (levenberg.cpp:69-72); later iterations inherit them from the accepted trial
            errors();
            k_lm_reduce<<<1, 256, 0, st>>>(B, h_stop, 0);
            count();
            if (int rc = exchange_scalars()) return rc;
            constexpr int kLag = 2;
            bool done = false;
            int slot = 0;
            for (; slot < iterations && !done; slot++) {
                poll_stop();
                if (int rc = enqueue_slot(slot == 0)) return rc;
                ORBS_CUDA(cudaMemcpyAsync(&h_ctl[slot % orbo_handle::kCtlCopies], B.ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, st));
                ORBS_CUDA(cudaEventRecord(h->slot_done[slot % orbo_handle::kCtlCopies], st));
                if (slot >= kLag) {
                    const int q = (slot - kLag) % orbo_handle::kCtlCopies;
                    while (cudaEventQuery(h->slot_done[q]) == cudaErrorNotReady) poll_stop();
                    if (h_ctl[q].state == LM_DONE) done = true;     // early stop (nbad rule, stop flag, failed trials): the slots already enqueued are no-ops
                }
            }
            // the common case ends here: `iterations` accepted trials.  Rejected trials need further slots, one at a time.
            for (int guard = 0; guard < 10 * iterations + 4; guard++) {
                const int q = (slot - 1) % orbo_handle::kCtlCopies;
                while (cudaEventQuery(h->slot_done[q]) == cudaErrorNotReady) poll_stop();
                if (h_ctl[q].state == LM_DONE) break;
                poll_stop();
                if (int rc = enqueue_slot(false)) return rc;
                ORBS_CUDA(cudaMemcpyAsync(&h_ctl[slot % orbo_handle::kCtlCopies], B.ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, st));
                ORBS_CUDA(cudaEventRecord(h->slot_done[slot % orbo_handle::kCtlCopies], st));
                slot++;
            }
        }
        ORBS_CUDA(cudaMemcpyAsync(&h_ctl[0], B.ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        *final_ctl = h_ctl[0];
        return ORBS_OK;
    }
};

}  // namespace

extern "C" int orbo_bundle_adjust(orbo_handle *h, int K, float *poses, const uint8_t *fixed, const double *intr, int P, float *points,
                                  int E, const int32_t *e_kf, const int32_t *e_pt, const float *e_uv, const float *e_inv_sigma2,
                                  int two_stage, int its0, int its1, int robust, const volatile uint8_t *stop_flag,
                                  double *e_chi2, uint8_t *e_depth_ok, uint8_t *e_outlier, int32_t *stats)
{
    ORBS_REQUIRE(h && poses && fixed && intr && points && e_kf && e_pt && e_uv && e_inv_sigma2, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(K > 0 && P > 0 && E > 0, ORBS_E_INVALID, "empty graph");
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (h->nranks == 1 && stop_flag && *stop_flag) { if (stats) stats[3] = 1; return 1; }   // Optimizer.cc:678-680 (sharded: decided collectively below)
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const auto t_begin = std::chrono::steady_clock::now();
    StageTrace trace;
    trace.mark("begin");

    // ---- host-side graph layout: edges grouped by point (stable); every array that goes to the device is carved out of one pinned arena.  The float32 inputs
    // (observations, weights, points) are staged as they are and widened to fp64 on the device
    const bool want_edges = e_chi2 || e_depth_ok || e_outlier;
    auto up = [](size_t b) { return (b + 63) & ~(size_t)63; };
    const size_t arena = up((P + 1) * sizeof(int)) + 2 * up((size_t)E * sizeof(int)) + up((K + 1) * sizeof(int)) + up(2 * (size_t)E * sizeof(float)) +
                         up((size_t)E * sizeof(float)) + up(3 * (size_t)P * sizeof(float)) + (want_edges ? up((size_t)E * sizeof(double)) + 2 * up((size_t)E) : 0) + 64;
    if (int rc = h->ba_host.reserve(arena)) return rc;
    uint8_t *ap = h->ba_host.as<uint8_t>();
    auto carve = [&](size_t bytes) { uint8_t *p = ap; ap += up(bytes); return p; };
    int *pt_start = (int *)carve((P + 1) * sizeof(int)), *kf_s = (int *)carve((size_t)E * sizeof(int)), *pt_s = (int *)carve((size_t)E * sizeof(int));
    int *pose_start = (int *)carve((K + 1) * sizeof(int));
    float *obs_s = (float *)carve(2 * (size_t)E * sizeof(float)), *w_s = (float *)carve((size_t)E * sizeof(float)), *pts_s = (float *)carve(3 * (size_t)P * sizeof(float));
    double *chi2_s = want_edges ? (double *)carve((size_t)E * sizeof(double)) : nullptr;
    uint8_t *depth_s = want_edges ? carve((size_t)E) : nullptr, *outl_s = want_edges ? carve((size_t)E) : nullptr;
    memset(pt_start, 0, (P + 1) * sizeof(int)); memset(pose_start, 0, (K + 1) * sizeof(int));
    trace.mark("arena");
    // pass A: validate + count per point and per keyframe;  pass B: stable scatter into point order (a straight conversion copy when the caller's edges
    // already come grouped by ascending point, as a graph built by walking the map points does);  the second CSR (edges of a keyframe, ascending) is
    // a stable device sort of the uploaded keyframe column (below)
    bool grouped = true;
    {
        bool ok = true;
        int prev = 0;
        for (int e = 0; e < E; e++) {
            const int kf = e_kf[e], pt = e_pt[e];
            if ((unsigned)kf >= (unsigned)K || (unsigned)pt >= (unsigned)P) { ok = false; break; }
            grouped &= pt >= prev; prev = pt;
            pt_start[pt + 1]++; pose_start[kf + 1]++;
        }
        ORBS_REQUIRE(ok, ORBS_E_INVALID, "edge references a vertex out of range");
    }
    long long pair_cap = 0;
    for (int p = 0; p < P; p++) { const long long m = pt_start[p + 1]; pair_cap += m * (m + 1) / 2; pt_start[p + 1] += pt_start[p]; }
    ORBS_REQUIRE(pair_cap < (1ll << 31), ORBS_E_INVALID, "too many co-observations for one bundle adjustment (2^31 keyframe pairs)");
    pair_cap = std::max(pair_cap, 1ll);
    for (int k = 0; k < K; k++) pose_start[k + 1] += pose_start[k];
    trace.mark("pass_A_count");
    std::vector<int> order;                                                  // sorted position -> caller's edge index (empty: identity)
    if (grouped) {
        // (splitting these copies over helper threads was measured: the copy gets faster, the DMA out of lines owned by several cores slower by the same amount)
        memcpy(kf_s, e_kf, (size_t)E * sizeof(int)); memcpy(pt_s, e_pt, (size_t)E * sizeof(int));
        memcpy(obs_s, e_uv, 2 * (size_t)E * sizeof(float)); memcpy(w_s, e_inv_sigma2, (size_t)E * sizeof(float));
    } else {
        order.resize(E);
        std::vector<int> fill(pt_start, pt_start + P);
        for (int e = 0; e < E; e++) {
            const int pt = e_pt[e], j = fill[pt]++;
            order[j] = e; kf_s[j] = e_kf[e]; pt_s[j] = pt;
            obs_s[2 * j] = e_uv[2 * e]; obs_s[2 * j + 1] = e_uv[2 * e + 1]; w_s[j] = e_inv_sigma2[e];
        }
    }
    memcpy(pts_s, points, 3 * (size_t)P * sizeof(float));
    trace.mark("pass_B_scatter");

    // ---- device buffers
    Stager S(&h->ba_pool, st, ORBS_MEM_HOST);
    BaRun D;
    D.h = h; D.st = st; D.S = &S; D.stop = stop_flag; D.fixed = fixed; D.trace = &trace;
    D.h_ctl = h->h_scalars.as<LmCtl>(); D.h_stop = reinterpret_cast<int *>(D.h_ctl + orbo_handle::kCtlCopies);
    *D.h_stop = 0;
    BaDev &B = D.B;
    memset(&B, 0, sizeof B);
    B.K = K; B.P = P; B.E = E;
    B.nranks = h->nranks; B.rank = h->rank; B.lead = h->rank == 0 ? 1 : 0;
    const float *d_T = S.in(poses, (size_t)K * 16);
    const uint8_t *d_fixed = S.in(fixed, K);
    B.intr = S.in(intr, (size_t)K * 4);
    const float *d_pts_f = S.in(pts_s, 3 * (size_t)P);
    B.pt = S.scratch<double>(3 * (size_t)P);
    B.pt_bak = S.scratch<double>(3 * (size_t)P);
    B.pose = S.scratch<Se3>(K); B.pose_bak = S.scratch<Se3>(K);
    B.pt_start = S.in(pt_start, (size_t)P + 1); B.e_kf = S.in(kf_s, E); B.e_point = S.in(pt_s, E);
    const float *d_obs_f = S.in(obs_s, 2 * (size_t)E), *d_w_f = S.in(w_s, E);
    double *d_obs = S.scratch<double>(2 * (size_t)E), *d_w = S.scratch<double>(E);
    B.e_obs = d_obs; B.e_w = d_w;
    B.e_level = S.scratch<uint8_t>(E); B.e_err = S.scratch<double>(2 * (size_t)E);
    B.e_W = S.scratch<double>(18 * (size_t)E); B.e_Z = S.scratch<double>(18 * (size_t)E);
    B.pose_start = S.in(pose_start, (size_t)K + 1);
    int *d_pose_edges = S.scratch<int>(E), *d_iota = S.scratch<int>(E), *d_kf_sorted = S.scratch<int>(E);
    B.pose_edges = d_pose_edges;
    D.d_pose_idx = S.scratch<int>(K); B.pose_idx = D.d_pose_idx;
    B.pt_active = S.scratch<uint8_t>(P);
    B.Hpp = S.scratch<double>(36 * (size_t)K); B.bp = S.scratch<double>(6 * (size_t)K + 8);
    B.Hll = S.scratch<double>(9 * (size_t)P); B.bl = S.scratch<double>(3 * (size_t)P);
    B.ptG = S.scratch<double>(6 * (size_t)P); B.ptg = S.scratch<double>(3 * (size_t)P); B.xl = S.scratch<double>(3 * (size_t)P);
    D.nt_max = (K + kPosesPerTile - 1) / kPosesPerTile;
    ORBS_REQUIRE(D.nt_max <= 1024, ORBS_E_INVALID, "more than 10240 keyframes in one bundle adjustment");
    D.d_rowbase = S.scratch<int>(K + 1); B.rowbase = D.d_rowbase;
    D.d_slot_of = S.scratch<int>((size_t)D.nt_max * D.nt_max + 4); B.slot_of = D.d_slot_of;
    D.d_tile_pose = S.scratch<int>((size_t)D.nt_max * kPosesPerTile + 4); B.tile_pose = D.d_tile_pose;
    D.d_adj = S.scratch<uint8_t>((size_t)D.nt_max * D.nt_max + 16);
    D.d_pose_act = S.scratch<int>(K + 2);
    D.err_blocks = (E + 255) / 256; D.pt_blocks = (P + 31) / 32; D.upd_blocks = (K + P + 255) / 256;
    B.n_chi = D.err_blocks; B.n_pt = D.pt_blocks;
    B.p_chi = S.scratch<double>(D.err_blocks + (size_t)D.pt_blocks + (6 * (size_t)K + 255) / 256 + 1 + ((size_t)K + P + 255) / 256 + 1);
    B.p_pt = B.p_chi + D.err_blocks; B.p_xp = B.p_pt + D.pt_blocks; B.p_diag = B.p_xp + (6 * (size_t)K + 255) / 256 + 1;
    B.scalars = S.scratch<double>(16 + sizeof(LmCtl) / sizeof(double) + 2);
    B.ctl = reinterpret_cast<LmCtl *>(B.scalars + 16);
    D.diag6 = S.scratch<double>(6 * (size_t)K + 64);
    D.pair_cap = pair_cap;
    D.d_pair_cnt = S.scratch<int>(P + 2); D.d_pair_start = S.scratch<int>(P + 2);
    D.d_key[0] = S.scratch<unsigned long long>(pair_cap); D.d_key[1] = S.scratch<unsigned long long>(pair_cap);
    D.d_val[0] = S.scratch<unsigned long long>(pair_cap); D.d_val[1] = S.scratch<unsigned long long>(pair_cap);
    D.d_head = S.scratch<int>(pair_cap); D.d_scan = S.scratch<int>(pair_cap); D.d_seg_start = S.scratch<int>(pair_cap + 2); D.d_nseg = S.scratch<int>(4);
    double *d_chi2 = S.scratch<double>(E); uint8_t *d_depth = S.scratch<uint8_t>(E), *d_outl = S.scratch<uint8_t>(E);
    float *d_Tout = S.scratch<float>((size_t)K * 16); float *d_pts_out = S.scratch<float>(3 * (size_t)P);
    if (S.rc) return S.rc;
    {
        const size_t nw = std::max(2 * (size_t)E, 3 * (size_t)P);
        k_ba_widen<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(2 * (size_t)E, d_obs_f, d_obs, (size_t)E, d_w_f, d_w, 3 * (size_t)P, d_pts_f, B.pt);
        h->launches++;
    }
    {   // edges of every keyframe in ascending (point-grouped) edge order: stable radix sort of the keyframe column, values = edge positions
        const int kbits = std::max(1, 32 - __builtin_clz((unsigned)std::max(K - 1, 1)));
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, B.e_kf, d_kf_sorted, d_iota, d_pose_edges, E, 0, kbits, st);
        if (int rc = h->ba_cub.reserve(need + 256)) return rc;
        k_ba_iota<<<(E + 255) / 256, 256, 0, st>>>(E, d_iota);
        ORBS_CUDA(cub::DeviceRadixSort::SortPairs(h->ba_cub.p, need, B.e_kf, d_kf_sorted, d_iota, d_pose_edges, E, 0, kbits, st));
        h->launches += 2;
    }
    // LocalBundleAdjustment: const float thHuberMono = sqrt(5.991) (Optimizer.cc:592); BundleAdjustment: const float thHuber2D = sqrt(5.99) (:106)
    B.delta = two_stage ? (double)(float)sqrt(5.991) : (double)(float)sqrt(5.99);
    B.dsqr = (double)(float)(B.delta * B.delta);   // RobustKernelHuber keeps delta^2 in a FLOAT member (robust_kernel_impl.h:84): pinned against the reference's object code
    ORBS_CUDA(cudaMemsetAsync(B.e_level, 0, E, st));
    ORBS_CUDA(cudaMemsetAsync(B.e_err, 0, 2 * (size_t)E * sizeof(double), st));
    ORBS_CUDA(cudaMemsetAsync(B.scalars, 0, (16 + sizeof(LmCtl) / sizeof(double) + 2) * sizeof(double), st));
    k_ba_import_poses<<<(K + 255) / 256, 256, 0, st>>>(K, d_T, B.pose);
    ORBS_CUDA(cudaMemcpyAsync(d_Tout, d_T, (size_t)K * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    h->launches++;

    // sharded: the stop decision is collective (every rank must take it or none)
    if (D.multi()) {
        int v = (stop_flag && *stop_flag) ? 1 : 0;
        ORBS_CUDA(cudaMemcpyAsync(D.d_pose_act, &v, sizeof(int), cudaMemcpyHostToDevice, st));
        if (int rc = D.allreduce(D.d_pose_act, 1, ncclInt32, ncclMax)) return rc;
        ORBS_CUDA(cudaMemcpyAsync(&v, D.d_pose_act, sizeof(int), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        if (v) { if (stats) stats[3] = 1; return 1; }
    }
    // initializeOptimization(level 0) + buildIndexMapping, sparse_optimizer.cpp:166-267
    D.pose_idx.assign(K, -2);
    bool changed = false;
    trace.mark("enqueue_uploads");
    int any = D.read_activity(&changed);
    if (any < 0) return ORBS_E_CUDA;
    trace.mark("uploads_and_activity");
    int rc = D.build_structure();
    if (rc) return rc;
    const auto t_loop = std::chrono::steady_clock::now();
    B.robust = two_stage ? 1 : (robust ? 1 : 0);
    LmCtl fin;
    memset(&fin, 0, sizeof fin);
    if ((rc = D.optimize(its0, any == 1, &fin))) return rc;
    if (two_stage && !*D.h_stop) {
        // chi2 / depth gating (Optimizer.cc:691-705) on the device; the structure is rebuilt only if a keyframe lost all its observations
        k_ba_edge_check<<<(E + 255) / 256, 256, 0, st>>>(B, d_chi2, d_depth, d_outl, 1);
        h->launches++;
        B.robust = 0;
        any = D.read_activity(&changed);
        if (any < 0) return ORBS_E_CUDA;
        if (changed && (rc = D.build_structure())) return rc;
        if ((rc = D.optimize(its1, any == 1, &fin))) return rc;
    }
    if (const char *tp = getenv("ORBS_RS_TRACE")) {
        // profiling aid: one more solve of the last system with per-task timestamps, dumped as text (task type i j ndep t_start t_deps t_end sm)
        const int ntk = D.rsplan.ntasks;
        long long *d_tr = nullptr;
        if (ntk > 0 && cudaMalloc(&d_tr, (size_t)ntk * 4 * sizeof(long long)) == cudaSuccess) {
            cudaMemsetAsync(d_tr, 0, (size_t)ntk * 4 * sizeof(long long), st);
            RsBuf rb = D.rsbuf; rb.trace = d_tr;
            k_rs_solve<<<D.solve_ctas, 256, kRsSmemBytes, st>>>(D.rsplan, rb, nullptr, ++h->rs_epoch);
            std::vector<long long> tr((size_t)ntk * 4);
            cudaMemcpyAsync(tr.data(), d_tr, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            cudaFree(d_tr);
            if (FILE *f = fopen(tp, "w")) {
                for (int t = 0; t < ntk; t++) { const RsTask &k = D.plan.tasks[t]; fprintf(f, "%d %d %d %d %d %lld %lld %lld %lld\n", t, k.type, k.i, k.j, k.ndep, tr[4 * t], tr[4 * t + 1], tr[4 * t + 2], tr[4 * t + 3]); }
                fclose(f);
            }
        }
    }
    const auto t_loop_end = std::chrono::steady_clock::now();          // optimize() returned after a stream synchronisation
    trace.mark("lm_loops");
    k_ba_edge_check<<<(E + 255) / 256, 256, 0, st>>>(B, d_chi2, d_depth, d_outl, 0);
    k_ba_export_poses<<<(K + 255) / 256, 256, 0, st>>>(K, B.pose, d_fixed, d_Tout);
    k_ba_export_points<<<(3 * P + 255) / 256, 256, 0, st>>>(P, B.pt, d_pts_out);
    h->launches += 3;
    if (want_edges) {
        ORBS_CUDA(cudaMemcpyAsync(chi2_s, d_chi2, E * sizeof(double), cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaMemcpyAsync(depth_s, d_depth, E, cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaMemcpyAsync(outl_s, d_outl, E, cudaMemcpyDeviceToHost, st));
    }
    ORBS_CUDA(cudaMemcpyAsync(poses, d_Tout, (size_t)K * 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaMemcpyAsync(points, d_pts_out, 3 * (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaStreamSynchronize(st));
    trace.mark("download");
    if (want_edges) {                                                                         // Optimizer.cc:734-766
        if (order.empty()) {
            if (e_chi2) memcpy(e_chi2, chi2_s, (size_t)E * sizeof(double));
            if (e_depth_ok) memcpy(e_depth_ok, depth_s, (size_t)E);
            if (e_outlier) memcpy(e_outlier, outl_s, (size_t)E);
        } else
            for (int j = 0; j < E; j++) {
                const int e = order[j];
                if (e_chi2) e_chi2[e] = chi2_s[j];
                if (e_depth_ok) e_depth_ok[e] = depth_s[j];
                if (e_outlier) e_outlier[e] = outl_s[j];
            }
    }
    {
        const auto t_end = std::chrono::steady_clock::now();
        auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        h->ba_timing[0] = sec(t_loop, t_loop_end); h->ba_timing[1] = sec(t_begin, t_end); h->ba_timing[2] = sec(t_begin, t_loop);
        h->ba_timing[3] = (double)B.nt * TS;
        h->ba_skyline[0] = B.nt; h->ba_skyline[1] = B.ns; h->ba_skyline[2] = D.plan.nlevels;
    }
    trace.mark("edge_flags");
    trace.dump();
    if (stats) { stats[0] = fin.lm_iterations; stats[1] = fin.lm_trials; stats[2] = fin.chol_failures; stats[3] = 0; }
    if (D.multi() && h->peer_on) {
        int perr = 0;
        ORBS_CUDA(cudaMemcpy(&perr, &h->peer.ctl[h->rank]->error, sizeof(int), cudaMemcpyDeviceToHost));
        ORBS_REQUIRE(perr == 0, ORBS_E_CUDA, "sharded bundle adjustment: a rank did not arrive at a peer-memory exchange (timed out waiting for its flag)");
    }
    return ORBS_OK;
}

extern "C" int orbo_comm_mode(const orbo_handle *h)
{
    return !h || h->nranks <= 1 ? 0 : (h->peer_on ? 2 : 1);            // 0 single GPU, 1 NCCL all-reduce, 2 NVLink peer-memory exchange
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
// Essential-graph (Sim3 pose graph) optimisation: host-driven Levenberg over the kernels of pose_graph.cuh, normal equations solved by
// the tiled sparse solver of the BA (tile_solver.cuh; own symbolic factorisation: 9 Sim3 vertices per 64-row tile)
namespace {

struct PgHost {
    orbo_handle *h; cudaStream_t st; PgDev P;
    int nbuf = 0;
    TilePlanHost plan;
    RsPlan rsplan = {}; RsBuf rsbuf = {};
    int eblocks = 0, solve_ctas = 1;
    template <typename T> T *alloc(size_t n, int *rc) { if (*rc) return nullptr; if (nbuf >= 24) { *rc = ORBS_E_INVALID; return nullptr; } *rc = h->pg_bufs[nbuf].reserve(std::max<size_t>(n, 1) * sizeof(T)); return h->pg_bufs[nbuf++].as<T>(); }

    // tile groups of kPgPerTile consecutive free vertices; adjacency from the edges; elimination order + symbolic factorisation + tasks;
    // the per-target-block contribution lists of the normal equations (sorted: deterministic accumulation)
    int schedule(int K, const std::vector<int> &hidx, int E, const int32_t *e_i, const int32_t *e_j)
    {
        const int nA = P.nA;
        const int ng = (nA + kPgPerTile - 1) / kPgPerTile;
        std::vector<uint8_t> adj((size_t)ng * ng, 0);
        for (int e = 0; e < E; e++) {
            const int a = hidx[e_i[e]], b = hidx[e_j[e]];
            if (a < 0 || b < 0) continue;
            adj[(size_t)(a / kPgPerTile) * ng + b / kPgPerTile] = 1; adj[(size_t)(b / kPgPerTile) * ng + a / kPgPerTile] = 1;
        }
        plan.build(adj, ng);
        const int nt = plan.nt;
        P.nt = nt; P.ns = plan.ns;
        std::vector<int> rowbase(std::max(nA, 1));
        std::vector<uint8_t> rowpad((size_t)nt * TS, 1);
        for (int ip = 0; ip < nA; ip++) {
            rowbase[ip] = TS * plan.pos[ip / kPgPerTile] + 7 * (ip % kPgPerTile);
            for (int a = 0; a < 7; a++) rowpad[rowbase[ip] + a] = 0;
        }
        // contributions: (row global row, col global row, edge, side_row, side_col), sorted
        struct C { int rr, rc, e, sr, sc; };
        std::vector<C> cs; cs.reserve((size_t)E * 3);
        for (int e = 0; e < E; e++) {
            const int hs[2] = {hidx[e_i[e]], hidx[e_j[e]]};
            for (int s = 0; s < 2; s++) if (hs[s] >= 0) cs.push_back({rowbase[hs[s]], rowbase[hs[s]], e, s, s});
            if (hs[0] >= 0 && hs[1] >= 0) {
                const int r0 = rowbase[hs[0]], r1 = rowbase[hs[1]];
                if (r0 > r1) cs.push_back({r0, r1, e, 0, 1}); else cs.push_back({r1, r0, e, 1, 0});
            }
        }
        std::stable_sort(cs.begin(), cs.end(), [](const C &a, const C &b) { return a.rr != b.rr ? a.rr < b.rr : a.rc < b.rc; });
        std::vector<PgBlock> blocks; std::vector<PgEntry> entries(cs.size());
        for (size_t q = 0; q < cs.size(); q++) {
            entries[q] = {cs[q].e, cs[q].sr, cs[q].sc, 0};
            if (q == 0 || cs[q].rr != cs[q - 1].rr || cs[q].rc != cs[q - 1].rc) {
                PgBlock b = {};
                b.slot = plan.slot_of[(size_t)(cs[q].rr >> 6) * nt + (cs[q].rc >> 6)];
                b.r0 = cs[q].rr & 63; b.c0 = cs[q].rc & 63; b.diag = cs[q].rr == cs[q].rc; b.start = (int)q; b.n = 0; b.rs = cs[q].rr;
                blocks.push_back(b);
            }
            blocks.back().n++;
        }
        P.n_blocks = (int)blocks.size();
        int rc = ORBS_OK;
        RsTask *d_tasks = alloc<RsTask>(plan.tasks.size(), &rc); int2 *d_deps = alloc<int2>(plan.deps.size(), &rc);
        int *d_rowbase = alloc<int>(nA, &rc); uint8_t *d_rowpad = alloc<uint8_t>(rowpad.size(), &rc);
        PgBlock *d_blocks = alloc<PgBlock>(blocks.size(), &rc); PgEntry *d_entries = alloc<PgEntry>(entries.size(), &rc);
        if (rc) return rc;
        ORBS_CUDA(cudaMemcpyAsync(d_tasks, plan.tasks.data(), plan.tasks.size() * sizeof(RsTask), cudaMemcpyHostToDevice, st));
        if (!plan.deps.empty()) ORBS_CUDA(cudaMemcpyAsync(d_deps, plan.deps.data(), plan.deps.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_rowbase, rowbase.data(), std::max(nA, 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaMemcpyAsync(d_rowpad, rowpad.data(), rowpad.size(), cudaMemcpyHostToDevice, st));
        if (!blocks.empty()) ORBS_CUDA(cudaMemcpyAsync(d_blocks, blocks.data(), blocks.size() * sizeof(PgBlock), cudaMemcpyHostToDevice, st));
        if (!entries.empty()) ORBS_CUDA(cudaMemcpyAsync(d_entries, entries.data(), entries.size() * sizeof(PgEntry), cudaMemcpyHostToDevice, st));
        ORBS_CUDA(cudaStreamSynchronize(st));                  // the host vectors die here
        rsplan.tasks = d_tasks; rsplan.ntasks = (int)plan.tasks.size(); rsplan.deps = d_deps; rsplan.nt = nt; rsplan.ns = plan.ns;
        solve_ctas = std::max(1, std::min(rsplan.ntasks, h->sm_count));
        P.rowbase = d_rowbase; P.row_pad = d_rowpad; P.blocks = d_blocks; P.entries = d_entries;
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
    // A <- H + lambda I, factor, solve; ok = the factorisation succeeded
    int solve(double lambda, bool *ok)
    {
        ORBS_CUDA(cudaMemsetAsync(rsbuf.flags, 0, 4 * sizeof(int), st));
        ORBS_CUDA(cudaMemcpyAsync(P.A, P.H, (size_t)P.ns * TS2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        k_pg_prepare<<<(P.nt * TS + 255) / 256, 256, 0, st>>>(P, lambda);
        k_rs_solve<<<solve_ctas, 256, kRsSmemBytes, st>>>(rsplan, rsbuf, nullptr, ++h->rs_epoch);
        h->launches += 2;
        int f = 0;
        ORBS_CUDA(cudaMemcpyAsync(&f, rsbuf.flags, sizeof(int), cudaMemcpyDeviceToHost, st));
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
    memset(&P, 0, sizeof P);
    std::vector<int> hidx(K);
    int nA = 0;
    for (int k = 0; k < K; k++) hidx[k] = fixed[k] ? -1 : nA++;
    if (nA == 0) return ORBS_OK;
    P.K = K; P.E = E; P.nA = nA; P.fix_scale = fix_scale ? 1 : 0;
    static_assert(sizeof(Sim3) == 8 * sizeof(double), "Sim3 must be 8 packed doubles (r xyzw, t, s)");
    int rc = ORBS_OK;
    if ((rc = G.schedule(K, hidx, E, e_i, e_j))) return rc;
    const int nt = P.nt, ns = P.ns, ld = nt * TS;
    ORBS_REQUIRE((size_t)ns * TS2 * sizeof(double) * 3 <= ((size_t)64 << 30), ORBS_E_INVALID, "pose graph too large for one GPU");
    G.eblocks = (E + 127) / 128;
    P.verts = G.alloc<Sim3>(K, &rc); P.backup = G.alloc<Sim3>(K, &rc);
    int *d_hidx = G.alloc<int>(K, &rc), *d_ei = G.alloc<int>(E, &rc), *d_ej = G.alloc<int>(E, &rc);
    Sim3 *d_meas = G.alloc<Sim3>(E, &rc);
    P.err = G.alloc<double>((size_t)7 * E, &rc);
    P.eJ = G.alloc<double>((size_t)98 * E, &rc);
    // one buffer: H | A | L | Linv | b | y | x | partial | scalars, then the dataflow flags
    const size_t nT = (size_t)ns * TS2;
    double *big = G.alloc<double>(3 * nT + (size_t)nt * TS2 + 3 * (size_t)ld + G.eblocks + 12 + (size_t)G.plan.n_part * TS2, &rc);
    const size_t n_flg = (size_t)ns + 2 * (size_t)nt + (size_t)G.plan.n_part + 8;
    int *flg = G.alloc<int>(n_flg, &rc);
    if (rc) return rc;
    P.H = big; P.A = big + nT; G.rsbuf.L = P.A + nT; G.rsbuf.Linv = G.rsbuf.L + nT; P.b = G.rsbuf.Linv + (size_t)nt * TS2;
    G.rsbuf.y = P.b + ld; P.x = G.rsbuf.y + ld; P.partial = P.x + ld; P.scalars = P.partial + G.eblocks; G.rsbuf.part = big + ((3 * nT + (size_t)nt * TS2 + 3 * (size_t)ld + G.eblocks + 8 + 1) & ~(size_t)1);   // 16-byte aligned (double2 accesses)
    G.rsbuf.A = P.A; G.rsbuf.b = P.b; G.rsbuf.x = P.x;
    G.rsbuf.counters = flg; G.rsbuf.flags = flg + 4; G.rsbuf.done_slot = flg + 8; G.rsbuf.done_y = G.rsbuf.done_slot + ns; G.rsbuf.done_x = G.rsbuf.done_y + nt; G.rsbuf.done_part = G.rsbuf.done_x + nt;
    P.hidx = d_hidx; P.e_i = d_ei; P.e_j = d_ej; P.meas = d_meas;
    cudaStream_t st = h->stream;
    ORBS_CUDA(cudaMemsetAsync(flg, 0, n_flg * sizeof(int), st));
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
        ORBS_CUDA(cudaMemsetAsync(P.H, 0, nT * sizeof(double), st));
        ORBS_CUDA(cudaMemsetAsync(P.b, 0, (size_t)ld * sizeof(double), st));
        k_pg_jac<<<(E * 14 + 127) / 128, 128, 0, st>>>(P);
        k_pg_accum<<<(P.n_blocks + 7) / 8, 256, 0, st>>>(P);
        h->launches += 2;
        ORBS_CUDA(cudaMemcpyAsync(hb.data(), P.b, sizeof(double) * ld, cudaMemcpyDeviceToHost, st));
        ORBS_CUDA(cudaStreamSynchronize(st));
        if (it == 0) {
            if (lambda_init > 0) lambda = lambda_init;                                                   // setUserLambdaInit
            else {                                                                                       // computeLambdaInit: 1e-5 * max diagonal
                std::vector<double> diag(ld);
                for (int t = 0; t < nt; t++)
                    ORBS_CUDA(cudaMemcpy2DAsync(diag.data() + (size_t)t * TS, sizeof(double), P.H + (size_t)t * TS2, (TS + 1) * sizeof(double), sizeof(double), TS, cudaMemcpyDeviceToHost, st));
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
            ORBS_CUDA(cudaMemcpyAsync(hx.data(), P.x, sizeof(double) * ld, cudaMemcpyDeviceToHost, st));
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
