// vocab.cu -- DBoW2 vocabulary transform on sm_100a (orbv_*): descriptor -> vocabulary word + node at `levelsup`, then the BowVector and
// FeatureVector of every frame.  Replaces S/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1262 (transform), BowVector.cpp:34-84
// (addWeight, normalize), FORB.cpp:81-101 (distance), as called by Frame::ComputeBoW / KeyFrame::ComputeBoW (S/src/Frame.cc:395-402).
// The tree (k = 10, L = 6 for ORBvoc: 1.08 M nodes x 32 B = 35 MB) lives in HBM and is served from the 126 MB L2; one thread walks one
// descriptor (k XOR + POPC distances per level, the first minimum wins like the reference's strict '<').  One CTA per frame then sorts
// (word, feature) and (node, feature) keys in shared memory and builds the two maps in the reference's order; the fp64 sums run in
// feature order and in ascending word order exactly like the std::map loops they replace.
#include <mutex>
#include "common.cuh"

namespace orbs {

struct VocabView {
    int n_nodes, L;
    const uint4 *desc;            // [n_nodes, 2]
    const int *child_start;       // [n_nodes + 1]
    const int *child_ids;
    const int *word_id;           // [n_nodes], leaves only
    const double *weight;         // [n_nodes], leaves only
};

__device__ __forceinline__ int hamming256v(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// transform(feature, word_id, weight, &nid, levelsup), TemplatedVocabulary.h:1217-1262
__global__ void __launch_bounds__(256)
k_vocab_walk(const VocabView V, int slab, int levelsup, const uint4 *__restrict__ desc, const int *__restrict__ counts,
             int *__restrict__ word_of, int *__restrict__ node_of, double *__restrict__ weight_of)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[f]) return;
    const size_t q = (size_t)f * slab + i;
    const uint4 a0 = __ldg(&desc[2 * q]), a1 = __ldg(&desc[2 * q + 1]);
    const int nid_level = V.L - levelsup;
    int nid = 0, final_id = 0, level = 0;
    for (;;) {
        const int c0 = V.child_start[final_id], c1 = V.child_start[final_id + 1];
        if (c0 == c1) break;                                            // isLeaf()
        ++level;
        int best = V.child_ids[c0];
        int best_d = hamming256v(a0, a1, __ldg(&V.desc[2 * best]), __ldg(&V.desc[2 * best + 1]));
        for (int c = c0 + 1; c < c1; c++) {
            const int id = V.child_ids[c];
            const int d = hamming256v(a0, a1, __ldg(&V.desc[2 * id]), __ldg(&V.desc[2 * id + 1]));
            if (d < best_d) { best_d = d; best = id; }
        }
        final_id = best;
        if (level == nid_level) nid = final_id;
    }
    const double w = V.weight[final_id];
    word_of[q] = w > 0 ? V.word_id[final_id] : -1;                      // "if(w > 0) // not stopped"
    node_of[q] = w > 0 ? nid : -1;
    weight_of[q] = w;
}

// in-place bitonic sort of n 64-bit keys in shared memory (padded to a power of two with ~0)
__device__ void block_bitonic_sort(unsigned long long *keys, int n_pow2)
{
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n_pow2; t += blockDim.x) {
                const int p = t ^ j;
                if (p > t) {
                    const unsigned long long a = keys[t], b = keys[p];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { keys[t] = b; keys[p] = a; }
                }
            }
            __syncthreads();
        }
}

// exclusive prefix sum of one int per thread over the block (<= 1024 threads); returns the thread's offset, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp /*[32]*/, int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? s_warp[lane] : 0, x = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += o; }
        s_warp[lane] = x - w;
        if (lane == 31) *total = x;
    }
    __syncthreads();
    const int off = s_warp[wid] + inc - v;
    __syncthreads();
    return off;
}

// BowVector (v.addWeight in feature order, then L1 normalisation in ascending word order) and FeatureVector (fv.addFeature) of one frame.
// After the sort every word / node is a run of keys: the run heads are numbered by a block scan, each head thread sums its run in feature
// order (runs are independent std::map entries), and only the L1 norm -- one fp64 sum over the ascending words -- stays sequential.
__global__ void __launch_bounds__(1024)
k_vocab_assemble(int slab, int n_pow2, const int *__restrict__ counts, const int *__restrict__ word_of, const int *__restrict__ node_of,
                 const double *__restrict__ weight_of, int *__restrict__ bow_ids, double *__restrict__ bow_vals, int *__restrict__ bow_counts,
                 int *__restrict__ fv_nodes, int *__restrict__ fv_start, int *__restrict__ fv_items, int *__restrict__ fv_counts)
{
    extern __shared__ __align__(16) unsigned long long s_keys[];
    __shared__ int s_warp[32], s_total, s_valid;
    __shared__ double s_norm;
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, N = counts[f];
    const size_t o = (size_t)f * slab;
    const int per = (n_pow2 + nt - 1) / nt;                                   // contiguous keys per thread
    for (int pass = 0; pass < 2; pass++) {
        const int *key_of = pass == 0 ? word_of : node_of;
        for (int t = tid; t < n_pow2; t += nt) {
            unsigned long long k = ~0ull;
            if (t < N && key_of[o + t] >= 0) k = ((unsigned long long)(unsigned)key_of[o + t] << 32) | (unsigned)t;
            s_keys[t] = k;
        }
        __syncthreads();
        block_bitonic_sort(s_keys, n_pow2);
        // run heads in this thread's chunk
        const int t0 = tid * per, t1 = min(t0 + per, n_pow2);
        int heads = 0, valid = 0;
        for (int t = t0; t < t1; t++) {
            const unsigned long long k = s_keys[t];
            if (k == ~0ull) break;
            valid++;
            if (t == 0 || (unsigned)(s_keys[t - 1] >> 32) != (unsigned)(k >> 32)) heads++;
        }
        int out = block_exclusive_scan(heads, s_warp, &s_total);
        const int n_out = s_total;
        block_exclusive_scan(valid, s_warp, &s_valid);
        const int n_valid = s_valid;
        int *st = fv_start + (size_t)f * (slab + 1);
        for (int t = t0; t < t1; t++) {
            const unsigned long long k = s_keys[t];
            if (k == ~0ull) break;
            const unsigned id = (unsigned)(k >> 32);
            if (t != 0 && (unsigned)(s_keys[t - 1] >> 32) == id) continue;
            if (pass == 0) {
                double v = weight_of[o + (unsigned)(k & 0xffffffffu)];       // insert, then "vit->second += v" in feature order
                for (int u = t + 1; u < n_pow2 && s_keys[u] != ~0ull && (unsigned)(s_keys[u] >> 32) == id; u++)
                    v = __dadd_rn(v, weight_of[o + (unsigned)(s_keys[u] & 0xffffffffu)]);
                bow_ids[o + out] = (int)id; bow_vals[o + out] = v;
            } else { fv_nodes[o + out] = (int)id; st[out] = t; }
            out++;
        }
        __syncthreads();
        if (pass == 0) {
            double *s_vals = reinterpret_cast<double *>(s_keys);             // the keys of this pass are no longer needed
            for (int a = tid; a < n_out; a += nt) s_vals[a] = bow_vals[o + a];
            __syncthreads();
            if (tid == 0) {                                                  // BowVector::normalize(L1), BowVector.cpp:62-84: one sum in ascending word order
                double norm = 0.0;
                for (int a = 0; a < n_out; a++) norm = __dadd_rn(norm, fabs(s_vals[a]));
                s_norm = norm; bow_counts[f] = n_out;
            }
            __syncthreads();
            const double norm = s_norm;
            if (norm > 0.0) for (int a = tid; a < n_out; a += nt) bow_vals[o + a] = __ddiv_rn(s_vals[a], norm);
        } else {
            if (tid == 0) { st[n_out] = n_valid; fv_counts[f] = n_out; }
            for (int t = tid; t < n_valid; t += nt) fv_items[o + t] = (int)(s_keys[t] & 0xffffffffu);
        }
        __syncthreads();
    }
}

}  // namespace orbs

using namespace orbs;

struct orbv_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    std::mutex mu;
    int k = 0, L = 0, n_nodes = 0;
    DevBuf desc, child_start, child_ids, word_id, weight;
    StagePool pool;
};

constexpr size_t kAssembleSmemMax = 128 * 1024;      // 16384 (word, feature) keys

extern "C" {

int orbv_create(orbv_handle **out, int device, int k, int L, int n_nodes, const uint8_t *node_desc, const int32_t *child_start, const int32_t *child_ids,
                const int32_t *word_id, const double *weight)
{
    ORBS_REQUIRE(out, ORBS_E_INVALID, "orbv_create: null out pointer");
    *out = nullptr;
    ORBS_REQUIRE(node_desc && child_start && child_ids && word_id && weight, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_nodes > 1 && L >= 1 && k >= 1, ORBS_E_INVALID, "empty vocabulary");
    ORBS_REQUIRE(child_start[0] == 0 && child_start[n_nodes] == n_nodes - 1, ORBS_E_INVALID, "child lists must cover every node but the root exactly once");
    ORBS_CUDA(cudaSetDevice(device));
    orbv_handle *h = new orbv_handle();
    h->device = device; h->k = k; h->L = L; h->n_nodes = n_nodes;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    int rc = ORBS_OK;
    if (!rc) rc = h->desc.reserve((size_t)n_nodes * 32);
    if (!rc) rc = h->child_start.reserve((size_t)(n_nodes + 1) * 4);
    if (!rc) rc = h->child_ids.reserve((size_t)n_nodes * 4);
    if (!rc) rc = h->word_id.reserve((size_t)n_nodes * 4);
    if (!rc) rc = h->weight.reserve((size_t)n_nodes * 8);
    if (!rc) {
        e = cudaMemcpyAsync(h->desc.p, node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->child_start.p, child_start, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->child_ids.p, child_ids, (size_t)(n_nodes - 1) * 4, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->word_id.p, word_id, (size_t)n_nodes * 4, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->weight.p, weight, (size_t)n_nodes * 8, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "vocabulary upload", __FILE__, __LINE__);
    }
    if (!rc) {   // per-device function attribute shared by every vocabulary handle: the static bound, never lowered
        e = cudaFuncSetAttribute(k_vocab_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAssembleSmemMax);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaFuncSetAttribute(k_vocab_assemble)", __FILE__, __LINE__);
    }
    if (rc) { orbv_destroy(h); return rc; }
    *out = h;
    return ORBS_OK;
}

int orbv_destroy(orbv_handle *h)
{
    if (!h) return ORBS_OK;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->desc.release(); h->child_start.release(); h->child_ids.release(); h->word_id.release(); h->weight.release(); h->pool.release();
    delete h;
    return ORBS_OK;
}

long long orbv_kernel_launches(const orbv_handle *h) { return h ? h->launches : 0; }

int orbv_transform(orbv_handle *h, int n_frames, const uint8_t *desc, const int32_t *counts, int slab, int levelsup,
                   int32_t *word_of_feature, int32_t *node_of_feature, int32_t *bow_ids, double *bow_vals, int32_t *bow_counts,
                   int32_t *fv_nodes, int32_t *fv_start, int32_t *fv_items, int32_t *fv_counts, int memspace)
{
    ORBS_REQUIRE(h && desc && counts && bow_ids && bow_vals && bow_counts && fv_nodes && fv_start && fv_items && fv_counts, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && slab > 0, ORBS_E_INVALID, "non-positive size");
    ORBS_REQUIRE(slab <= 16384, ORBS_E_INVALID, "at most 16384 features per frame");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    Stager S(&h->pool, h->stream, memspace);
    const size_t n = (size_t)n_frames * slab;
    const uint4 *dd = (const uint4 *)S.in(desc, n * 32);
    const int32_t *dc = S.in(counts, n_frames);
    int32_t *dword = word_of_feature ? S.inout(word_of_feature, n, false) : S.scratch<int32_t>(n);
    int32_t *dnode = node_of_feature ? S.inout(node_of_feature, n, false) : S.scratch<int32_t>(n);
    double *dw = S.scratch<double>(n);
    int32_t *dbi = S.inout(bow_ids, n, false), *dbc = S.inout(bow_counts, n_frames, false);
    double *dbv = S.inout(bow_vals, n, false);
    int32_t *dfn = S.inout(fv_nodes, n, false), *dfs = S.inout(fv_start, (size_t)n_frames * (slab + 1), false), *dfi = S.inout(fv_items, n, false),
            *dfc = S.inout(fv_counts, n_frames, false);
    if (S.rc) return S.rc;
    ORBS_REQUIRE((uintptr_t)dd % 16 == 0, ORBS_E_INVALID, "descriptor array must be 16-byte aligned");
    if (memspace == ORBS_MEM_HOST) {
        ORBS_CUDA(cudaMemsetAsync(dword, 0xff, n * 4, h->stream)); ORBS_CUDA(cudaMemsetAsync(dnode, 0xff, n * 4, h->stream));
        ORBS_CUDA(cudaMemsetAsync(dbi, 0xff, n * 4, h->stream)); ORBS_CUDA(cudaMemsetAsync(dbv, 0, n * 8, h->stream));
        ORBS_CUDA(cudaMemsetAsync(dfn, 0xff, n * 4, h->stream)); ORBS_CUDA(cudaMemsetAsync(dfs, 0, (size_t)n_frames * (slab + 1) * 4, h->stream));
        ORBS_CUDA(cudaMemsetAsync(dfi, 0xff, n * 4, h->stream));
    }
    VocabView V;
    V.n_nodes = h->n_nodes; V.L = h->L; V.desc = h->desc.as<uint4>(); V.child_start = h->child_start.as<int>(); V.child_ids = h->child_ids.as<int>();
    V.word_id = h->word_id.as<int>(); V.weight = h->weight.as<double>();
    k_vocab_walk<<<dim3((slab + 255) / 256, n_frames), 256, 0, h->stream>>>(V, slab, levelsup, dd, dc, dword, dnode, dw);
    int n_pow2 = 1;
    while (n_pow2 < slab) n_pow2 <<= 1;
    const size_t smem = (size_t)n_pow2 * sizeof(unsigned long long);
    ORBS_REQUIRE(smem <= kAssembleSmemMax, ORBS_E_INVALID, "too many features per frame for the on-chip BoW assembly (slab > 16384)");
    k_vocab_assemble<<<n_frames, 1024, smem, h->stream>>>(slab, n_pow2, dc, dword, dnode, dw, dbi, dbv, dbc, dfn, dfs, dfi, dfc);
    h->launches += 2;
    ORBS_CUDA(cudaGetLastError());
    return S.finish();
}

}  // extern "C"
