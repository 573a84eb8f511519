// common.cuh -- shared host/device helpers for the orbslamm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>
#include "../../include/orbslamm_b200.h"

namespace orbs {

void set_last_error(const std::string &s);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define ORBS_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return orbs::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define ORBS_REQUIRE(cond, code, msg)                                          \
    do {                                                                       \
        if (!(cond)) { orbs::set_last_error(msg); return (code); }             \
    } while (0)

// growable device buffer
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n)
    {
        if (n <= bytes) return ORBS_OK;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
        bytes = want;
        return ORBS_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n)
    {
        if (n <= bytes) return ORBS_OK;
        if (p) { cudaFreeHost(p); p = nullptr; bytes = 0; }
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__);
        bytes = want;
        return ORBS_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Optional per-kernel CUDA-event timing on the launching stream (bench roofline numbers).
struct KernelTimer {
    static constexpr int kMaxKernels = 16;
    bool enabled = false;
    struct Rec { cudaEvent_t a, b; int id; };
    std::vector<Rec> recs;
    size_t used = 0;
    double total_ms[kMaxKernels] = {0};
    long long count[kMaxKernels] = {0};
    void begin(int id, cudaStream_t st)
    {
        if (!enabled) return;
        if (used == recs.size()) { Rec r; cudaEventCreate(&r.a); cudaEventCreate(&r.b); r.id = id; recs.push_back(r); }
        recs[used].id = id;
        cudaEventRecord(recs[used].a, st);
    }
    void end(cudaStream_t st)
    {
        if (!enabled) return;
        cudaEventRecord(recs[used].b, st);
        used++;
    }
    // call after the stream is synchronised
    void collect()
    {
        for (size_t i = 0; i < used; i++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, recs[i].a, recs[i].b) == cudaSuccess) { total_ms[recs[i].id] += ms; count[recs[i].id]++; }
        }
        used = 0;
    }
    void reset() { used = 0; for (int i = 0; i < kMaxKernels; i++) { total_ms[i] = 0; count[i] = 0; } }
    void release() { for (auto &r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } recs.clear(); used = 0; }
};

// host<->device staging for ORBS_MEM_HOST calls: inputs are copied to pooled device buffers, outputs copied back in finish()
struct StagePool {
    static constexpr int kSlots = 96;
    DevBuf buf[kSlots];
    void release() { for (auto &b : buf) b.release(); }
};

struct Stager {
    StagePool *pool; cudaStream_t stream; int memspace; int rc = ORBS_OK; int used = 0;
    struct Out { void *host; void *dev; size_t bytes; };
    std::vector<Out> outs;
    Stager(StagePool *p, cudaStream_t st, int ms) : pool(p), stream(st), memspace(ms) {}
    DevBuf *next(size_t bytes)
    {
        if (used >= StagePool::kSlots) { set_last_error("internal: staging pool exhausted"); rc = ORBS_E_INVALID; return nullptr; }
        DevBuf &b = pool->buf[used++];
        if ((rc = b.reserve(bytes + 16))) return nullptr;
        return &b;
    }
    template <typename T> const T *in(const T *p, size_t n)
    {
        if (memspace == ORBS_MEM_DEVICE || !p || rc) return p;
        DevBuf *b = next(n * sizeof(T));
        if (!b) return nullptr;
        cudaError_t e = cudaMemcpyAsync(b->p, p, n * sizeof(T), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "stage in", __FILE__, __LINE__); return nullptr; }
        return b->as<T>();
    }
    template <typename T> T *inout(T *p, size_t n, bool copy_in = true)
    {
        if (memspace == ORBS_MEM_DEVICE || !p || rc) return p;
        DevBuf *b = next(n * sizeof(T));
        if (!b) return nullptr;
        if (copy_in) {
            cudaError_t e = cudaMemcpyAsync(b->p, p, n * sizeof(T), cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) { rc = cuda_fail(e, "stage inout", __FILE__, __LINE__); return nullptr; }
        }
        outs.push_back({p, b->p, n * sizeof(T)});
        return b->as<T>();
    }
    template <typename T> T *scratch(size_t n)      // device-only scratch in both memory spaces
    {
        if (rc) return nullptr;
        DevBuf *b = next(n * sizeof(T));
        return b ? b->as<T>() : nullptr;
    }
    int finish()
    {
        if (rc) return rc;
        if (memspace == ORBS_MEM_DEVICE) return ORBS_OK;
        for (auto &o : outs) ORBS_CUDA(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, stream));
        ORBS_CUDA(cudaStreamSynchronize(stream));
        return ORBS_OK;
    }
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace orbs
