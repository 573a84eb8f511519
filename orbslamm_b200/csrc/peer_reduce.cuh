// peer_reduce.cuh -- the exchange step of the sharded bundle adjustment over NVLink peer memory.
//
// The sharded solve (optimizer.cu) needs, once per LM trial, the SUM over the ranks of the packed reduced system (structurally nonzero 64x64 fp64 tiles +
// right-hand side: 4.9 MB at 500 keyframes) and of five scalars.  NCCL does that in ~100 us + ~40 us at 8 GPUs, which is as long as everything else in
// the trial but the factorisation.  Here every rank maps every other rank's buffers (cudaIpc handles, exchanged once through the NCCL communicator) and
// the reduction is one or two small kernels on the BA stream:
//   up to 4 ranks:  k_peer_all_reduce      every rank reads every rank's whole partial buffer and writes the sum, in rank order (bit-identical on every rank),
//                                          into a private buffer the solver reads: ONE flag synchronisation
//   more ranks:     k_peer_reduce_scatter  rank r sums chunk r of all ranks' buffers, in rank order, into its own chunk
//                   k_peer_all_gather      every rank copies the other ranks' reduced chunks (into the private buffer)
//   Nobody may overwrite its shared buffer while a peer still reads it: the scalar exchange that follows the solve in the same trial (k_peer_scalars: every
//   rank signals after its own reduction kernels, in stream order, and waits for everybody) is that barrier, so no separate one is needed.
// Synchronisation: per-phase epoch flags in peer memory (st.release.sys after __threadfence_system, ld.acquire.sys spins with a bound: a rank that never
// arrives sets an error flag instead of hanging the device).  The epoch only grows, flags are never reset.
#pragma once
#include "common.cuh"
#include "tile_solver.cuh"

namespace orbs {

constexpr int kMaxPeers = 8;
constexpr unsigned kPeerSpinLimit = 1u << 27;       // ~ a second

struct PeerCtlDev {                                 // one per rank, device memory, mapped by every peer
    unsigned flags[4][kMaxPeers];                   // [phase][source rank] = epoch of the source's last signal
    double scal[2][kMaxPeers][8];                   // [epoch parity][source rank][k]: the five LM scalars of a trial
    int error;
};

struct PeerDev {
    int n, rank;
    PeerCtlDev *ctl[kMaxPeers];                     // ctl[rank] = own block
    double *sys[kMaxPeers];                         // sys[rank] = own packed system buffer
};

__device__ __forceinline__ void peer_st_release(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned peer_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 peer_ld2(const double *p)
{
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

// one warp: lane r < n tells rank r that this rank reached `phase` of exchange `epoch`
__device__ __forceinline__ void peer_signal(const PeerDev &P, int phase, unsigned epoch, int lane)
{
    __threadfence_system();
    if (lane < P.n) peer_st_release(&P.ctl[lane]->flags[phase][P.rank], epoch);
}

// one warp: wait until every rank signalled `phase` of `epoch` (flags only grow: >=, with wrap-safe comparison)
__device__ __forceinline__ void peer_wait(const PeerDev &P, int phase, unsigned epoch, int lane)
{
    if (lane < P.n) {
        const unsigned *f = &P.ctl[P.rank]->flags[phase][lane];
        unsigned spins = 0;
        while ((int)(peer_ld_acquire(f) - epoch) < 0) {
            if (++spins > kPeerSpinLimit) { P.ctl[P.rank]->error = 1; break; }
        }
    }
    __syncwarp();
}

// n2 = number of double2 units in [A | b]; chunk p = [p n2 / n, (p + 1) n2 / n)
__global__ void __launch_bounds__(256)
k_peer_reduce_scatter(const PeerDev P, const LmCtl *__restrict__ ctl, long long n2, unsigned epoch)
{
    if (ctl->state == LM_DONE) return;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        if (blockIdx.x == 0) peer_signal(P, 0, epoch, lane);       // my partial system is complete (stream order: the Schur kernels ran before this one)
        peer_wait(P, 0, epoch, lane);
    }
    __syncthreads();
    const long long c0 = P.rank * n2 / P.n, c1 = (P.rank + 1) * n2 / P.n;
    for (long long i = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c1; i += (long long)gridDim.x * blockDim.x) {
        double2 s = make_double2(0.0, 0.0);
        for (int p = 0; p < P.n; p++) { const double2 v = peer_ld2(P.sys[p] + 2 * i); s.x += v.x; s.y += v.y; }
        reinterpret_cast<double2 *>(P.sys[P.rank])[i] = s;
    }
}

__global__ void __launch_bounds__(256)
k_peer_all_gather(const PeerDev P, const LmCtl *__restrict__ ctl, long long n2, unsigned epoch, double *__restrict__ out)
{
    if (ctl->state == LM_DONE) return;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        if (blockIdx.x == 0) peer_signal(P, 1, epoch, lane);       // my chunk is reduced (the previous kernel finished)
        peer_wait(P, 1, epoch, lane);
    }
    __syncthreads();
    for (int q = 0; q < P.n; q++) {
        const int p = (P.rank + q) % P.n;                          // own chunk first, then the neighbours: the ranks do not all pull from rank 0 at once
        const long long c0 = p * n2 / P.n, c1 = (p + 1) * n2 / P.n;
        for (long long i = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c1; i += (long long)gridDim.x * blockDim.x)
            reinterpret_cast<double2 *>(out)[i] = peer_ld2(P.sys[p] + 2 * i);
    }
}

// one-shot all-reduce for a few ranks: out (private) = sum over the ranks of their shared partial buffers
__global__ void __launch_bounds__(256)
k_peer_all_reduce(const PeerDev P, const LmCtl *__restrict__ ctl, long long n2, unsigned epoch, double *__restrict__ out)
{
    if (ctl->state == LM_DONE) return;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        if (blockIdx.x == 0) peer_signal(P, 0, epoch, lane);
        peer_wait(P, 0, epoch, lane);
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        double2 s = make_double2(0.0, 0.0);
        for (int p = 0; p < P.n; p++) { const double2 v = peer_ld2(P.sys[p] + 2 * i); s.x += v.x; s.y += v.y; }
        reinterpret_cast<double2 *>(out)[i] = s;
    }
}

// sum of `scalars[0..4]` over the ranks, in rank order on every rank: each rank writes its five values into everybody's slot table, then adds up its own table
__global__ void k_peer_scalars(const PeerDev P, const LmCtl *__restrict__ ctl, double *__restrict__ scalars, unsigned epoch)
{
    if (ctl->state == LM_DONE) return;
    const int lane = threadIdx.x, par = epoch & 1;
    if (lane < P.n) {
        double *dst = P.ctl[lane]->scal[par][P.rank];
#pragma unroll
        for (int k = 0; k < 5; k++) dst[k] = scalars[k];
    }
    peer_signal(P, 3, epoch, lane);
    peer_wait(P, 3, epoch, lane);
    if (lane < 5) {
        double s = 0.0;
        for (int p = 0; p < P.n; p++) s += *reinterpret_cast<volatile double *>(&P.ctl[P.rank]->scal[par][p][lane]);
        scalars[lane] = s;
    }
}

}  // namespace orbs
