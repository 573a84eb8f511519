// 1-D bulk copy (cp.async.bulk, UBLKCP) check (scratch)
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstdint>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const uint8_t *img, int pitch, int x0, int y0, int rows, int rowbytes, uint8_t *out, int *err)
{
    __shared__ __align__(128) uint8_t raw[66 * 80];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(rows * rowbytes) : "memory");
        __syncwarp();
        for (int r = tid; r < rows; r += 32)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(raw + r * rowbytes)), "l"(img + (size_t)(y0 + r) * pitch + x0), "r"(rowbytes), "r"(smem_u32(&mbar)) : "memory");
    }
    unsigned done = 0;
    for (int spin = 0; !done && spin < (1 << 16); spin++)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    if (!__syncthreads_and((int)done)) { if (tid == 0) *err = 3; return; }
    for (int i = tid; i < rows * rowbytes; i += blockDim.x) out[i] = raw[i];
}
int main()
{
    const int h = 376, pitch = 1280, rows = 38, rowbytes = 64, x0 = 1200, y0 = 300;
    std::vector<uint8_t> img((size_t)pitch * h);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t *d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    uint8_t *o; int *e; cudaMalloc(&o, rows * rowbytes); cudaMalloc(&e, 4); cudaMemset(e, 0, 4);
    k<<<1, 128>>>(d, pitch, x0, y0, rows, rowbytes, o, e);
    cudaError_t ce = cudaDeviceSynchronize();
    std::vector<uint8_t> ho(rows * rowbytes); int he = 0;
    cudaMemcpy(ho.data(), o, ho.size(), cudaMemcpyDeviceToHost); cudaMemcpy(&he, e, 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < rows; r++) for (int x = 0; x < rowbytes; x++) bad += ho[r * rowbytes + x] != img[(size_t)(y0 + r) * pitch + x0 + x];
    printf("bulk 1d: cuda=%s err=%d mismatches=%d\n", cudaGetErrorString(ce), he, bad);
    return 0;
}
