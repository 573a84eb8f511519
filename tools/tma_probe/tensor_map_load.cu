// canonical CUDA-guide TMA example (scratch)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <dlfcn.h>
#include <cstdlib>
#include <vector>
#include <cstdint>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 64, BH = 32;
__constant__ CUtensorMap c_map;
__global__ void k(const __grid_constant__ CUtensorMap tensor_map_p, const CUtensorMap *gmap, int mode, int x, int y, uint8_t *out)
{
    const CUtensorMap *tmp = mode == 0 ? &tensor_map_p : mode == 1 ? gmap : &c_map;
    __shared__ alignas(128) uint8_t smem_buffer[BH][BW];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, tmp, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem_buffer[i / BW][i % BW];
}
int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int w = 1241, h = 376, pitch = 1280;
    std::vector<uint8_t> img((size_t)pitch * h);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t *d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSuccess;
    const int how = argc > 2 ? atoi(argv[2]) : 0;          // 0: runtime default, 1: by version 12000, 2: dlsym(libcuda)
    if (how == 0) cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    else if (how == 1) cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    else { cudaFree(0); void *lib = dlopen("libcuda.so.1", RTLD_NOW); p = lib ? dlsym(lib, "cuTensorMapEncodeTiled") : nullptr; }
    printf("entry point how=%d p=%p\n", how, p);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Fn fn = (Fn)p;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}, strides[1] = {(cuuint64_t)pitch}; cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d q=%d\n", (int)r, (int)q);
    uint8_t *o; cudaMalloc(&o, BW * BH);
    const int x = 1209, y = 350;
    CUtensorMap *gm; cudaMalloc(&gm, sizeof tm); cudaMemcpy(gm, &tm, sizeof tm, cudaMemcpyHostToDevice); cudaMemcpyToSymbol(c_map, &tm, sizeof tm);
    k<<<1, 128>>>(tm, gm, mode, x, y, o);
    cudaError_t ce = cudaDeviceSynchronize();
    std::vector<uint8_t> ho(BW * BH);
    cudaMemcpy(ho.data(), o, BW * BH, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < BH; yy++) for (int xx = 0; xx < BW; xx++) {
        const int gx = x + xx, gy = y + yy;
        const uint8_t ref = (gx < w && gy < h) ? img[(size_t)gy * pitch + gx] : 0;
        bad += ho[yy * BW + xx] != ref;
    }
    printf("mode %d (0 param, 1 global, 2 const): cuda=%s mismatches=%d\n", mode, cudaGetErrorString(ce), bad);
    return 0;
}
