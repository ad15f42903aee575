// which tiled-TMA configurations load a (BX x BY) f64 box on this GPU?  one configuration per process (argv[1])
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, uint32_t bytes, double *out, int n) {
    extern __shared__ __align__(128) double tile[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(tile_s), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar_s) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}" : "=r"(done) : "r"(bar_s) : "memory");
    for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = tile[t];
}
int main(int argc, char **argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 50, H = 44, Hp = 48, Wp = 64, Z = 4;
    std::vector<double> h((size_t)Z * Wp * Hp);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *out; cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 34 * 34 * 8 * 2);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)Z};
    cuuint64_t strides[2] = {(cuuint64_t)Hp * 8, (cuuint64_t)Wp * Hp * 8};
    cuuint32_t box[3] = {34, 34, 1}; cuuint32_t es[3] = {1, 1, 1};
    int c0 = -1, c1 = -1, c2 = 1;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (mode == 1) { box[0] = 32; box[1] = 32; c0 = 0; c1 = 0; }
    if (mode == 2) { c0 = 2; c1 = 3; }
    if (mode == 3) { dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; dims[0] = 2 * H; box[0] = 68; c0 = -2; }
    if (mode == 4) { l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE; }
    if (mode == 5) { dt = CU_TENSOR_MAP_DATA_TYPE_UINT64; }
    if (mode == 6) { box[0] = 32; box[1] = 34; c0 = 0; }
    uint32_t bytes = box[0] * box[1] * (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 8);
    CUresult cr = ((EncodeFn)fn)(&map, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d encode %d bytes %u ", mode, (int)cr, bytes);
    k<<<1, 128, 34 * 34 * 8 + 1024>>>(map, c0, c1, c2, bytes, out, (int)(bytes / 8));
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<double> o(bytes / 8); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        printf("   first row: %g %g %g %g ... second row start %g\n", o[0], o[1], o[2], o[3], o[box[0] * (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 0.5 : 1)]);
    }
    return 0;
}
