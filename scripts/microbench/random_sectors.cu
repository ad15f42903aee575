// Random-access read throughput of HBM on B200: every thread issues UNROLL independent loads of GRAN bytes at
// pseudo-random GRAN-aligned offsets of a buffer much larger than L2.  Prints useful GB/s per granularity.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int GRAN>
__global__ void k_random(const uint4 *__restrict__ buf, uint64_t n_gran, int iters, unsigned long long *sink) {
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    constexpr int U = 8;
    for (int it = 0; it < iters; ++it) {
        uint4 v[U][GRAN / 16];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            uint64_t g = (s >> 20) % n_gran;
            const uint4 *p = buf + g * (GRAN / 16);
#pragma unroll
            for (int k = 0; k < GRAN / 16; ++k) v[u][k] = p[k];
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < GRAN / 16; ++k) acc += v[u][k].x ^ v[u][k].w;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int GRAN>
void run(const uint4 *buf, size_t bytes, unsigned long long *sink) {
    const int blocks = 148 * 16, threads = 256, iters = 64;
    uint64_t n_gran = bytes / GRAN;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_random<GRAN><<<blocks, threads>>>(buf, n_gran, 4, sink);
    cudaEventRecord(a);
    k_random<GRAN><<<blocks, threads>>>(buf, n_gran, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double useful = (double)blocks * threads * iters * 8 * GRAN;
    printf("random %3d-byte reads: %8.1f GB/s useful  (%.3f ms, %.2f G accesses/s)\n", GRAN, useful / ms / 1e6, ms,
           (double)blocks * threads * iters * 8 / ms / 1e6);
}

int main() {
    size_t bytes = 16ull << 30;
    uint4 *buf;
    unsigned long long *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    cudaMemset(sink, 0, 8);
    run<16>(buf, bytes, sink);
    run<32>(buf, bytes, sink);
    run<64>(buf, bytes, sink);
    run<128>(buf, bytes, sink);
    run<256>(buf, bytes, sink);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
