// Random 32-byte reads: dependence on (a) total footprint, (b) per-block window (an "environment" of `win` bytes that
// the block stays inside), to separate TLB reach / page locality from DRAM random-access limits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_random(const uint4 *__restrict__ buf, uint64_t n_sec_total, uint64_t n_sec_win, int iters,
                         unsigned long long *sink) {
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint64_t n_win = n_sec_total / n_sec_win;
    uint64_t win = ((blockIdx.x * 0x9E3779B97F4A7C15ull) >> 24) % n_win;      // this block's window
    const uint4 *base = buf + win * n_sec_win * 2;
    uint32_t mask = (uint32_t)(n_sec_win - 1);                                // power of two
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint4 v[8][2];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint4 *p = base + (uint64_t)((uint32_t)(s >> 24) & mask) * 2;
            v[u][0] = p[0]; v[u][1] = p[1];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u][0].x ^ v[u][1].w;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

int main() {
    size_t cap = 32ull << 30;
    uint4 *buf; unsigned long long *sink;
    if (cudaMalloc(&buf, cap) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, cap);
    const int blocks = 148 * 16, threads = 256, iters = 64;
    size_t totals[] = {64ull << 20, 512ull << 20, 4ull << 30, 16ull << 30, 32ull << 30};
    size_t wins[] = {1ull << 20, 32ull << 20, 0};
    for (size_t total : totals)
        for (size_t w : wins) {
            size_t win = w ? w : total;
            if (win > total) continue;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            k_random<<<blocks, threads>>>(buf, total / 32, win / 32, 4, sink);
            cudaEventRecord(a);
            k_random<<<blocks, threads>>>(buf, total / 32, win / 32, iters, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double n = (double)blocks * threads * iters * 8;
            printf("footprint %6zu MB, per-block window %6zu MB: %7.1f GB/s (%.1f G sectors/s)\n", total >> 20, win >> 20,
                   n * 32 / ms / 1e6, n / ms / 1e6);
        }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
