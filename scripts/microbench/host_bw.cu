// host_bw.cu -- what bounds the end-to-end (host-buffer) path on the GPU box: pinned D2H copy bandwidth, and how fast
// T host threads can WRITE a dense f32 observation array (regular vs non-temporal stores) while reading a packed one.
// nvcc -O3 -Xcompiler -mavx2 -o host_bw host_bw.cu -lpthread
#include <cuda_runtime.h>
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    const size_t MB = 1 << 20;
    const size_t out_bytes = (argc > 1 ? atol(argv[1]) : 728) * MB, in_bytes = out_bytes / 4;
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    char *d = nullptr, *hp = nullptr;
    cudaMalloc(&d, out_bytes);
    cudaMemset(d, 1, out_bytes);
    cudaHostAlloc(&hp, out_bytes, cudaHostAllocDefault);
    memset(hp, 0, out_bytes);
    for (size_t sz : {in_bytes, out_bytes}) {
        cudaMemcpy(hp, d, sz, cudaMemcpyDeviceToHost);
        double t0 = now();
        for (int r = 0; r < 5; ++r) cudaMemcpy(hp, d, sz, cudaMemcpyDeviceToHost);
        double dt = (now() - t0) / 5;
        printf("D2H pinned %zu MB: %.2f ms = %.1f GB/s\n", sz / MB, dt * 1e3, sz / dt / 1e9);
    }
    // chunked D2H on a stream (16 chunks) to see the per-chunk overhead
    {
        cudaStream_t s; cudaStreamCreate(&s);
        double t0 = now();
        for (int r = 0; r < 5; ++r) {
            for (int c = 0; c < 16; ++c) cudaMemcpyAsync(hp + c * (in_bytes / 16), d + c * (in_bytes / 16), in_bytes / 16, cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
        }
        double dt = (now() - t0) / 5;
        printf("D2H pinned %zu MB in 16 chunks: %.2f ms = %.1f GB/s\n", in_bytes / MB, dt * 1e3, in_bytes / dt / 1e9);
    }
    char *src = (char *)aligned_alloc(4096, in_bytes), *dst = hp;   // expansion writes into the pinned output buffer
    memset(src, 3, in_bytes);
    for (int mode = 0; mode < 3; ++mode)
        for (int T : {1, 2, 4, 8, 16, 32, 64}) {
            if ((unsigned)T > 2 * std::thread::hardware_concurrency()) continue;
            auto work = [&](int t) {
                size_t o0 = out_bytes / T * t, o1 = out_bytes / T * (t + 1), i0 = in_bytes / T * t;
                if (mode == 0) { memset(dst + o0, t + 1, o1 - o0); return; }
                const __m256i *in = (const __m256i *)(src + (i0 & ~size_t(31)));
                __m256i *out = (__m256i *)(dst + (o0 & ~size_t(31)));
                size_t n = (o1 - o0) / 128;
                for (size_t k = 0; k < n; ++k) {          // read 32 B, write 128 B (a 4x expansion)
                    __m256i v = _mm256_load_si256(in + k);
                    __m256i a = _mm256_add_epi32(v, v), b = _mm256_add_epi32(a, v), c = _mm256_add_epi32(b, v);
                    if (mode == 1) {
                        _mm256_store_si256(out + 4 * k, v); _mm256_store_si256(out + 4 * k + 1, a);
                        _mm256_store_si256(out + 4 * k + 2, b); _mm256_store_si256(out + 4 * k + 3, c);
                    } else {
                        _mm256_stream_si256(out + 4 * k, v); _mm256_stream_si256(out + 4 * k + 1, a);
                        _mm256_stream_si256(out + 4 * k + 2, b); _mm256_stream_si256(out + 4 * k + 3, c);
                    }
                }
                _mm_sfence();
            };
            double best = 1e9;
            for (int r = 0; r < 3; ++r) {
                double t0 = now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; ++t) th.emplace_back(work, t);
                for (auto &x : th) x.join();
                best = std::min(best, now() - t0);
            }
            printf("%s T=%2d: %.2f ms = %.1f GB/s written\n", mode == 0 ? "memset        " : (mode == 1 ? "expand regular" : "expand stream "),
                   T, best * 1e3, out_bytes / best / 1e9);
        }
    return 0;
}
