// standalone check of k_diffuse_tma against a CPU stencil
#include "../../antsrl_b200/csrc/ants_kernels.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using ants::Params;
int main() {
    const int E = 2, P = 2, W = 50, H = 44, Hp = 48, Wp = 64;
    Params p; memset(&p, 0, sizeof p);
    p.E = E; p.P = P; p.W = W; p.H = H; p.Hp = Hp; p.Wp = Wp; p.plane = (int64_t)Wp * Hp; p.N = 4;
    p.filt_center = (1 - 8 * 0.02) * 0.99; p.filt_ring = 0.02 * 0.99; p.phero_max_val = 255; p.has_max_val = 1;
    size_t cells = (size_t)E * p.plane;
    std::vector<double> h(2 * cells * P, 0.0);
    for (int ep = 0; ep < E * P; ++ep) for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y) {
        double v = ((x * 31 + y * 17 + ep * 7) % 23 == 0) ? 100.0 + x : 0.0;
        if ((x + 2 * y) % 11 == 0) v = -v - 0.0;     // wall cells: negative sign
        h[(size_t)ep * p.plane + (size_t)x * Hp + y] = v;
    }
    double *d; cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    p.phero_pl = d; p.phero_alt = d + cells * P;
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)(2 * E * P)};
    cuuint64_t strides[2] = {(cuuint64_t)Hp * 8, (cuuint64_t)p.plane * 8};
    cuuint32_t box[3] = {ants::kStY + 2 + ants::kStPadY, ants::kStX + 2, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult cr = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)cr);
    int nbx = (W + 31) / 32, nby = (H + 31) / 32;
    ants::k_diffuse_tma<<<nbx * nby * E * P, 256>>>(p, map, 0, nbx, nby);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<double> out(cells * P);
    cudaMemcpy(out.data(), p.phero_alt, out.size() * 8, cudaMemcpyDeviceToHost);
    double maxerr = 0; long bad = 0;
    for (int ep = 0; ep < E * P; ++ep) for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y) {
        double acc = 0;
        for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) {
            int gx = x + dx, gy = y + dy; double v = 0;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = fmax(h[(size_t)ep * p.plane + (size_t)gx * Hp + gy], 0.0);
            acc += v * ((dx == 0 && dy == 0) ? p.filt_center : p.filt_ring);
        }
        acc = acc < 0.01 ? 0 : acc; acc = fmin(acc, 255.0);
        bool wall = std::signbit(h[(size_t)ep * p.plane + (size_t)x * Hp + y]);
        double ref = wall ? -acc : acc, got = out[(size_t)ep * p.plane + (size_t)x * Hp + y];
        double err = fabs(ref - got); if (err > maxerr) maxerr = err; if (err > 1e-9 * (1 + fabs(ref)) || std::signbit(ref) != std::signbit(got)) ++bad;
    }
    printf("max err %g, bad %ld\n", maxerr, bad);
    return 0;
}
