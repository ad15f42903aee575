// ants_pack.cuh -- the observation in the form that crosses PCIe on the host-buffer path (ants_step_host /
// ants_step_host_packed): visible samples only (37 of 49 with the generator's mask, environment_generator.py:35-41),
// 12 bytes per sample instead of 4 C:
//     [f32 value_a][f32 value_b][u16 food][u8 flags][u8 0]
// value_a / value_b = the (up to two) pheromone channels as the f32 the dense observation holds (RL_api.py:124-125),
// food = the food channel as an integer count (RL_api.py:126-127; what the reference's maps hold), flags bit k = the k-th
// 0/1 channel (ants, anthill, walls, rocks: RL_api.py:128-142).  The host expander (ants_host_unpack.cpp) rebuilds the
// dense (N, S, S, C) f32 array bit for bit.  A food amount that is not an integer in [0, 65535) raises `fail`: that
// step is then copied dense.
#pragma once

namespace ants {

constexpr int kPackSampleBytes = 12;

struct PackArgs {
    int32_t V, S2, C;
    int32_t flag_ch[8];            // channel behind bit k of the flags byte, -1 = unused
    int32_t val_ch[2];             // channels stored as f32, -1 = unused
    int32_t food_ch;               // channel stored as u16, -1 = none
    uint8_t vis[228];              // sample index (row-major) of visible sample v
};

__global__ void __launch_bounds__(256)
k_pack_obs(const __grid_constant__ PackArgs a, const float *__restrict__ obs, uint8_t *__restrict__ packed, int64_t ant0,
           int64_t n_ants, uint32_t *__restrict__ fail) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_ants * a.V) return;
    const int64_t la = idx / a.V;
    const int v = (int)(idx - la * a.V);
    const int64_t ant = ant0 + la;
    const float *o = obs + (ant * a.S2 + a.vis[v]) * a.C;
    const float va = a.val_ch[0] >= 0 ? o[a.val_ch[0]] : 0.f;
    const float vb = a.val_ch[1] >= 0 ? o[a.val_ch[1]] : 0.f;
    uint32_t food = 0u, flags = 0u;
    if (a.food_ch >= 0) {
        const float f = o[a.food_ch];
        if (f >= 0.f && f < 65535.f && f == floorf(f)) food = (uint32_t)f;
        else atomicOr(fail, 1u);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (a.flag_ch[k] >= 0) {
            const float f = o[a.flag_ch[k]];
            if (f == 1.f) flags |= 1u << k;
            else if (f != 0.f) atomicOr(fail, 1u);
        }
    uint32_t *dst = reinterpret_cast<uint32_t *>(packed + (ant * a.V + v) * kPackSampleBytes);
    dst[0] = __float_as_uint(va);
    dst[1] = __float_as_uint(vb);
    dst[2] = food | (flags << 16);
}

}  // namespace ants
