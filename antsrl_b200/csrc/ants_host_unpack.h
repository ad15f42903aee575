// ants_host_unpack.h -- host-side expander of packed observations (see ants_pack.cuh for the format).
#pragma once
#include <stdint.h>

struct AntsUnpackPlan {
    int32_t V, S2, C;
    int32_t vis[228];              // sample index of visible sample v
    int32_t flag_ch[8], val_ch[2], food_ch;
    // derived by ants_unpack_plan_finish()
    float templ[228 * 16];         // one ant's dense observation with -1 in the masked samples (RL_api.py:147-148), 0 elsewhere
    float lut[256][8];             // flags byte -> the 0/1 channels at their channel positions (C <= 8)
    int32_t perm[8];               // lane c of an output sample <- lane {0: value_a, 1: value_b, 2: food}
    uint32_t valmask[8];           // lanes that take the permuted values instead of the lut
    int32_t n_fix, fix[228 * 8];   // floats of masked samples a vector store of the preceding sample spills into (C < 8)
    int32_t simd;                  // 1 = the AVX2 path serves this layout on this CPU
};

void ants_unpack_plan_finish(AntsUnpackPlan *plan);
// expands n_ants consecutive ants: packed -> out (out = the dense observation of the first of them)
void ants_unpack_range(const AntsUnpackPlan *plan, const uint8_t *packed, int64_t n_ants, float *out);
