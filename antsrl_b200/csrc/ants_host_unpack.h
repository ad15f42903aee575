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
    int32_t simd;                  // 1 = AVX2 per-sample path, 2 = AVX-512 per-sample path, 3 = AVX-512 table path
    // AVX-512 table path: output vector j = floats [16 j, 16 j + 16) of an ant's dense observation.  All its lanes come
    // from at most 16 consecutive dwords of the packed record (3 dwords per visible sample), starting at vec_src[j]
    int32_t n_vec;
    int32_t vec_src[116];          // byte offset of the 64-byte window in the ant's packed record, -1 = no visible lane
    int32_t vec_idx[116][16];      // lane <- dword of the window
    uint32_t vec_bit[116][16];     // flag lanes: the bit of the sample's third dword; 0 elsewhere
    float vec_base[116][16];       // -1 in the lanes of masked samples, 0 elsewhere
    uint16_t vec_val[116], vec_food[116], vec_store[116];   // lanes taking the raw f32 / the u16 count; lanes that exist
};

void ants_unpack_plan_finish(AntsUnpackPlan *plan);
// expands n_ants consecutive ants: packed -> out (out = the dense observation of the first of them)
void ants_unpack_range(const AntsUnpackPlan *plan, const uint8_t *packed, int64_t n_ants, float *out);
