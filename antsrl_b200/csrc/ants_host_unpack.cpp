// ants_host_unpack.cpp -- rebuilds the dense (N, S, S, C) f32 observation from its packed PCIe form on the host, bit for
// bit what ants_step_host used to copy (ants_pack.cuh).  Runs on the worker threads of ants_abi.cu while the next chunk
// is still crossing PCIe.  The output is written with non-temporal stores (whole cache lines, no read-for-ownership):
// the expansion is bound by host memory bandwidth (scripts/microbench/host_bw.cu: 145 GB/s written with streaming stores
// against 78 GB/s with regular ones on the 16-core B200 host).
#include "ants_host_unpack.h"

#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

void ants_unpack_plan_finish(AntsUnpackPlan *p) {
    const int n = p->S2 * p->C;
    for (int k = 0; k < n; ++k) p->templ[k] = -1.f;
    for (int v = 0; v < p->V; ++v)
        for (int c = 0; c < p->C; ++c) p->templ[p->vis[v] * p->C + c] = 0.f;
    for (int f = 0; f < 256; ++f)
        for (int c = 0; c < 8; ++c) {
            float x = 0.f;
            for (int k = 0; k < 8; ++k)
                if (p->flag_ch[k] == c && ((f >> k) & 1)) x = 1.f;
            p->lut[f][c] = x;
        }
    for (int c = 0; c < 8; ++c) {
        p->perm[c] = 3;            // lane 3 of the loaded sample: never selected
        p->valmask[c] = 0u;
        if (c == p->val_ch[0]) { p->perm[c] = 0; p->valmask[c] = 0xFFFFFFFFu; }
        if (c == p->val_ch[1]) { p->perm[c] = 1; p->valmask[c] = 0xFFFFFFFFu; }
        if (c == p->food_ch) { p->perm[c] = 2; p->valmask[c] = 0xFFFFFFFFu; }
    }
    // an 8-float store of sample s covers floats [s C, s C + 8): those beyond s C + C belong to the following samples.
    // Visible followers are rewritten afterwards (ascending order); masked ones are put back to -1 from this list.
    bool visible[228] = {false};
    for (int v = 0; v < p->V; ++v) visible[p->vis[v]] = true;
    p->n_fix = 0;
    for (int v = 0; v < p->V; ++v)
        for (int k = p->vis[v] * p->C + p->C; k < p->vis[v] * p->C + 8 && k < n; ++k)
            if (!visible[k / p->C]) {
                bool seen = false;
                for (int j = 0; j < p->n_fix; ++j) seen |= p->fix[j] == k;
                if (!seen) p->fix[p->n_fix++] = k;
            }
    // ... and the spill of the previous ant's last sample into this ant's first floats (the staging buffer holds
    // several ants back to back)
    for (int k = 0; k < 8 - p->C && k < n; ++k)
        if (!visible[k / p->C]) {
            bool seen = false;
            for (int j = 0; j < p->n_fix; ++j) seen |= p->fix[j] == k;
            if (!seen) p->fix[p->n_fix++] = k;
        }
    p->simd = 0;
#if defined(__x86_64__)
    if (p->C <= 8 && __builtin_cpu_supports("avx2")) p->simd = 1;
    if (p->simd && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && !getenv("ANTS_NO_AVX512")) p->simd = 2;
#endif
}

static void unpack_scalar(const AntsUnpackPlan *p, const uint8_t *packed, int64_t n_ants, float *out) {
    const int n = p->S2 * p->C;
    for (int64_t a = 0; a < n_ants; ++a) {
        float *o = out + a * n;
        memcpy(o, p->templ, (size_t)n * sizeof(float));
        const uint8_t *rec = packed + a * p->V * 12;
        for (int v = 0; v < p->V; ++v, rec += 12) {
            float va, vb;
            uint16_t food;
            memcpy(&va, rec, 4); memcpy(&vb, rec + 4, 4); memcpy(&food, rec + 8, 2);
            const unsigned flags = rec[10];
            float *s = o + p->vis[v] * p->C;
            if (p->val_ch[0] >= 0) s[p->val_ch[0]] = va;
            if (p->val_ch[1] >= 0) s[p->val_ch[1]] = vb;
            if (p->food_ch >= 0) s[p->food_ch] = (float)food;
            for (int k = 0; k < 8; ++k)
                if (p->flag_ch[k] >= 0) s[p->flag_ch[k]] = ((flags >> k) & 1u) ? 1.f : 0.f;
        }
    }
}

#if defined(__x86_64__)
// streaming copy of `bytes` from a cache-resident buffer: whole 32-byte vectors go out with non-temporal stores
__attribute__((target("avx2"))) static void stream_out(float *dst, const float *src, size_t bytes) {
    uint8_t *d = (uint8_t *)dst;
    const uint8_t *s = (const uint8_t *)src;
    size_t head = ((uintptr_t)d & 31) ? 32 - ((uintptr_t)d & 31) : 0;
    if (head > bytes) head = bytes;
    memcpy(d, s, head);
    d += head; s += head; bytes -= head;
    const size_t nv = bytes / 32;
    for (size_t k = 0; k < nv; ++k)
        _mm256_stream_si256((__m256i *)d + k, _mm256_loadu_si256((const __m256i *)s + k));
    memcpy(d + nv * 32, s + nv * 32, bytes - nv * 32);
}

__attribute__((target("avx2"))) static void unpack_avx2(const AntsUnpackPlan *p, const uint8_t *packed, int64_t n_ants, float *out) {
    const int n = p->S2 * p->C, C = p->C, V = p->V;
    constexpr int B = 8;                                           // ants staged in the cache-resident buffer
    alignas(64) float buf[B * 228 * 8 + 16];
    for (int b = 0; b < B; ++b) memcpy(buf + b * n, p->templ, (size_t)n * sizeof(float));
    const __m256i perm = _mm256_loadu_si256((const __m256i *)p->perm);
    const __m128i foodmask = _mm_set_epi32(0, 0xFFFF, 0, 0);
    for (int64_t a0 = 0; a0 < n_ants; a0 += B) {
        const int nb = (int)(n_ants - a0 < B ? n_ants - a0 : B);
        for (int b = 0; b < nb; ++b) {
            float *o = buf + b * n;
            const uint8_t *rec = packed + (a0 + b) * V * 12;
            for (int v = 0; v < V; ++v, rec += 12) {
                const __m128i x = _mm_loadu_si128((const __m128i *)rec);                 // [a][b][food | flags << 16][next]
                const __m128 foodf = _mm_cvtepi32_ps(_mm_and_si128(x, foodmask));
                const __m128 vals = _mm_blend_ps(_mm_castsi128_ps(x), foodf, 12);        // [a][b][food as f32][0.0]
                // value channels take lanes 0..2, the 0/1 channels lane 3 (= 0.0) and then the lut's bits
                const __m256 spread = _mm256_permutevar8x32_ps(_mm256_castps128_ps256(vals), perm);
                const __m256 lut = _mm256_loadu_ps(p->lut[rec[10]]);
                _mm256_storeu_ps(o + p->vis[v] * C, _mm256_or_ps(lut, spread));
            }
            for (int k = 0; k < p->n_fix; ++k) o[p->fix[k]] = -1.f;
        }
        stream_out(out + a0 * n, buf, (size_t)nb * n * sizeof(float));
    }
    _mm_sfence();
}
#endif

#if defined(__x86_64__)
// AVX-512 (F + VL): a sample is stored through a lane mask (no spill into the next sample, no repair list) and the
// staged ants leave in 64-byte non-temporal stores
__attribute__((target("avx512f,avx512vl,avx2"))) static void stream_out512(float *dst, const float *src, size_t bytes) {
    uint8_t *d = (uint8_t *)dst;
    const uint8_t *s = (const uint8_t *)src;
    size_t head = ((uintptr_t)d & 63) ? 64 - ((uintptr_t)d & 63) : 0;
    if (head > bytes) head = bytes;
    memcpy(d, s, head);
    d += head; s += head; bytes -= head;
    const size_t nv = bytes / 64;
    for (size_t k = 0; k < nv; ++k)
        _mm512_stream_si512((__m512i *)d + k, _mm512_loadu_si512((const __m512i *)s + k));
    memcpy(d + nv * 64, s + nv * 64, bytes - nv * 64);
}

__attribute__((target("avx512f,avx512vl,avx2"))) static void unpack_avx512(const AntsUnpackPlan *p, const uint8_t *packed, int64_t n_ants,
                                                                        float *out) {
    const int n = p->S2 * p->C, C = p->C, V = p->V;
    constexpr int B = 8;
    alignas(64) float buf[B * 228 * 8 + 16];
    int off[228];
    for (int v = 0; v < V; ++v) off[v] = p->vis[v] * C;
    for (int b = 0; b < B; ++b) memcpy(buf + b * n, p->templ, (size_t)n * sizeof(float));
    const __m256i perm = _mm256_loadu_si256((const __m256i *)p->perm);
    const __m128i foodmask = _mm_set_epi32(0, 0xFFFF, 0, 0);
    const __mmask8 lanes = (__mmask8)((1u << C) - 1u);
    for (int64_t a0 = 0; a0 < n_ants; a0 += B) {
        const int nb = (int)(n_ants - a0 < B ? n_ants - a0 : B);
        for (int b = 0; b < nb; ++b) {
            float *o = buf + b * n;
            const uint8_t *rec = packed + (a0 + b) * V * 12;
            for (int v = 0; v < V; ++v, rec += 12) {
                const __m128i x = _mm_loadu_si128((const __m128i *)rec);
                const __m128 foodf = _mm_cvtepi32_ps(_mm_and_si128(x, foodmask));
                const __m128 vals = _mm_blend_ps(_mm_castsi128_ps(x), foodf, 12);
                const __m256 spread = _mm256_permutevar8x32_ps(_mm256_castps128_ps256(vals), perm);
                const __m256 lut = _mm256_loadu_ps(p->lut[rec[10]]);
                _mm256_mask_storeu_ps(o + off[v], lanes, _mm256_or_ps(lut, spread));
            }
        }
        stream_out512(out + a0 * n, buf, (size_t)nb * n * sizeof(float));
    }
    _mm_sfence();
}
#endif

void ants_unpack_range(const AntsUnpackPlan *p, const uint8_t *packed, int64_t n_ants, float *out) {
#if defined(__x86_64__)
    if (p->simd == 2 && n_ants > 1) {
        unpack_avx512(p, packed, n_ants - 1, out);
        unpack_scalar(p, packed + (n_ants - 1) * (int64_t)p->V * 12, 1, out + (n_ants - 1) * (int64_t)p->S2 * p->C);
        return;
    }
    if (p->simd && n_ants > 1) {
        // (the vector path loads 16 bytes per 12-byte sample: the last ant goes through the scalar path so that nothing
        //  past the caller's buffer is read)
        unpack_avx2(p, packed, n_ants - 1, out);
        unpack_scalar(p, packed + (n_ants - 1) * (int64_t)p->V * 12, 1, out + (n_ants - 1) * (int64_t)p->S2 * p->C);
        return;
    }
#endif
    unpack_scalar(p, packed, n_ants, out);
}
