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
    // tables of the AVX-512 path
    int vidx[228];
    for (int k = 0; k < 228; ++k) vidx[k] = -1;
    for (int v = 0; v < p->V; ++v) vidx[p->vis[v]] = v;
    bool table_ok = n <= 116 * 16;
    p->n_vec = (n + 15) / 16;
    for (int j = 0; table_ok && j < p->n_vec; ++j) {
        int vf = -1;
        for (int l = 0; l < 16; ++l) {
            const int f = 16 * j + l;
            if (f < n && vidx[f / p->C] >= 0 && (vf < 0 || vidx[f / p->C] < vf)) vf = vidx[f / p->C];
        }
        p->vec_src[j] = vf < 0 ? -1 : 12 * vf;
        p->vec_val[j] = p->vec_food[j] = p->vec_store[j] = 0;
        for (int l = 0; l < 16; ++l) {
            const int f = 16 * j + l;
            p->vec_idx[j][l] = 0; p->vec_bit[j][l] = 0u; p->vec_base[j][l] = 0.f;
            if (f >= n) continue;
            p->vec_store[j] |= (uint16_t)(1u << l);
            const int smp = f / p->C, c = f - smp * p->C, v = vidx[smp];
            if (v < 0) { p->vec_base[j][l] = -1.f; continue; }
            const int rel = 3 * (v - vf);
            if (c == p->val_ch[0]) { p->vec_idx[j][l] = rel; p->vec_val[j] |= (uint16_t)(1u << l); }
            else if (c == p->val_ch[1]) { p->vec_idx[j][l] = rel + 1; p->vec_val[j] |= (uint16_t)(1u << l); }
            else if (c == p->food_ch) { p->vec_idx[j][l] = rel + 2; p->vec_food[j] |= (uint16_t)(1u << l); }
            else {
                p->vec_idx[j][l] = rel + 2;
                for (int k = 0; k < 8; ++k)
                    if (p->flag_ch[k] == c) p->vec_bit[j][l] = 1u << (16 + k);
            }
            if (p->vec_idx[j][l] > 15) table_ok = false;       // (fewer than 4 channels: more than 5 samples per vector)
        }
    }
    p->simd = 0;
#if defined(__x86_64__)
    if (p->C <= 8 && __builtin_cpu_supports("avx2")) p->simd = 1;
    if (p->simd && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && !getenv("ANTS_NO_AVX512")) {
        p->simd = 2;
        if (table_ok && !getenv("ANTS_NO_AVX512_TABLE")) p->simd = 3;
    }
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

#if defined(__x86_64__)
// AVX-512, one output vector at a time: 16 consecutive floats of an ant's dense observation come from one 64-byte window
// of its packed record through one permute; the u16 counts are converted and the flag bits tested on the same register
__attribute__((target("avx512f,avx512vl,avx2"))) static void unpack_avx512_table(const AntsUnpackPlan *p, const uint8_t *packed,
                                                                              int64_t n_ants, float *out) {
    const int n = p->S2 * p->C, NV = p->n_vec;
    const int64_t bpa = (int64_t)p->V * 12;
    const __m512i lo16 = _mm512_set1_epi32(0xFFFF);
    const __m512 one = _mm512_set1_ps(1.f);
    // The ants' observations form one flat float stream at `out`.  They are assembled in a cache-resident staging buffer
    // whose 64-byte phase equals the destination's, and every line of the stream leaves with a non-temporal store as soon
    // as it is complete (ant by ant: the write-combining buffers drain while the next ant is assembled), so no line of
    // the destination is read and only the first and last partial line of the range see ordinary stores.
    constexpr int EPOCH = 8;                                   // ants per pass over the staging buffer
    alignas(64) float buf[16 + EPOCH * 228 * 8 + 32];
    const int64_t total = n_ants * n;
    int64_t written = 0, streamed = 0;                         // flat floats assembled / delivered
    int64_t epoch0 = 0;                                        // flat index held at stage[0]
    float *stage = buf + (((uintptr_t)out & 63) >> 2);         // stage[k - epoch0] <-> out[k], same alignment
    bool head_done = (((uintptr_t)out & 63) == 0);
    for (int64_t a = 0; a < n_ants; ++a) {
        if (a > 0 && a % EPOCH == 0) {                         // new pass: carry the incomplete line to the front
            const int64_t left = written - streamed;           // < 16 floats, and out + streamed is line aligned
            memmove(buf, stage + (streamed - epoch0), (size_t)left * sizeof(float));
            stage = buf; epoch0 = streamed;
        }
        float *o = stage + (written - epoch0);
        const uint8_t *rec = packed + a * bpa;
        for (int j = 0; j < NV; ++j) {
            __m512 r = _mm512_loadu_ps(p->vec_base[j]);
            if (p->vec_src[j] >= 0) {
                const __m512i win = _mm512_loadu_si512((const void *)(rec + p->vec_src[j]));
                const __m512i g = _mm512_permutexvar_epi32(_mm512_loadu_si512((const void *)p->vec_idx[j]), win);
                r = _mm512_mask_mov_ps(r, (__mmask16)p->vec_val[j], _mm512_castsi512_ps(g));
                r = _mm512_mask_mov_ps(r, (__mmask16)p->vec_food[j], _mm512_cvtepi32_ps(_mm512_and_si512(g, lo16)));
                r = _mm512_mask_mov_ps(r, _mm512_test_epi32_mask(g, _mm512_loadu_si512((const void *)p->vec_bit[j])), one);
            }
            _mm512_mask_storeu_ps(o + 16 * j, (__mmask16)p->vec_store[j], r);
        }
        written += n;
        if (!head_done) {                                      // the first, partial line of the range
            const int64_t head = 16 - (((uintptr_t)out & 63) >> 2);
            if (written < head) continue;
            memcpy(out, stage, (size_t)head * sizeof(float));
            streamed = head; head_done = true;
        }
        for (; streamed + 16 <= written; streamed += 16)
            _mm512_stream_si512((__m512i *)(out + streamed), _mm512_load_si512((const void *)(stage + (streamed - epoch0))));
    }
    if (!head_done) streamed = 0;                              // (a range shorter than its first partial line)
    memcpy(out + streamed, stage + (streamed - epoch0), (size_t)(total - streamed) * sizeof(float));
    _mm_sfence();
}
#endif

void ants_unpack_range(const AntsUnpackPlan *p, const uint8_t *packed, int64_t n_ants, float *out) {
#if defined(__x86_64__)
    if (p->simd == 3 && n_ants > 6) {
        // (a window reads 64 bytes from some sample of the record: the last ants of the range go through the scalar path
        //  so that nothing past the caller's buffer is read)
        const int64_t tail = (64 + (int64_t)p->V * 12 - 1) / ((int64_t)p->V * 12) + 1;
        unpack_avx512_table(p, packed, n_ants - tail, out);
        unpack_scalar(p, packed + (n_ants - tail) * (int64_t)p->V * 12, tail, out + (n_ants - tail) * (int64_t)p->S2 * p->C);
        return;
    }
    if (p->simd >= 2 && n_ants > 1) {
        unpack_avx512(p, packed, n_ants - 1, out);
        unpack_scalar(p, packed + (n_ants - 1) * (int64_t)p->V * 12, 1, out + (n_ants - 1) * (int64_t)p->S2 * p->C);
        return;
    }
    if (p->simd >= 1 && n_ants > 1) {
        // (the vector path loads 16 bytes per 12-byte sample: the last ant goes through the scalar path so that nothing
        //  past the caller's buffer is read)
        unpack_avx2(p, packed, n_ants - 1, out);
        unpack_scalar(p, packed + (n_ants - 1) * (int64_t)p->V * 12, 1, out + (n_ants - 1) * (int64_t)p->S2 * p->C);
        return;
    }
#endif
    unpack_scalar(p, packed, n_ants, out);
}
