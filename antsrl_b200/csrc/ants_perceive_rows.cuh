// ants_perceive_rows.cuh -- the perception + reward kernel for the generator's default channel list
// (RL_api.py:96-165, reward_custom.py:79-106, ants.py:119-121), "row per lane" mapping.
//
// The generic kernel (k_perceive in ants_kernels.cuh) spreads a chunk's samples over the lanes by a flat sample
// index: every sample re-derives its ant and its window position and re-reads the ant's frame from shared memory,
// and the exploration counts need a ballot split at the ant boundaries -- ncu showed ~490 warp instructions per ant
// with the issue slots as the limiter.  Here a warp still owns 32 consecutive ants and walks them in chunks of 4
// (the bulk-store granule: 4 * S*S * C * 4 bytes is always a multiple of 16), but lane l of the chunk is
//     ant  a = l / S        row  i = l % S        (4 * S <= 32 lanes active; 28 of 32 for the default 7x7 window)
// and loops over the S columns j of its row.  Everything that depends on (ant, row) -- cos/sin of the frame, the
// shifted position, st*Y and ct*Y of RL_api.py:110-111, the anthill, the record base, the row's mask bits -- lives in
// registers for the whole row; the column offset X = off[j] is a kernel-parameter constant after unrolling, so a
// sample costs two f64 multiplies, four adds and two conversions for its cell (bit-identical to the reference's
// ct*X - st*Y, st*X + ct*Y: every product is rounded separately, -fmad=false), one 128-bit record load (compact
// records; two for f64 records), a branch-free decode and C shared-memory stores.  All record loads of a row are in
// flight before the first is decoded.  Exploration counts are per-lane adds, summed over the S rows in the epilogue.
// The staged (4 x S*S x C) f32 tile leaves with one TMA bulk store, as before.
#pragma once

namespace ants {

#ifndef ANTS_ROWS_THREADS
#define ANTS_ROWS_THREADS 32
#endif
constexpr int kRowsThreads = ANTS_ROWS_THREADS;
constexpr int kRowsGroup = 4;
#ifndef ANTS_ROWS_FLAT
#define ANTS_ROWS_FLAT 0           // 1 = hand the 32 * S rows of a warp to the lanes without idle lanes (measured slower: 0.279 vs 0.241 ms)
#endif
#ifndef ANTS_ROWS_TILES
#define ANTS_ROWS_TILES (ANTS_ROWS_FLAT ? 2 : 1)
#endif
constexpr int kRowsTiles = ANTS_ROWS_TILES;   // staging tiles per warp (2 = the bulk store of chunk g drains while g + 1 is staged)

struct RowPrep {                 // 72 bytes per ant, shared memory (phase A -> phase B)
    double ct, st;               // cos / sin(theta + pi/2)
    double xf, yf;               // position shifted forward by perception_fwd_delta
    unsigned long long rocks;    // further candidate rocks (beyond the first) whose disc can reach the window
    int e, flags;                // environment; 2 = a rock may reach the window, 4 = more than one
    double rcx, rcy, rrad;       // the first candidate rock
};

// one conditional add or subtract wraps a sample coordinate onto the torus (np.mod on ints, RL_api.py:118-119) when
// the window reaches less than one map size past the border: of v, v + n, v - n the one in [0, n) is the smallest as
// an unsigned number
__device__ __forceinline__ int wrap1(int v, int n) {
    return (int)min(min((unsigned)v, (unsigned)(v + n)), (unsigned)(v - n));
}
// np.round(v).astype(int) (half to even, RL_api.py:117) for |v| < 2^31 without the quarter-rate F2I.F64: adding
// 1.5 * 2^52 leaves the rounded integer (two's complement) in the low word of the sum -- the add itself rounds to
// nearest-even, exactly like rint()
__device__ __forceinline__ int round_half_even(double v) { return __double2loint(v + 6755399441055744.0); }
__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2 (2^-22 relative); arguments here are in [-15, 0]
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 16-byte record load.  (Measured alternatives, all slower on the cfg4 shard: ld.global.L2::128B / L2::64B prefetch
// sizes and the non-coherent path, 0.35 ms against 0.29 ms.)
__device__ __forceinline__ uint4 ld_record16(const uint8_t *rp) { return *reinterpret_cast<const uint4 *>(rp); }

// value of a pheromone field that is neither zero nor a live boxed deposit outside walls (rare): kept out of line
__device__ __noinline__ float phero_obs_slow(const Params &p, const uint8_t *rp, int k, uint32_t now, uint32_t now_abs) {
    const double inv_max = 1.0 / p.phero_max_val;
    return (float)(phero_value(p, rp, k, now, now_abs) * inv_max);
}

// exact rock test of one sample cell against the candidate rocks (RL_api.py:132-135, strict <)
__device__ __noinline__ float rock_channel(const Params &p, int e, unsigned long long rm, int ix, int iy) {
    const double *rc = p.rock_c + (int64_t)e * p.R * 2;
    const double *rr = p.rock_rad + (int64_t)e * p.R;
    while (rm) {
        const int r = __ffsll((long long)rm) - 1;
        rm &= rm - 1;
        const double ddx = (double)ix - rc[2 * r], ddy = (double)iy - rc[2 * r + 1];
        if (sqrt(ddx * ddx + ddy * ddy) < rr[r]) return 1.f;
    }
    return 0.f;
}

template <int LAYOUT, int REC, int S>      // REC: 0 = f64 records, 1 = compact 16-byte, 2 = compact 8-byte
#ifndef ANTS_ROWS_OCC
#define ANTS_ROWS_OCC ((ANTS_ROWS_FLAT ? 4 : 5) * 128 / ANTS_ROWS_THREADS)   // resident blocks per SM the register budget allows (shared memory: 4 / 5 of 128 threads)
#endif
#ifndef ANTS_ROWS_UNR
#define ANTS_ROWS_UNR 7
#endif
#ifndef ANTS_ROWS_ROCKS_WARP
#define ANTS_ROWS_ROCKS_WARP 1     // 1 = the rock channel of an ant whose window a rock may reach is evaluated by the whole
                                   // warp (its S*S samples over the lanes, ballots back to the row lanes) instead of by the
                                   // ant's S row lanes while the other rows of the chunk idle
#endif
#ifndef ANTS_ROWS_MAGIC
#define ANTS_ROWS_MAGIC 1          // 1 = 8-byte records: (float)age through the 2^23 exponent trick (PRMT + FFMA) instead of I2F + FMUL
#endif
#ifndef ANTS_ROWS_BASE64
#define ANTS_ROWS_BASE64 1         // 1 = the environment's record base as one opaque 64-bit register (one IMAD.WIDE per sample address)
#endif
#ifndef ANTS_ROWS_FULLWARP
#define ANTS_ROWS_FULLWARP 0       // 1 = a warp with all of its 32 ants runs a copy of the chunk loop without the ragged-end bookkeeping
#endif
#ifndef ANTS_ROWS_SYM
#define ANTS_ROWS_SYM 1            // 1 = the window offsets are (j - radius) * DELTA (RL_api.py:92-93), so ct * X and st * X are
                                   // computed for the positive columns only: a product's sign flips exactly with its factor's
#endif
#ifndef ANTS_ROWS_PIPE
#define ANTS_ROWS_PIPE 0           // 1 = the record loads of the NEXT chunk are issued before the current chunk is decoded
                                   // (double-buffered registers; compact records only)
#endif
#ifndef ANTS_ROWS_PREFETCH
#define ANTS_ROWS_PREFETCH 0       // samples of the NEXT chunk's row whose records are prefetched into L2 (0 = off, S = all).
                                   // Measured on the cfg4 shard: 0.240 ms without, 0.254 with 4, 0.269 with 7 -- the
                                   // kernel has no issue slots to spare for the address arithmetic
#endif
__global__ void __launch_bounds__(kRowsThreads, ANTS_ROWS_OCC)
k_perceive_rows(const __grid_constant__ Params p, float *__restrict__ obs, float *__restrict__ agent_state, float *__restrict__ state_out,
                double *__restrict__ reward_out, uint32_t obs_gen, uint32_t occ_gen, int is_step, int rw_alias,
                uint32_t now, uint32_t now_abs, int64_t ant0, int64_t ant_end) {
    // (ants [ant0, ant_end) of the batch: ants_rollout runs groups of environments on streams of their own)
    pdl_begin();
    static_assert(LAYOUT == 1 || LAYOUT == 2, "default channel lists only");
    constexpr bool REC16 = (REC == 1), REC8 = (REC == 2);
    constexpr bool FLAT = (ANTS_ROWS_FLAT != 0) && S == 7 && REC != 0;
    constexpr int SH = REC8 ? 3 : (REC16 ? 4 : 5);                    // log2(record bytes)
    static_assert(kRowsGroup * S <= 32, "a chunk's rows must fit one warp");
    constexpr int S2 = S * S, C = (LAYOUT == 2) ? 7 : 6, SC = S2 * C;
    constexpr int G = kRowsGroup, ROWS = G * S, TILE = G * SC;       // TILE * 4 bytes is a multiple of 16
    constexpr int NW = kRowsThreads / 32;
    constexpr int UNR = (REC != 0) ? (S < ANTS_ROWS_UNR ? S : ANTS_ROWS_UNR) : (S + 1) / 2;                      // record loads in flight per lane
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_obs = reinterpret_cast<float *>(smem_raw);               // [NW][TILE]
    RowPrep *prep = reinterpret_cast<RowPrep *>(s_obs + NW * kRowsTiles * TILE);   // [threads]
    uint8_t *s_rowcnt = reinterpret_cast<uint8_t *>(prep + kRowsThreads);   // [threads][S]

    const int tid = threadIdx.x;
    const int64_t base = ant0 + (int64_t)blockIdx.x * kRowsThreads;
    const bool explore_on = p.explore_on != 0;

    // ---- phase A (thread per ant): frame, reward terms that do not need the exploration count, small outputs.
    // The kernel's pointers may alias as far as the compiler knows, so a load after a store waits for it: every load
    // of the ant's state is issued first, the stores come last, and the dependent chain is state -> rock grid word ->
    // rock discs (ncu, round 2: phase A / C were 12 % of the instructions and 29 % of the warps' stall samples).
    double r_other = 0.0, r_mult = 1.0;
    int my_rock_flags = 0;
    int rs_prev = 0;                                                           // reward_state, used by phase C
    {
        const int64_t i = base + tid;
        if (i < ant_end) {
            const int e = (int)(i / p.N);
            const double x = p.x[i], y = p.y[i], th = p.theta[i], hold = p.holding[i];
            const double hprev = rw_alias ? hold : p.rw_holding_prev[i];       // Q18
            const double pdist = (p.reward_kind == 0) ? p.rw_prev_dist[i] : 0.0;
            const int32_t *hl = p.hill + 4 * e;
            const int32_t hx = hl[0], hy = hl[1];
            const float seed_f = (float)p.seed[i];
            const uint8_t mand = p.mandibles[i];
            const bool want_state = state_out != nullptr;
            const double act0 = (want_state && p.P > 0) ? p.act[i] : 0.0;
            const double act1 = (want_state && p.P > 1) ? p.act[(int64_t)p.EN + i] : 0.0;
            if (is_step) rs_prev = p.reward_state[i];
            double s0, c0, st, ct;
            sincos(th, &s0, &c0);
            RowPrep q;
            q.xf = x; q.yf = y;
            if (p.fwd_delta != 0.0) { q.xf = x + c0 * p.fwd_delta; q.yf = y + s0 * p.fwd_delta; }   // RL_api.py:103-104
            q.e = e;
            // grid candidates (the word is on its way while the frame and the reward terms are computed)
            unsigned long long rm = (LAYOUT == 2) ? rock_candidates(p, e, q.xf, q.yf) : 0ull;
            sincos(th + 3.141592653589793 * 0.5, &st, &ct);                    // RL_api.py:101,107-108
            q.ct = ct; q.st = st;
            const double d = hold - hprev;
            double nd = 0.0;
            if (p.reward_kind == 0) {                                          // All_Rewards, reward_custom.py:79-106
                const double r_food = d < 0.0 ? 0.0 : d;
                const double r_hill = d < 0.0 ? 1.0 : 0.0;
                const double ddx = x - (double)hx, ddy = y - (double)hy;
                nd = sqrt(ddx * ddx + ddy * ddy);
                const double heading = (pdist > nd && hold > 0.0) ? 0.1 : 0.0;
                r_other = r_food * p.f_food + r_hill * p.f_anthill + heading * p.f_heading;
                r_mult = (hold == 0.0) ? p.f_explore : p.f_explore_hold;
            } else if (p.reward_kind == 2) {                                   // Food_Reward, :37-40
                r_other = d < 0.0 ? 10.0 : d;
            }
            q.flags = 0;
            q.rcx = q.rcy = 0.0; q.rrad = -1.0;
            q.rocks = 0ull;
            if (LAYOUT == 2 && rm) {
                // narrowed to the rocks whose disc can reach the window: a sample lies within
                // radius*DELTA*sqrt(2) + 0.5 (rounding) of (xf, yf) on the torus.  The discs of the first two
                // candidates are loaded together (more than two are rare)
                const double *rc = p.rock_c + (int64_t)e * p.R * 2;
                const double *rr = p.rock_rad + (int64_t)e * p.R;
                const double reach = (double)p.radius * p.delta * 1.4142135623730951 + 1.0;
                auto near = [&](double cx, double cy, double rad) -> bool {
                    double dx = fabs(q.xf - cx), dy = fabs(q.yf - cy);
                    dx = fmin(dx, fabs((double)p.W - dx));
                    dy = fmin(dy, fabs((double)p.H - dy));
                    const double L = rad + reach;
                    return dx < L && dy < L && dx * dx + dy * dy < L * L;
                };
                const int ra = __ffsll((long long)rm) - 1;
                unsigned long long it = rm & (rm - 1);
                const int rb = it ? __ffsll((long long)it) - 1 : ra;
                it &= it - 1;
                const double ax = rc[2 * ra], ay = rc[2 * ra + 1], arad = rr[ra];
                const double bx = rc[2 * rb], by = rc[2 * rb + 1], brad = rr[rb];
                const bool ka = near(ax, ay, arad), kb = (rb != ra) && near(bx, by, brad);
                unsigned long long keep = (ka ? 1ull << ra : 0ull) | (kb ? 1ull << rb : 0ull);
                while (it) {
                    const int r = __ffsll((long long)it) - 1;
                    it &= it - 1;
                    if (near(rc[2 * r], rc[2 * r + 1], rr[r])) keep |= 1ull << r;
                }
                if (keep) {
                    const int r = __ffsll((long long)keep) - 1;            // the first candidate rides in the prep record
                    keep &= keep - 1;
                    if (r == ra) { q.rcx = ax; q.rcy = ay; q.rrad = arad; }
                    else if (r == rb) { q.rcx = bx; q.rcy = by; q.rrad = brad; }
                    else { q.rcx = rc[2 * r]; q.rcy = rc[2 * r + 1]; q.rrad = rr[r]; }
                    q.flags = keep ? 6 : 2;
                    q.rocks = keep;
                }
                my_rock_flags = q.flags;
            }
            prep[tid] = q;
            // the stores
            if (p.reward_kind == 0) { p.rw_prev_dist[i] = nd; p.rw_holding_prev[i] = hold; }
            else if (p.reward_kind == 2) p.rw_holding_prev[i] = hold;
            agent_state[2 * i] = (float)hold;                                  // RL_api.py:160-162
            agent_state[2 * i + 1] = seed_f;
            if (want_state) {                                                  // RL_api.py:155-158
                float *so = state_out + i * (2 + p.P);
                so[0] = (float)mand;
                so[1] = (float)hold;
                if (p.P > 0) so[2] = act0 > 0.0 ? 1.f : 0.f;
                if (p.P > 1) so[3] = act1 > 0.0 ? 1.f : 0.f;
                for (int k = 2; k < p.P; ++k) so[2 + k] = p.act[(int64_t)k * p.EN + i] > 0.0 ? 1.f : 0.f;
            }
        }
    }
    __syncwarp();        // a warp only reads the prep records of its own 32 ants
    const uint32_t rock_ants = (LAYOUT == 2) ? __ballot_sync(0xffffffffu, my_rock_flags != 0) : 0u;   // bit a: ant a of the warp

    // ---- phase B: a lane processes one window row of one ant (its S columns); two ways of handing out the rows
    const int warp = tid >> 5, lane = tid & 31;
    const int W = p.W, H = p.H;
    const int nby64m = (p.nby << 6) - 64;
    // ages that still show a value: the table's last entry is the 0 the reference's < 0.01 cut produced
    const uint32_t tab_len = p.tab_len > 0 ? (uint32_t)p.tab_len - 1u : 0u;
    const float decay_c = (float)p.log2_keep;                       // obs = 2^(age * log2(keep)), see below
    const float decay_c2 = -8388608.f * decay_c;
    // 2^23 + table length as a float (8-byte records: live ages are < 2^15, everything else reads >= 2^15)
    const float tab_len_f = __uint_as_float(0x4B000000u | (tab_len < 0x8000u ? tab_len : 0x8000u));
    const bool eager = (REC == 0) && !p.lazy;                       // f64 fields hold plain current values (no decay on read)
    const bool eager_planes = eager && p.diffuse != 0;              // ... in the diffusion planes (sign bit = wall)
    const double inv_max = 1.0 / p.phero_max_val;
    const bool any_plain = !eager && *p.plain_flag != 0u;           // lazy field: plain (non-boxed) values may exist
    // boxed deposit b = box | t: (box | now) - b = now - t = age; zero and plain values give an "age" >= 2^22
    const uint32_t nowb = REC16 ? box32(now_abs) : (REC8 ? now_abs : (now_abs & kBoxMask));
    // 8-byte records: both 15-bit deposit steps of a cell are aged in one subtraction (bit 15 of each half set in the
    // minuend: no borrow crosses the halves)
    const uint32_t nowb2 = ((now_abs & kBox8Mask) | kBox8) * 0x00010001u;
    const uint32_t ogs = obs_gen << 8;
    float *wobs0 = s_obs + warp * kRowsTiles * TILE;
    constexpr bool kLateWait = (UNR >= S);
    constexpr bool ROCKS_WARP = (ANTS_ROWS_ROCKS_WARP != 0) && !FLAT && LAYOUT == 2;
    // one row: sample cells, record loads, decode, staging-tile stores; returns the row's count of unexplored samples
    // phase 2 = the whole row; phase 0 = only issue the record loads into (lo, hi, cell); phase 1 = only consume them
    // exact rock test of one sample cell against the ant's candidate rocks (RL_api.py:132-135): strict sqrt(d2) < r,
    // decided on the squares unless d2 is within 1e-12 of r^2
    auto rock_hit = [&](const RowPrep &q, const int ix, const int iy) -> bool {
        const double ddx = (double)ix - q.rcx, ddy = (double)iy - q.rcy, rad = q.rrad;
        const double d2 = ddx * ddx + ddy * ddy, r2 = rad * rad;
        bool hit = d2 < r2 * 0.999999999999;
        if (!hit && d2 <= r2 * 1.000000000001) hit = sqrt(d2) < rad;
        if (!hit && (q.flags & 4)) hit = rock_channel(p, q.e, q.rocks, ix, iy) != 0.f;
        return hit;
    };
    auto row_body = [&](const RowPrep &q, const double offY, const uint32_t mrow, const uint32_t orow_s,
                        const uint32_t amask, const RowPrep *qn, uint4 (&lo)[UNR], uint4 (&hi)[UNR], uint32_t (&cell)[UNR],
                        const int phase, const uint32_t rbits_in) -> int {
        const double ct = q.ct, st = q.st, xf = q.xf, yf = q.yf;
        const int e = q.e;
        const uint8_t *cells = p.cells + (((int64_t)e * p.plane) << SH);
        if (ANTS_ROWS_BASE64) asm volatile("" : "+l"(cells));
        const double *plane0 = eager_planes ? p.phero_pl + (int64_t)e * 2 * p.plane : nullptr;   // [e][k = 0, 1]
        const double stY = st * offY, ctY = ct * offY;             // RL_api.py:110-111
        const int flags = q.flags;
        uint32_t rbits = rbits_in;                                 // bit j: a rock covers the sample at column j
        int cnt = 0;
        constexpr int RH = S / 2;
        double ctX[RH > 0 ? RH : 1], stX[RH > 0 ? RH : 1];         // ct * X, st * X of the columns right of the centre
        if (ANTS_ROWS_SYM && phase != 1) {
#pragma unroll
            for (int k = 0; k < RH; ++k) { ctX[k] = ct * p.off_c[RH + 1 + k]; stX[k] = st * p.off_c[RH + 1 + k]; }
        }
#pragma unroll
        for (int j0 = 0; j0 < S; j0 += UNR) {
            if (phase != 1) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int j = j0 + u;
                if (j < S) {
                    // sample cell, RL_api.py:110-119: round_half_even(rot(theta + pi/2) * offset + xy_f) mod (W, H)
                    double cX, sX;                         // ct * X, st * X with X = off_c[j] = -off_c[S - 1 - j]
                    if (!ANTS_ROWS_SYM) { const double X = p.off_c[j]; cX = ct * X; sX = st * X; }
                    else if (j == RH) { cX = 0.0; sX = 0.0; }      // (the sign of a zero is lost in the sums below)
                    else if (j > RH) { cX = ctX[j - RH - 1]; sX = stX[j - RH - 1]; }
                    else { cX = -ctX[RH - 1 - j]; sX = -stX[RH - 1 - j]; }
                    const double rx = cX - stY;
                    const double ry = sX + ctY;
                    int ix = round_half_even(rx + xf), iy = round_half_even(ry + yf);
                    ix = wrap1(ix, W); iy = wrap1(iy, H);
                    // cidx(): 8 x 8 blocks of 64 records; (x & 7) * 8 = 8 x - 64 (x >> 3), (y >> 3) * 64 + (y & 7) = y + 56 (y >> 3)
                    cell[u] = (uint32_t)((ix >> 3) * nby64m + ix * 8 + (iy >> 3) * 56 + iy);
                    const uint8_t *rp = cells + ((size_t)cell[u] << SH);
                    if (REC8) {                    // the whole cell in one 64-bit load, four cells per sector
                        const uint2 v8 = *reinterpret_cast<const uint2 *>(rp);
                        lo[u] = make_uint4(v8.x, v8.y, 0u, 0u);
                    } else if (REC == 0 && eager_planes) {  // diffusion: the two pheromone values come from the row-major planes
                        const double *pv = plane0 + (int64_t)ix * p.Hp + iy;
                        const double d0 = pv[0], d1 = pv[p.plane];
                        lo[u] = make_uint4((uint32_t)__double2loint(d0), (uint32_t)__double2hiint(d0) & 0x7FFFFFFFu,
                                           (uint32_t)__double2loint(d1), (uint32_t)__double2hiint(d1) & 0x7FFFFFFFu);
                    } else {
                        lo[u] = ld_record16(rp);
                    }
                    if (REC == 0) hi[u] = *reinterpret_cast<const uint4 *>(rp + 16);
                }
            }
            }
            if (phase == 0) return 0;
            if (ANTS_ROWS_PREFETCH > 0 && qn != nullptr && j0 == 0) {
                // (experiment, off by default) while this row's record loads are in flight: prefetch the same row of the
                // ant this lane serves in the NEXT chunk.  ncu (bench batch) showed 30 % of all warp stall samples on the
                // first use of the loaded records, but the extra address arithmetic costs more than the L2 hits give back.
                const double ct2 = qn->ct, st2 = qn->st, xf2 = qn->xf, yf2 = qn->yf;
                const uint8_t *cells2 = p.cells + (((int64_t)qn->e * p.plane) << SH);
                const double stY2 = st2 * offY, ctY2 = ct2 * offY;
#pragma unroll
                for (int k = 0; k < ANTS_ROWS_PREFETCH && k < S; ++k) {
                    const int j = ANTS_ROWS_PREFETCH >= S ? k : (k * (S - 1)) / (ANTS_ROWS_PREFETCH > 1 ? ANTS_ROWS_PREFETCH - 1 : 1);
                    const double X = p.off_c[j];
                    int ix = round_half_even((ct2 * X - stY2) + xf2), iy = round_half_even((st2 * X + ctY2) + yf2);
                    ix = wrap1(ix, W); iy = wrap1(iy, H);
                    const uint32_t c2 = (uint32_t)((ix >> 3) * nby64m + ix * 8 + (iy >> 3) * 56 + iy);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(cells2 + ((size_t)c2 << SH)));
                }
            }
            if (LAYOUT == 2 && !ROCKS_WARP && flags) {   // a rock may reach this ant's window (few ants): RL_api.py:132-135
#pragma unroll 1
                for (int j = j0; j < S && j < j0 + UNR; ++j) {
                    const double X = p.off_c[j];                           // the same arithmetic as above
                    const double rx = ct * X - stY;
                    const double ry = st * X + ctY;
                    int ix = round_half_even(rx + xf), iy = round_half_even(ry + yf);
                    ix = wrap1(ix, W); iy = wrap1(iy, H);
                    rbits |= (rock_hit(q, ix, iy) ? 1u : 0u) << j;
                }
            }
            if (kLateWait) {
                if (lane == 0) bulk_store_wait_read<FLAT ? 0 : kRowsTiles - 1>();
                __syncwarp(amask);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int j = j0 + u;
                if (j < S) {
                    uint8_t *rp = const_cast<uint8_t *>(cells) + ((size_t)cell[u] << SH);
                    bool wl, occupied, fresh, seen_now, hill;
                    uint32_t age0, age1;
                    float m0 = 0.f, m1 = 0.f;      // 2^23 + age as floats (8-byte records, ANTS_ROWS_MAGIC)
                    float v5;                      // food as the f32 observation shows it
                    if (REC8) {
                        const uint32_t pk = lo[u].y >> 16;                 // [hill|occ][wall|explored]
                        occupied = (pk & 0x7Fu) == occ_gen;
                        hill = (pk & 0x80u) != 0;
                        wl = (pk & 0x8000u) != 0;
                        fresh = (pk & 0x7F00u) == 0u;
                        seen_now = (pk & 0x7F00u) == ogs;
                        v5 = (float)(lo[u].y & 0xFFFFu);                   // (an escaped amount is patched in below)
                        // per half: age = (now - step) mod 2^15 for a boxed deposit (bit 15 set), >= 2^15 (= expired) else
                        const uint32_t d2 = nowb2 - (lo[u].x & 0x7FFF7FFFu);
                        const uint32_t a2 = (d2 & 0x7FFF7FFFu) | (~lo[u].x & 0x80008000u);
                        age0 = a2 & 0xFFFFu; age1 = a2 >> 16;
                        if (ANTS_ROWS_MAGIC) {     // bytes [age lo, age hi, 00, 4B] = the float 2^23 + age
                            m0 = __uint_as_float(__byte_perm(a2, 0x4B000000u, 0x7610u));
                            m1 = __uint_as_float(__byte_perm(a2, 0x4B000000u, 0x7632u));
                        }
                        if (explore_on && fresh) rp[7] = (uint8_t)(((pk >> 8) & 0x80u) | obs_gen);
                    } else if (REC16) {
                        const uint32_t pk = lo[u].w;
                        occupied = (pk & 0x7Fu) == occ_gen;
                        hill = (pk & 0x80u) != 0;
                        wl = (pk & 0x8000u) != 0;
                        fresh = (pk & 0x7F00u) == 0u;
                        seen_now = (pk & 0x7F00u) == ogs;
                        v5 = __uint_as_float(lo[u].z);
                        age0 = nowb - lo[u].x; age1 = nowb - lo[u].y;
                        if (explore_on && fresh) rp[13] = (uint8_t)(((pk >> 8) & 0x80u) | obs_gen);
                    } else {
                        occupied = (hi[u].z >> 16) == occ_gen;
                        wl = (hi[u].w & 1u) != 0;
                        hill = (hi[u].w & 2u) != 0;
                        fresh = (hi[u].z & 0xFFFFu) == 0u;
                        seen_now = (hi[u].z & 0xFFFFu) == obs_gen;
                        v5 = (float)__hiloint2double((int)hi[u].y, (int)hi[u].x);
                        // f64 fields: boxed <=> the high word carries the NaN box; the deposit step is the low word
                        const bool bx0 = p.lazy && (lo[u].y & 0xFFF80000u) == 0x7FF80000u;
                        const bool bx1 = p.lazy && (lo[u].w & 0xFFF80000u) == 0x7FF80000u;
                        age0 = bx0 ? ((nowb - lo[u].x) & kBoxMask) : 0xFFFFFFFFu;
                        age1 = bx1 ? ((nowb - lo[u].z) & kBoxMask) : 0xFFFFFFFFu;
                        if (explore_on && fresh) *reinterpret_cast<uint16_t *>(rp + 24) = (uint16_t)obs_gen;
                    }
                    if (explore_on) cnt += (fresh || seen_now) ? 1 : 0;    // gather-before-scatter, Q7
                    // pheromone channels, RL_api.py:124-125.  A saturated deposit of age k shows
                    // (float)(max_val * keep^k / max_val) = keep^k (the reference's per-step rounding moves it by
                    // ~1e-16 k): evaluated as 2^(k log2 keep) in f32, < 1.2e-6 relative (bar 1e-5); the < 0.01
                    // cut is the exact table length; inside a wall only a deposit of this very update shows.
                    float v1, v2;
                    if (REC8 && ANTS_ROWS_MAGIC) {
                        // (2^23 + age) c - 2^23 c in one fused rounding = the rounded product age * c (2^23 c is exact);
                        // the floats 2^23 + n order like the integers n
                        const float limf = wl ? 8388609.f : tab_len_f;
                        v1 = m0 < limf ? ex2_approx(__fmaf_rn(m0, decay_c, decay_c2)) : 0.f;
                        v2 = m1 < limf ? ex2_approx(__fmaf_rn(m1, decay_c, decay_c2)) : 0.f;
                    } else {
                        const uint32_t lim = wl ? 1u : tab_len;
                        v1 = age0 < lim ? ex2_approx((float)age0 * decay_c) : 0.f;
                        v2 = age1 < lim ? ex2_approx((float)age1 * decay_c) : 0.f;
                    }
                    if (REC == 0 && eager) {       // eager f64 fields (dense / tiles / diffusion): phero / max_val
                        v1 = (float)(__hiloint2double((int)lo[u].y, (int)lo[u].x) * inv_max);
                        v2 = (float)(__hiloint2double((int)lo[u].w, (int)lo[u].z) * inv_max);
                    }
                    const float v0 = occupied ? 1.f : 0.f;                                   // :136-142
                    const float v3 = hill ? 1.f : 0.f;                                       // :130-131 (disc bit of the record)
                    const float v4 = wl ? 1.f : 0.f;                                         // :128-129
                    const float v6 = ((rbits >> j) & 1u) ? 1.f : 0.f;                        // :132-135
                    const uint32_t vis = (mrow >> j) & 1u;     // masked slots keep the -1 written once above
                    const uint32_t oaddr = orow_s + (uint32_t)(j * C * 4);
                    if (LAYOUT == 2)
                        asm volatile("{\n .reg .pred pv;\n setp.ne.u32 pv, %0, 0;\n"
                                     " @pv st.shared.f32 [%1], %2;\n @pv st.shared.f32 [%1+4], %3;\n"
                                     " @pv st.shared.f32 [%1+8], %4;\n @pv st.shared.f32 [%1+12], %5;\n"
                                     " @pv st.shared.f32 [%1+16], %6;\n @pv st.shared.f32 [%1+20], %7;\n"
                                     " @pv st.shared.f32 [%1+24], %8;\n}"
                                     ::"r"(vis), "r"(oaddr), "f"(v0), "f"(v1), "f"(v2), "f"(v3), "f"(v4), "f"(v5), "f"(v6) : "memory");
                    else
                        asm volatile("{\n .reg .pred pv;\n setp.ne.u32 pv, %0, 0;\n"
                                     " @pv st.shared.f32 [%1], %2;\n @pv st.shared.f32 [%1+4], %3;\n"
                                     " @pv st.shared.f32 [%1+8], %4;\n @pv st.shared.f32 [%1+12], %5;\n"
                                     " @pv st.shared.f32 [%1+16], %6;\n @pv st.shared.f32 [%1+20], %7;\n}"
                                     ::"r"(vis), "r"(oaddr), "f"(v0), "f"(v1), "f"(v2), "f"(v3), "f"(v4), "f"(v5) : "memory");
                }
            }
            if (any_plain) {   // plain pheromone values (bool activations, imports, eager modes) may exist: patch them in
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int j = j0 + u;
                    if (j >= S || !((mrow >> j) & 1u)) continue;
                    const uint8_t *rp = cells + ((size_t)cell[u] << SH);
                    bool pl0, pl1;
                    if (REC8) {
                        const uint2 r2 = *reinterpret_cast<const uint2 *>(rp);
                        pl0 = (r2.x & 0xFFFFu) == 1u; pl1 = (r2.x >> 16) == 1u;
                        if ((r2.y & 0xFFFFu) == kFoodEsc)                  // a non-integer amount of food
                            asm volatile("st.shared.f32 [%0+20], %1;" ::"r"(orow_s + (uint32_t)(j * C * 4)), "f"((float)ld_food(p, rp)) : "memory");
                    } else {
                        const uint4 r4 = *reinterpret_cast<const uint4 *>(rp);
                        pl0 = REC16 ? (r4.x != 0u && !is_boxed32(r4.x))
                                    : ((r4.x | r4.y) != 0u && !(p.lazy && (r4.y & 0xFFF80000u) == 0x7FF80000u));
                        pl1 = REC16 ? (r4.y != 0u && !is_boxed32(r4.y))
                                    : ((r4.z | r4.w) != 0u && !(p.lazy && (r4.w & 0xFFF80000u) == 0x7FF80000u));
                    }
                    const uint32_t oaddr = orow_s + (uint32_t)(j * C * 4);
                    if (pl0) asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(oaddr), "f"(phero_obs_slow(p, rp, 0, now, now_abs)) : "memory");
                    if (pl1) asm volatile("st.shared.f32 [%0+8], %1;" ::"r"(oaddr), "f"(phero_obs_slow(p, rp, 1, now, now_abs)) : "memory");
                }
            }
        }
        return cnt;
    };
    // flush n_in ants' staged (S2 x C) f32 observations: one TMA bulk store when 16 B granular, else plain stores
    auto flush = [&](const float *wobs, int64_t i0, int n_in) {
        float *dst = obs + i0 * SC;
        const uint32_t bytes = (uint32_t)(n_in * SC * 4);
        if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) bulk_store_s2g(dst, wobs, bytes);
        } else {
            __syncwarp();
            for (int t = lane; t < n_in * SC; t += 32) dst[t] = wobs[t];
            __syncwarp();
        }
    };
    const int64_t wbase = base + warp * 32;
    const int n_valid = (ant_end - wbase >= 32) ? 32 : (ant_end > wbase ? (int)(ant_end - wbase) : 0);   // ants of this warp
    if (FLAT) {
        // All 32 ants of the warp as 32 * S rows: iteration t hands rows 32t .. 32t+31 to the lanes (no idle lanes:
        // S iterations instead of 32 / 4 chunks of 4 * S <= 32 rows).  Ant a stages into slot (a / 4) & 1 of two
        // 4-ant tiles; with S = 7 an iteration touches chunks t and t + 1 only, and chunk t is complete after it.
        const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(wobs0);
        for (int rr = lane; rr < 2 * ROWS; rr += 32) {             // masked samples read -1 (RL_api.py:147-148): written once
            const int slot = rr / ROWS, rem = rr - slot * ROWS, a4 = rem / S, li = rem - a4 * S;
            const uint32_t m = p.mask_rows[li];
            const uint32_t o = tiles_s + (uint32_t)((slot * TILE + (a4 * S2 + li * S) * C) * 4);
            for (int j = 0; j < S; ++j)
                if (!((m >> j) & 1u))
                    for (int c = 0; c < C; ++c)
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(o + (uint32_t)((j * C + c) * 4)), "f"(-1.f) : "memory");
        }
        __syncwarp();
        const int chunks_total = (n_valid + G - 1) / G;
        int flushed = 0;
        for (int t = 0; t < S; ++t) {
            const int nrows = n_valid * S - 32 * t;                // rows left for this iteration
            if (nrows <= 0) break;
            const int r = 32 * t + lane;
            const int la = r / S, li = r - la * S;                 // ant of the warp, window row
            const uint32_t amask = nrows >= 32 ? 0xffffffffu : ((1u << nrows) - 1u);
            if (lane < nrows) {
                const uint32_t orow_s = tiles_s + (uint32_t)((((la >> 2) & 1) * TILE + ((la & 3) * S2 + li * S) * C) * 4);
                uint4 lo[UNR], hi[UNR];
                uint32_t cell[UNR];
                const int cnt = row_body(prep[warp * 32 + la], p.off_c[li], p.mask_rows[li], orow_s, amask, nullptr, lo, hi, cell, 2, 0u);
                s_rowcnt[(warp * 32 + la) * S + li] = (uint8_t)cnt;
            }
            const bool last = nrows <= 32;
            int c_hi = last ? chunks_total : (32 * (t + 1)) / ROWS;
            if (c_hi > chunks_total) c_hi = chunks_total;
            for (; flushed < c_hi; ++flushed) {
                const int left = n_valid - G * flushed;
                flush(wobs0 + (flushed & 1) * TILE, wbase + G * flushed, left < G ? left : G);
            }
        }
    } else {
        const int la = lane / S, li = lane - la * S;
        const bool lane_on = lane < ROWS;
        const double offY = p.off_c[lane_on ? li : 0];
        uint32_t mrow = p.mask_rows[lane_on ? li : 0];
        // this lane's row of the staging tile (fixed for the whole kernel), as a shared-space address
        uint32_t orow_s0 = (uint32_t)__cvta_generic_to_shared(wobs0 + ((lane_on ? la : 0) * S2 + (lane_on ? li : 0) * S) * C);
        asm volatile("" : "+r"(orow_s0), "+r"(mrow));                // keep them in registers (no rematerialisation)
        // masked samples read -1 in every channel (RL_api.py:147-148) and their tile slots are never written again
        if (lane_on) {
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (!((mrow >> j) & 1u))
#pragma unroll
                    for (int c = 0; c < C; ++c)
#pragma unroll
                        for (int t = 0; t < kRowsTiles; ++t)
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(orow_s0 + (uint32_t)((t * TILE + j * C + c) * 4)), "f"(-1.f) : "memory");
        }
        constexpr bool PIPE = (ANTS_ROWS_PIPE != 0) && REC != 0 && UNR >= S && kRowsTiles == 1;
        uint4 loA[UNR], hiA[UNR], loB[UNR], hiB[UNR];
        uint32_t cellA[UNR], cellB[UNR];
        // the rock channel of the chunk's ants a rock may reach (bits of fa), by the whole warp: the S*S samples of
        // such an ant go over the lanes (the same arithmetic as in row_body, so the same cells), the hits come back
        // to the ant's row lanes as ballots.  Returns this lane's row bits.  (Warp-converged call.)
        auto rock_rows = [&](const int g, uint32_t fa) -> uint32_t {
            uint32_t rb = 0u;
            while (fa) {
                const int a = __ffs((int)fa) - 1;
                fa &= fa - 1u;
                const RowPrep &q = prep[warp * 32 + g + a];
                const double ct = q.ct, st = q.st, xf = q.xf, yf = q.yf;
#pragma unroll 1
                for (int s0 = 0; s0 < S2; s0 += 32) {
                    const int s = s0 + lane;
                    bool hit = false;
                    if (s < S2) {
                        const int i = s / S, j = s - i * S;
                        const double X = p.off_c[j], Y = p.off_c[i];
                        const double stY = st * Y, ctY = ct * Y;
                        const double rx = ct * X - stY;
                        const double ry = st * X + ctY;
                        int ix = round_half_even(rx + xf), iy = round_half_even(ry + yf);
                        ix = wrap1(ix, W); iy = wrap1(iy, H);
                        hit = rock_hit(q, ix, iy);
                    }
                    const uint32_t b = __ballot_sync(0xffffffffu, hit);
                    const int sh = li * S - s0;                    // this lane's row starts at sample li * S of the window
                    if (lane_on && la == a && sh > -S && sh < 32)
                        rb |= (sh >= 0 ? (b >> sh) : (b << -sh)) & ((1u << S) - 1u);
                }
            }
            return rb;
        };
        // one chunk of 4 ants: (issue +) decode + flush
        // (FULLW: all 32 ants of the warp exist, G ants in every chunk; a constant at each inlined call site)
        auto chunk = [&](const bool FULLW, const int g, uint4 (&lo)[UNR], uint4 (&hi)[UNR], uint32_t (&cell)[UNR], const int phase) {
            const int64_t i0 = wbase + g;
            const int n_in = FULLW ? G : ((ant_end - i0 < G) ? (int)(ant_end - i0) : G);
            const int tsel = (kRowsTiles > 1) ? ((g / G) & 1) : 0;
            float *wobs = wobs0 + tsel * TILE;
            const uint32_t orow_s = orow_s0 + (uint32_t)(tsel * TILE * 4);
            const uint32_t amask = (n_in * S >= 32) ? 0xffffffffu : ((1u << (n_in * S)) - 1u);   // the lanes with a row
            if (!kLateWait) {
                if (lane == 0) bulk_store_wait_read<kRowsTiles - 1>();
                __syncwarp();
            }
            const bool row_on = lane_on && (FULLW || la < n_in);
            const RowPrep *qn = (ANTS_ROWS_PREFETCH > 0 && g + G < 32 && i0 + G + la < ant_end) ? &prep[warp * 32 + g + G + la] : nullptr;
            const uint32_t fa = ROCKS_WARP ? ((rock_ants >> g) & ((1u << G) - 1u)) : 0u;     // warp-uniform
            uint32_t rb = 0u;
            if (ROCKS_WARP && phase == 2 && UNR >= S) {
                // record loads first, the (rare) rock evaluation while they are in flight, then the decode
                if (row_on) row_body(prep[warp * 32 + g + la], offY, mrow, orow_s, amask, qn, lo, hi, cell, 0, 0u);
                if (fa) rb = rock_rows(g, fa);
                if (row_on) s_rowcnt[(warp * 32 + g + la) * S + li] =
                    (uint8_t)row_body(prep[warp * 32 + g + la], offY, mrow, orow_s, amask, qn, lo, hi, cell, 1, rb);
            } else {
                if (ROCKS_WARP && fa && phase != 0) rb = rock_rows(g, fa);
                if (row_on) s_rowcnt[(warp * 32 + g + la) * S + li] =
                    (uint8_t)row_body(prep[warp * 32 + g + la], offY, mrow, orow_s, amask, qn, lo, hi, cell, phase, rb);
            }
            flush(wobs, i0, n_in);
        };
        // the record loads of a chunk alone
        auto issue = [&](const int g, uint4 (&lo)[UNR], uint4 (&hi)[UNR], uint32_t (&cell)[UNR]) {
            const int64_t i0 = wbase + g;
            if (g < 32 && i0 < ant_end && lane_on && i0 + la < ant_end)
                row_body(prep[warp * 32 + g + la], offY, mrow, 0u, 0u, nullptr, lo, hi, cell, 0, 0u);
        };
        if (PIPE) {
            // chunk g + 1 is on its way while chunk g is decoded: two register sets, the loop unrolled by two so that
            // neither is indexed dynamically
            issue(0, loA, hiA, cellA);
            for (int g = 0; g < 32; g += 2 * G) {
                if (wbase + g >= ant_end) break;
                issue(g + G, loB, hiB, cellB);
                chunk(false, g, loA, hiA, cellA, 1);
                if (wbase + g + G >= ant_end) break;
                issue(g + 2 * G, loA, hiA, cellA);
                chunk(false, g + G, loB, hiB, cellB, 1);
            }
        } else if (ANTS_ROWS_FULLWARP && n_valid == 32) {
#pragma unroll 1
            for (int g = 0; g < 32; g += G) chunk(true, g, loA, hiA, cellA, 2);
        } else {
#pragma unroll 1
            for (int g = 0; g < 32; g += G) {
                if (wbase + g >= ant_end) break;
                chunk(false, g, loA, hiA, cellA, 2);
            }
        }
    }
    __syncwarp();

    // ---- phase C: reward epilogue, thread per ant (this warp's own ants)
    {
        const int64_t i = base + tid;
        if (i < ant_end) {
            int count = 0;
            if (explore_on) {
#pragma unroll
                for (int k = 0; k < S; ++k) count += s_rowcnt[tid * S + k];
            }
            double reward;
            if (p.reward_kind == 1) {
                reward = (double)count / 10.0;                                 // reward_custom.py:19
            } else if (p.reward_kind == 0) {
                reward = 0.0;
                if (explore_on) reward += ((double)count / 10.0) * r_mult;     // :89-94
                reward += r_other;                                             // :106
            } else {
                reward = r_other;
            }
            p.rewards[i] = reward;
            if (reward_out != nullptr) reward_out[i] = reward;
            if (is_step) {                                                     // ants.py:119-121 (Q16)
                int rs = rs_prev;                                              // (loaded in phase A)
                rs += ((reward - p.reward_threshold) > 0.0) ? 255 : 0;
                p.reward_state[i] = (uint8_t)(rs > 255 ? 255 : rs);
            }
        }
    }
    if (lane == 0) bulk_store_wait_read();   // smem must stay valid until the last bulk store has read it
}

}  // namespace ants
