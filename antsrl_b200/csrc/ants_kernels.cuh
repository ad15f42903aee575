// ants_kernels.cuh -- sm_100a device code of the AntsRL step loop (one batch of E independent environments).
//
// Design: ant-centric flat kernels over all E*N ants.  The reference's "last writer wins"
// fancy-index scatters (quirk Q1: ants.py:116, pheromone.py:39, RL_api.py:141) and its gather-before-scatter
// exploration reward (Q7: reward_custom.py:89-93) are resolved without per-environment barriers through
// generation-stamped planes:
//   owner[e][x][y] : u32 = (phase << 16) | ant   written with atomicMax -> highest ant index of the current
//                    scatter phase owns the cell (food pickup/drop, pheromone deposit)
//   meta[e][x][y]  : u32 = (occ_gen << 16) | explored_gen; occ_gen == current step  => an ant stands here;
//                    explored_gen == 0 or == current observation => the cell was unexplored before this observation
// All position / angle / sample-coordinate arithmetic is f64 in the reference's operation order (compiled with
// -fmad=false) so that truncated / rounded cell indices agree with numpy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ants {

constexpr int kMaxCh = 16;
constexpr int kTile = 16;                 // pheromone activity tile: 16 x 16 cells

struct Params {
    int32_t E, N, W, H, Hp, P, R;
    int64_t EN;                            // E * N
    int64_t plane;                         // W * Hp cells per env plane
    int32_t radius, S, S2, C;
    int32_t has_mask;
    int32_t ch_kind[kMaxCh];
    int32_t ch_arg[kMaxCh];
    int32_t rule_n;
    int32_t rule_op[kMaxCh];               // mandible rule in perceived_objects order: 0 = food OR, 1 = anthill AND
    int32_t reward_kind, explore_on, has_max_val;
    int32_t tiles_x, tiles_y;              // activity tiles per plane
    double delta, fwd_delta, reward_threshold, max_speed, max_rot_speed, csr, bsr;
    double f_explore, f_food, f_anthill, f_explore_hold, f_heading;
    double filt_center, filt_ring, phero_max_val, max_hold;
    uint64_t rng_seed;
    int64_t env_id_base;
    // ants, [E*N]
    double *x, *y, *theta, *prev_x, *prev_y, *prev_theta, *holding, *seed;
    double *act;                           // [P][E*N]
    uint8_t *mandibles, *reward_state;
    double *rw_holding_prev, *rw_prev_dist, *rewards;
    // planes, row pitch Hp
    double *phero;                         // [E][P][W][Hp]
    double *phero_alt;                     // ping-pong target of the diffusion stencil (DIFFUSE_FACTOR != 0)
    double *food;                          // [E][W][Hp]
    uint8_t *walls;                        // [E][W][Hp]
    uint32_t *meta, *owner;                // [E][W][Hp]
    uint8_t *tile_active;                  // [E][P][tiles_x][tiles_y]
    int32_t *hill;                         // [E][4] = x, y, r, r*r
    double *hill_food;                     // [E]
    double *rock_c, *rock_rad, *rock_w;    // [E][R][2], [E][R], [E][R]
    unsigned long long *rock_grid;         // [E][ceil(W/32)][ceil(H/32)] bitmask of rocks near each 32x32 block
    const double *samp_px, *samp_py;       // [S2] perception_coords * DELTA (RL_api.py:92-93)
    const uint8_t *mask;                   // [S2]
    // scratch
    double *food_delta;                    // [E*N]
    uint32_t *commit_list, *commit_count;  // ants whose mandible action changes the food plane this step
    uint32_t *absorb_list, *absorb_count;  // (env, cell) pairs of food lying inside the anthill disc
    unsigned long long *tile_counter;      // tiles processed by the evaporation kernel
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ double pymod(double a, double b) {      // np.mod on floats (npy_remainder)
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0.0) != (r < 0.0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}
__device__ __forceinline__ int imod(int a, int n) {                // np.mod on ints
    int r = a % n;
    return r < 0 ? r + n : r;
}
// ndarray.astype(int) of a coordinate in [0, n]; n itself (quirk Q14: np.mod can return exactly n) maps to
// cell 0, as the reference's own occupancy map does (RL_api.py:137-139).
__device__ __forceinline__ int cell_of(double v, int n) {
    int c = (int)v;
    return c >= n ? c - n : c;
}
__device__ __forceinline__ bool in_hill(const int32_t *hl, int cx, int cy) {   // anthill.py:31-33 on integers
    int dx = hl[0] - cx, dy = hl[1] - cy;
    return dx * dx + dy * dy <= hl[3];
}

// Philox4x32-10, counter (ant, step, env, 0), key (seed_lo, seed_hi) -> one double in [0,1) built like
// numpy's random_sample: ((a >> 5) * 2^26 + (b >> 6)) / 2^53.  Mirrored by oracle.philox_uniform.
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint32_t env, uint32_t step, uint32_t ant) {
    uint32_t c0 = ant, c1 = step, c2 = env, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    double a = (double)(c0 >> 5), b = (double)(c1 >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

// ------------------------------------------------------------------------------------------------ step, part 1
// RLApi.step lines 178-196 for every ant: mandible rule, pickup / drop bookkeeping (ants.py:102-117),
// pheromone activation (ants.py:89-96), rotation (ants.py:62-67), forward move on the torus (ants.py:69-80),
// plus the occupancy stamp consumed by the "ants" perception channel (RL_api.py:136-142).
__global__ void __launch_bounds__(256)
k_step_move(Params p, const int8_t *__restrict__ rot, const int8_t *__restrict__ ph, uint32_t owner_stamp,
            uint32_t occ_gen, int all_stamp, double act_on) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.EN) return;
    int e = (int)(i / p.N);
    int a = (int)(i - (int64_t)e * p.N);
    double x = p.x[i], y = p.y[i], th = p.theta[i];
    double hold = p.holding[i];
    int m_old = p.mandibles[i] != 0;
    int pcx = cell_of(p.prev_x[i], p.W), pcy = cell_of(p.prev_y[i], p.H);      // RL_api.py:178
    int64_t pcell = (int64_t)e * p.plane + (int64_t)pcx * p.Hp + pcy;
    double f = p.food[pcell];
    const int32_t *hl = p.hill + 4 * e;
    bool hill = in_hill(hl, cell_of(x, p.W), cell_of(y, p.H));                  // RL_api.py:184
    int m = m_old;
    for (int k = 0; k < p.rule_n; ++k) {                                        // RL_api.py:180-184 (Q5)
        if (p.rule_op[k] == 0) m |= (f > 0.0) ? 1 : 0;
        else m &= hill ? 0 : 1;
    }
    bool closing = m && !m_old, opening = !m && m_old;                          // ants.py:103-104
    double taken = closing ? fmin(p.max_hold, fmax(0.0, f)) : 0.0;              // ants.py:111
    double dropped = opening ? hold : 0.0;                                      // ants.py:114
    double delta = dropped - taken;
    hold = hold + (taken - dropped);                                            // ants.py:117
    // ants.py:116 `qte[xy] += dropped - taken`: the highest ant index standing in the cell wins (Q1).  Ants in
    // a cell that holds no food and is outside the hill cannot change it and need not compete.
    if (all_stamp || f > 0.0 || hill) atomicMax(p.owner + pcell, owner_stamp | (uint32_t)a);
    if (delta != 0.0) {
        uint32_t slot = atomicAdd(p.commit_count, 1u);
        p.commit_list[slot] = (uint32_t)i;
        p.food_delta[i] = delta;
    }
    p.mandibles[i] = (uint8_t)m;
    p.holding[i] = hold;
    if (ph != nullptr) {                                                        // ants.py:89-96
        int v = ph[i];
        p.act[i] = (v == 1) ? act_on : 0.0;
        p.act[p.EN + i] = (v != 0 && v != 1) ? act_on : 0.0;
    }
    if (rot != nullptr) th = pymod(th + (double)rot[i] * p.max_rot_speed, 6.283185307179586);   // ants.py:62-67
    double fwd = (1.0 * p.max_speed) * (1.0 - hold * p.csr);                    // RL_api.py:194
    if (fwd < 0.0) fwd *= p.bsr;                                                // RL_api.py:195
    double s, c;
    sincos(th, &s, &c);
    x = pymod(x + c * fwd, (double)p.W);                                        // ants.py:69-80
    y = pymod(y + s * fwd, (double)p.H);
    p.x[i] = x; p.y[i] = y; p.theta[i] = th;
    int64_t ocell = (int64_t)e * p.plane + (int64_t)cell_of(x, p.W) * p.Hp + cell_of(y, p.H);
    reinterpret_cast<uint16_t *>(p.meta + ocell)[1] = (uint16_t)occ_gen;
}

// Occupancy stamp alone, for a stand-alone observation() (main.py:88).
__global__ void __launch_bounds__(256) k_occ_stamp(Params p, uint32_t occ_gen) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.EN) return;
    int e = (int)(i / p.N);
    int64_t ocell = (int64_t)e * p.plane + (int64_t)cell_of(p.x[i], p.W) * p.Hp + cell_of(p.y[i], p.H);
    reinterpret_cast<uint16_t *>(p.meta + ocell)[1] = (uint16_t)occ_gen;
}

// ------------------------------------------------------------------------------------------------ step, part 2
// The winners of the food scatter write `old + (dropped - taken)` (ants.py:116); food that now lies inside the
// anthill disc is queued for Anthill.update (anthill.py:41-46).
__global__ void __launch_bounds__(128) k_food_commit(Params p, uint32_t owner_stamp) {
    uint32_t n = *p.commit_count;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        int64_t i = p.commit_list[k];
        int e = (int)(i / p.N);
        int a = (int)(i - (int64_t)e * p.N);
        int pcx = cell_of(p.prev_x[i], p.W), pcy = cell_of(p.prev_y[i], p.H);
        int64_t pcell = (int64_t)e * p.plane + (int64_t)pcx * p.Hp + pcy;
        if (p.owner[pcell] != (owner_stamp | (uint32_t)a)) continue;
        double nv = p.food[pcell] + p.food_delta[i];
        p.food[pcell] = nv;
        if (nv != 0.0 && in_hill(p.hill + 4 * e, pcx, pcy)) {
            uint32_t s = atomicAdd(p.absorb_count, 1u);
            p.absorb_list[2 * s] = (uint32_t)e;
            p.absorb_list[2 * s + 1] = (uint32_t)(pcx * p.Hp + pcy);
        }
    }
}

// ------------------------------------------------------------------------------------------------ perception + reward
// RLApi.observation (RL_api.py:96-165) fused with reward.observation (reward_custom.py) and, when called from
// step, Ants.give_reward (ants.py:119-121).  Block = 256 threads = 256 consecutive ants.
//   phase A (thread per ant): f64 trigonometry of the rotated sampling frame, reward terms that do not need
//            the exploration count, agent_state / state outputs, rock culling.
//   phase B (warp per ant, lanes = samples): sample cell = round(rot(theta + pi/2) * offset + xy_f) mod (W,H),
//            gathers from the meta / pheromone / food / wall planes, analytic anthill disc, exact rock test on
//            the culled set, exploration count by ballot, mask, then the (S2 x C) f32 tile is staged in shared
//            memory and written with coalesced stores.
constexpr int kPerceiveThreads = 128;      // 4 warps = 128 consecutive ants per block
constexpr int kMaxGroup = 8;                // ants staged per TMA bulk store (4 or 8 ants: bytes are a multiple of 16)
constexpr int kGridShift = 5;               // rock grid cell = 32 x 32 map cells

struct AntPrep {
    double r_other, mult;        // reward terms without the exploration count; exploration multiplier
    unsigned long long rocks;    // candidate rocks (from the rock grid)
    int e, pad;                  // environment of the ant
};

// wrap an integer sample coordinate onto the torus (np.mod on ints); the fast path covers |offset| < n
__device__ __forceinline__ int wrap_coord(int v, int n) {
    if (v < 0) v += n; else if (v >= n) v -= n;
    return ((unsigned)v < (unsigned)n) ? v : imod(v, n);
}

// shared -> global bulk copy through the async proxy (TMA 1-D bulk store, SASS UBLKCP)
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// Candidate rocks of a point: every rock whose (radius + reach) box overlaps the point's 32x32 grid cell was
// OR-ed into that cell by rock_grid_mark, so one 8-byte load replaces a loop over all rocks.
__device__ __forceinline__ unsigned long long rock_candidates(const Params &p, int e, double x, double y) {
    const int gw = (p.W + 31) >> kGridShift, gh = (p.H + 31) >> kGridShift;
    int gx = cell_of(pymod(x, (double)p.W), p.W) >> kGridShift;
    int gy = cell_of(pymod(y, (double)p.H), p.H) >> kGridShift;
    return p.rock_grid[((int64_t)e * gw + gx) * gh + gy];
}
__device__ __forceinline__ void rock_grid_mark(const Params &p, int e, int r, double cx, double cy, double rad) {
    const int gw = (p.W + 31) >> kGridShift, gh = (p.H + 31) >> kGridShift;
    const double L = rad + ((double)p.radius * p.delta * 1.4142135623730951 + 1.75);
    unsigned long long *g = p.rock_grid + (int64_t)e * gw * gh;
    const unsigned long long bit = 1ull << r;
    // sample the box [c - L, c + L] every <= 32 cells (and at its far edge): hits every grid cell it overlaps
    for (double ox = -L;; ox += 32.0) {
        if (ox > L) ox = L;
        int gx = cell_of(pymod(cx + ox, (double)p.W), p.W) >> kGridShift;
        for (double oy = -L;; oy += 32.0) {
            if (oy > L) oy = L;
            int gy = cell_of(pymod(cy + oy, (double)p.H), p.H) >> kGridShift;
            atomicOr(g + gx * gh + gy, bit);
            if (oy >= L) break;
        }
        if (ox >= L) break;
    }
}

// LAYOUT 0: any perceived_objects list (switch per channel)
// LAYOUT 1: the generator's default list [ants, phero0, phero1, anthill, walls, food] (environment_generator.py:64-99)
// LAYOUT 2: the same plus rocks as 7th channel
// SFIX: perception window side known at compile time (7 = the default 7x7 window), 0 = run-time side.
//
// Block = blockDim.x threads = as many consecutive ants; warp w owns ants [32w, 32w+32) from start to end (no
// block barrier after the table load).
//   phase A (thread per ant): f64 trigonometry of the rotated sampling frame; ALL sample cells of the ant,
//            round(rot(theta + pi/2) * offset + xy_f) mod (W,H), with the products cos*offset / sin*offset
//            computed once per row / column, packed (x << 16 | y) into shared memory; reward terms that do not
//            need the exploration count; agent_state / state outputs; candidate rocks from the rock grid.
//   phase B (warp, flat sample index): the warp walks its ants in chunks of `group` ants; the chunk's
//            group*S2 samples are spread over the 32 lanes and processed in batches of U lane-slots whose
//            gathers (meta / pheromone / food / walls planes) are all issued before any is consumed; analytic
//            anthill disc, exact rock test on the candidates; exploration counts per ant with match_any + shared
//            atomics; the chunk's (group x S2 x C) f32 tile is staged in shared memory and leaves with ONE TMA
//            bulk store.
//   phase C (thread per ant): reward epilogue and Ants.give_reward.
template <int LAYOUT, int SFIX>
__global__ void __launch_bounds__(kPerceiveThreads)
k_perceive(Params p, float *__restrict__ obs, float *__restrict__ agent_state, float *__restrict__ state_out,
           double *__restrict__ reward_out, uint32_t obs_gen, uint32_t occ_gen, int is_step, int rw_alias,
           int group, uint32_t s2_magic, int dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = SFIX ? SFIX : p.S;
    const int S2 = S * S, C = p.C;
    const int SC = S2 * C;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    float *s_obs = reinterpret_cast<float *>(smem_raw);                        // [warps][group][S2*C], 16 B aligned
    uint32_t *s_cell = reinterpret_cast<uint32_t *>(s_obs + nwarps * group * SC);   // [threads][S2]
    AntPrep *prep = reinterpret_cast<AntPrep *>(s_cell + nthreads * S2 + ((nthreads * S2) & 1));
    double *s_off = reinterpret_cast<double *>(prep + nthreads);               // [S] = (k - r) * DELTA
    int *s_cnt = reinterpret_cast<int *>(s_off + S);
    uint8_t *s_mask = reinterpret_cast<uint8_t *>(s_cnt + nthreads);

    const int tid = threadIdx.x;
    for (int k = tid; k < S2; k += nthreads) s_mask[k] = p.has_mask ? p.mask[k] : 1;
    for (int k = tid; k < S; k += nthreads) s_off[k] = p.samp_px[k];          // row 0 of perception_coords[..., 0]
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * nthreads;

    // ---- phase A
    {
        const int64_t i = base + tid;
        s_cnt[tid] = 0;
        if (i < p.EN) {
            const int e = (int)(i / p.N);
            const double x = p.x[i], y = p.y[i], th = p.theta[i], hold = p.holding[i];
            double s0, c0, st, ct;
            sincos(th, &s0, &c0);
            sincos(th + 3.141592653589793 * 0.5, &st, &ct);                    // RL_api.py:101,107-108
            double xf = x, yf = y;
            if (p.fwd_delta != 0.0) { xf = x + c0 * p.fwd_delta; yf = y + s0 * p.fwd_delta; }   // :103-104
            // sample cells, RL_api.py:110-119: rel_x = cos*X - sin*Y, rel_y = sin*X + cos*Y with X = off[j],
            // Y = off[i]; abs = round_half_even(rel + xy_f) mod (W, H)
            uint32_t *mycell = s_cell + tid * S2;
            if (SFIX) {
                double cX[SFIX ? SFIX : 1], sX[SFIX ? SFIX : 1];
#pragma unroll
                for (int j = 0; j < SFIX; ++j) { const double o = s_off[j]; cX[j] = ct * o; sX[j] = st * o; }
#pragma unroll
                for (int ii = 0; ii < SFIX; ++ii) {
                    const double o = s_off[ii];
                    const double sY = st * o, cY = ct * o;
#pragma unroll
                    for (int j = 0; j < SFIX; ++j) {
                        const int ix = wrap_coord(__double2int_rn((cX[j] - sY) + xf), p.W);
                        const int iy = wrap_coord(__double2int_rn((sX[j] + cY) + yf), p.H);
                        mycell[ii * SFIX + j] = ((uint32_t)ix << 16) | (uint32_t)iy;
                    }
                }
            } else {
                for (int ii = 0; ii < S; ++ii) {
                    const double oy = s_off[ii];
                    const double sY = st * oy, cY = ct * oy;
                    for (int j = 0; j < S; ++j) {
                        const double ox = s_off[j];
                        const int ix = wrap_coord(__double2int_rn((ct * ox - sY) + xf), p.W);
                        const int iy = wrap_coord(__double2int_rn((st * ox + cY) + yf), p.H);
                        mycell[ii * S + j] = ((uint32_t)ix << 16) | (uint32_t)iy;
                    }
                }
            }
            AntPrep q;
            q.e = e; q.pad = 0;
            // reward.observation, reward_custom.py
            const double hprev = rw_alias ? hold : p.rw_holding_prev[i];       // Q18
            const double d = hold - hprev;
            q.r_other = 0.0; q.mult = 1.0;
            if (p.reward_kind == 0) {                                          // All_Rewards, :79-106
                const double r_food = d < 0.0 ? 0.0 : d;
                const double r_hill = d < 0.0 ? 1.0 : 0.0;
                const int32_t *hl = p.hill + 4 * e;
                const double ddx = x - (double)hl[0], ddy = y - (double)hl[1];
                const double nd = sqrt(ddx * ddx + ddy * ddy);
                const double heading = (p.rw_prev_dist[i] > nd && hold > 0.0) ? 0.1 : 0.0;
                p.rw_prev_dist[i] = nd;
                q.r_other = r_food * p.f_food + r_hill * p.f_anthill + heading * p.f_heading;
                q.mult = (hold == 0.0) ? p.f_explore : p.f_explore_hold;
                p.rw_holding_prev[i] = hold;
            } else if (p.reward_kind == 2) {                                   // Food_Reward, :37-40
                q.r_other = d < 0.0 ? 10.0 : d;
                p.rw_holding_prev[i] = hold;
            }
            q.rocks = (p.R > 0 && (LAYOUT == 2 || LAYOUT == 0)) ? rock_candidates(p, e, xf, yf) : 0ull;
            prep[tid] = q;
            agent_state[2 * i] = (float)hold;                                  // RL_api.py:160-162
            agent_state[2 * i + 1] = (float)p.seed[i];
            if (state_out != nullptr) {                                        // RL_api.py:155-158
                float *so = state_out + i * (2 + p.P);
                so[0] = (float)p.mandibles[i];
                so[1] = (float)hold;
                for (int k = 0; k < p.P; ++k) so[2 + k] = p.act[(int64_t)k * p.EN + i] > 0.0 ? 1.f : 0.f;
            }
        }
    }
    // which planes the configured channels read (block-uniform)
    bool need_food = LAYOUT != 0, need_walls = LAYOUT != 0, need_meta = LAYOUT != 0 || p.explore_on != 0;
    unsigned ph_need = LAYOUT != 0 ? 3u : 0u;
    if (dbg & 1) { need_food = need_walls = need_meta = false; ph_need = 0; }
    if (dbg & 16) { need_food = need_walls = false; ph_need = 0; }
    if (LAYOUT == 0) {
        for (int c = 0; c < C; ++c) {
            int kd = p.ch_kind[c];
            need_food |= kd == 4; need_walls |= kd == 3; need_meta |= kd == 0;
            if (kd == 1) ph_need |= 1u << p.ch_arg[c];
        }
    }
    constexpr int NPH = LAYOUT != 0 ? 2 : 4;
    const double inv_max = 1.0 / p.phero_max_val;   // obs is f32: x * (1/max) == x / max to well below 1e-5
    const bool has_mask = p.has_mask != 0;
    const bool explore_on = p.explore_on != 0;
    __syncwarp();

    // ---- phase B
    const int warp = tid >> 5, lane = tid & 31;
    float *wobs = s_obs + warp * group * SC;
    const uint32_t *wcell = s_cell + warp * 32 * S2;
    const AntPrep *wprep = prep + warp * 32;
    for (int g = 0; g < 32; g += group) {
        if (dbg & 8) break;
        const int64_t i0 = base + warp * 32 + g;
        if (i0 >= p.EN) break;
        int n_in = 32 - g < group ? 32 - g : group;
        if (p.EN - i0 < n_in) n_in = (int)(p.EN - i0);
        const int nsamp = n_in * S2;
        // the previous chunk's bulk store must have finished reading the staging tile
        if (lane == 0) bulk_store_wait_read();
        __syncwarp();
        const int iters = (nsamp + 31) >> 5;
        constexpr int U = 4;                       // lane-slots per batch: all gathers of a batch are in flight together
        for (int it0 = 0; it0 < iters; it0 += U) {
            int fs[U], es[U];
            uint32_t cw[U], mt[U];
            double fd[U], ph[U][NPH];
            uint8_t wl[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (it0 + u >= iters) break;                                   // warp-uniform
                const int f_raw = (it0 + u) * 32 + lane;
                const int f = f_raw < nsamp ? f_raw : 0;                       // idle tail lanes shadow sample 0
                const int aj = (int)(((uint32_t)f * s2_magic) >> 20);         // f / S2
                const uint32_t c2 = wcell[g * S2 + f];
                const int e = wprep[g + aj].e;
                const int cell = (int)(c2 >> 16) * p.Hp + (int)(c2 & 0xFFFFu);
                const int64_t eoff = (int64_t)e * p.plane + cell;
                fs[u] = f_raw; es[u] = e; cw[u] = c2;
                mt[u] = need_meta ? p.meta[eoff] : 0u;
                fd[u] = need_food ? p.food[eoff] : 0.0;
                wl[u] = need_walls ? p.walls[eoff] : (uint8_t)0;
                const double *php = p.phero + (int64_t)e * p.P * p.plane + cell;
#pragma unroll
                for (int kp = 0; kp < NPH; ++kp) ph[u][kp] = ((ph_need >> kp) & 1u) ? php[(int64_t)kp * p.plane] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (it0 + u >= iters) break;                                   // warp-uniform
                const bool valid = fs[u] < nsamp;
                const int f = valid ? fs[u] : 0;
                const int aj = (int)(((uint32_t)f * s2_magic) >> 20);
                const int e = es[u];
                const int ix = (int)(cw[u] >> 16), iy = (int)(cw[u] & 0xFFFFu);
                if (explore_on) {
                    const uint32_t eg = mt[u] & 0xFFFFu;
                    const bool unexplored = valid && ((eg == 0u) || (eg == obs_gen));     // gather-before-scatter, Q7
                    if (valid && eg == 0u && !(dbg & 4))
                        reinterpret_cast<uint16_t *>(p.meta + (int64_t)e * p.plane + ix * p.Hp + iy)[0] = (uint16_t)obs_gen;
                    const unsigned peers = __match_any_sync(0xffffffffu, aj);
                    const unsigned votes = __ballot_sync(0xffffffffu, unexplored) & peers;
                    if (votes && lane == __ffs(peers) - 1) atomicAdd(s_cnt + warp * 32 + g + aj, __popc(votes));
                }
                if (!valid) continue;
                const int s = f - aj * S2;
                const bool vis = s_mask[s] != 0;
                float *o = wobs + f * C;
                const int32_t *hl = p.hill + 4 * e;
                // mask * (perception + 1) - 1 (RL_api.py:147-148): the +1-1 round trip changes a value by at most
                // 2^-53 absolute, far below the resolution of the f32 observation, so visible samples pass through.
                if (LAYOUT != 0) {
                    if (vis) {
                        o[0] = ((mt[u] >> 16) == occ_gen) ? 1.f : 0.f;                                // :136-142
                        o[1] = (float)(ph[u][0] * inv_max);                                           // :124-125
                        o[2] = (float)(ph[u][1] * inv_max);
                        o[3] = in_hill(hl, ix, iy) ? 1.f : 0.f;                                       // :130-131
                        o[4] = wl[u] ? 1.f : 0.f;                                                     // :128-129
                        o[5] = (float)fd[u];                                                          // :126-127
                        if (LAYOUT == 2) {                                                            // :132-135
                            float rv = 0.f;
                            unsigned long long rm = wprep[g + aj].rocks;
                            const double *rc = p.rock_c + (int64_t)e * p.R * 2;
                            const double *rr = p.rock_rad + (int64_t)e * p.R;
                            while (rm) {
                                int r = __ffsll((long long)rm) - 1;
                                rm &= rm - 1;
                                double ddx = (double)ix - rc[2 * r], ddy = (double)iy - rc[2 * r + 1];
                                if (sqrt(ddx * ddx + ddy * ddy) < rr[r]) { rv = 1.f; break; }
                            }
                            o[6] = rv;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < (LAYOUT == 2 ? 7 : 6); ++c) o[c] = -1.f;
                    }
                } else {
                    for (int c = 0; c < C; ++c) {
                        double v;
                        switch (p.ch_kind[c]) {
                            case 0: v = ((mt[u] >> 16) == occ_gen) ? 1.0 : 0.0; break;
                            case 1: {
                                int a = p.ch_arg[c];
                                double pv = a == 0 ? ph[u][0] : a == 1 ? ph[u][1] : a == 2 ? ph[u][NPH > 2 ? 2 : 0]
                                                                                        : ph[u][NPH > 3 ? 3 : 0];
                                v = pv * inv_max;
                                break;
                            }
                            case 2: v = in_hill(hl, ix, iy) ? 1.0 : 0.0; break;
                            case 3: v = wl[u] ? 1.0 : 0.0; break;
                            case 4: v = fd[u]; break;
                            default: {
                                v = 0.0;
                                unsigned long long rm = wprep[g + aj].rocks;
                                const double *rc = p.rock_c + (int64_t)e * p.R * 2;
                                const double *rr = p.rock_rad + (int64_t)e * p.R;
                                while (rm) {
                                    int r = __ffsll((long long)rm) - 1;
                                    rm &= rm - 1;
                                    double ddx = (double)ix - rc[2 * r], ddy = (double)iy - rc[2 * r + 1];
                                    if (sqrt(ddx * ddx + ddy * ddy) < rr[r]) { v = 1.0; break; }
                                }
                            }
                        }
                        o[c] = (has_mask && !vis) ? -1.f : (float)v;
                    }
                }
            }
        }
        // flush the staged (n_in x S2 x C) f32 tile: one TMA bulk store when 16 B granular, else plain stores
        float *dst = obs + i0 * SC;
        const uint32_t bytes = (uint32_t)(n_in * SC * 4);
        if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(dbg & 2)) bulk_store_s2g(dst, wobs, bytes);
        } else {
            __syncwarp();
            for (int t = lane; t < n_in * SC; t += 32) dst[t] = wobs[t];
            __syncwarp();
        }
    }
    __syncwarp();

    // ---- phase C: reward epilogue, thread per ant (this warp's own ants)
    {
        const int64_t i = base + tid;
        if (i < p.EN) {
            const AntPrep &q = prep[tid];
            const int count = s_cnt[tid];
            double reward;
            if (p.reward_kind == 1) {
                reward = (double)count / 10.0;                                 // reward_custom.py:19
            } else if (p.reward_kind == 0) {
                reward = 0.0;
                if (explore_on) reward += ((double)count / 10.0) * q.mult;     // :89-94
                reward += q.r_other;                                           // :106
            } else {
                reward = q.r_other;
            }
            p.rewards[i] = reward;
            if (reward_out != nullptr) reward_out[i] = reward;
            if (is_step) {                                                     // ants.py:119-121 (Q16)
                int rs = p.reward_state[i];
                rs += ((reward - p.reward_threshold) > 0.0) ? 255 : 0;
                p.reward_state[i] = (uint8_t)(rs > 255 ? 255 : rs);
            }
        }
    }
    if (lane == 0) bulk_store_wait_read();   // smem must stay valid until the last bulk store has read it
}

// ------------------------------------------------------------------------------------------------ update, ants side
// Walls.update for ants (walls.py:24-28) and, when there are no rocks, the ant part of Ants.update
// (ants.py:124,130) plus the ownership stamp of the pheromone deposit (pheromone.py:39, Q1).
__global__ void __launch_bounds__(256)
k_collide(Params p, const double *__restrict__ noise, uint32_t step_id, uint32_t owner_stamp, int finish) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.EN) return;
    int e = (int)(i / p.N);
    int a = (int)(i - (int64_t)e * p.N);
    double x = p.x[i], y = p.y[i], th = p.theta[i];
    int64_t eoff = (int64_t)e * p.plane;
    if (p.walls[eoff + (int64_t)cell_of(x, p.W) * p.Hp + cell_of(y, p.H)]) {
        x = p.prev_x[i]; y = p.prev_y[i];
        double u = noise ? noise[i] : philox_uniform(p.rng_seed, (uint32_t)(p.env_id_base + e), step_id, (uint32_t)a);
        th += u - 0.5;                                                         // not re-wrapped (Q3)
        p.x[i] = x; p.y[i] = y; p.theta[i] = th;
    }
    if (finish) {
        p.prev_x[i] = x; p.prev_y[i] = y; p.prev_theta[i] = th;
        atomicMax(p.owner + eoff + (int64_t)cell_of(x, p.W) * p.Hp + cell_of(y, p.H), owner_stamp | (uint32_t)a);
        p.reward_state[i] = (uint8_t)((double)p.reward_state[i] * 0.9);        // ants.py:130
    }
}

// CircleObstacles.update, first half (circle_obstacles.py:35-40): ants push rocks.  One block per env, one warp
// per rock (looping); contributions are summed in ant order like np.sum(axis=0).
__global__ void __launch_bounds__(256) k_rocks_pushed(Params p) {
    int e = blockIdx.x;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    {   // the rock grid is rebuilt from the new centres
        const int gcells = ((p.W + 31) >> kGridShift) * ((p.H + 31) >> kGridShift);
        unsigned long long *g = p.rock_grid + (int64_t)e * gcells;
        for (int k = threadIdx.x; k < gcells; k += blockDim.x) g[k] = 0ull;
        __syncthreads();
    }
    const double *xs = p.x + (int64_t)e * p.N, *ys = p.y + (int64_t)e * p.N;
    for (int r = warp; r < p.R; r += nwarp) {
        double *c = p.rock_c + ((int64_t)e * p.R + r) * 2;
        double cx = c[0], cy = c[1], rad = p.rock_rad[(int64_t)e * p.R + r];
        double sx = 0.0, sy = 0.0;
        for (int a0 = 0; a0 < p.N; a0 += 32) {
            int a = a0 + lane;
            double px = 0.0, py = 0.0;
            bool hit = false;
            if (a < p.N) {
                double vx = cx - xs[a], vy = cy - ys[a];
                double d = sqrt(vx * vx + vy * vy);
                if (!(d > rad)) {
                    double fac = 1.0 - rad / (d + 0.001);
                    px = vx * fac; py = vy * fac; hit = true;
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                int l = __ffs(m) - 1;
                m &= m - 1;
                sx += __shfl_sync(0xffffffffu, px, l);
                sy += __shfl_sync(0xffffffffu, py, l);
            }
        }
        __syncwarp();
        if (lane == 0) {
            double wt = p.rock_w[(int64_t)e * p.R + r];
            double nx = cx - sx / wt, ny = cy - sy / wt;
            c[0] = nx;
            c[1] = ny;
            rock_grid_mark(p, e, r, nx, ny, rad);
        }
    }
}

// CircleObstacles.update, second half (circle_obstacles.py:53-58) + the ant part of Ants.update.
__global__ void __launch_bounds__(256) k_rocks_push_ants(Params p, uint32_t owner_stamp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.EN) return;
    int e = (int)(i / p.N);
    int a = (int)(i - (int64_t)e * p.N);
    double x = p.x[i], y = p.y[i];
    const double *rc = p.rock_c + (int64_t)e * p.R * 2;
    const double *rr = p.rock_rad + (int64_t)e * p.R;
    double sx = 0.0, sy = 0.0;
    for (int r = 0; r < p.R; ++r) {
        double vx = rc[2 * r] - x, vy = rc[2 * r + 1] - y;
        double d = sqrt(vx * vx + vy * vy);
        double rad = rr[r];
        if (!(d > rad)) {
            double fac = 1.0 - rad / (d + 0.001);
            sx += vx * fac; sy += vy * fac;
        }
    }
    x = pymod(x + sx, (double)p.W);                                            // translate_ants -> warp_xy
    y = pymod(y + sy, (double)p.H);
    p.x[i] = x; p.y[i] = y;
    p.prev_x[i] = x; p.prev_y[i] = y; p.prev_theta[i] = p.theta[i];
    int64_t eoff = (int64_t)e * p.plane;
    atomicMax(p.owner + eoff + (int64_t)cell_of(x, p.W) * p.Hp + cell_of(y, p.H), owner_stamp | (uint32_t)a);
    p.reward_state[i] = (uint8_t)((double)p.reward_state[i] * 0.9);
}

// ------------------------------------------------------------------------------------------------ pheromone field
// Dense pass for DIFFUSE_FACTOR == 0: walls.py:30 (zero inside walls), pheromone.py:44 reduced to its centre tap
// (bit-identical to convolve2d with an all-zero ring), pheromone.py:45 threshold, and the whole-plane clamp of
// pheromone.py:41.  Pure streaming: 16 B read + 16 B written per pair of cells, plus the wall mask.
__global__ void __launch_bounds__(256) k_evaporate_dense(Params p, int chunks) {
    const int64_t plane2 = p.plane >> 1;                                       // Hp is even
    const int64_t ep = blockIdx.x / chunks;                                    // e * P + p
    const int64_t chunk = blockIdx.x - ep * chunks;
    const int e = (int)(ep / p.P);
    double2 *ph = reinterpret_cast<double2 *>(p.phero + ep * p.plane);
    const uchar2 *wl = reinterpret_cast<const uchar2 *>(p.walls + (int64_t)e * p.plane);
    const double c = p.filt_center, mx = p.phero_max_val;
    const bool clamp = p.has_max_val && p.N > 0;
    for (int64_t j = chunk * blockDim.x + threadIdx.x; j < plane2; j += (int64_t)chunks * blockDim.x) {
        double2 v = ph[j];
        if (v.x == 0.0 && v.y == 0.0) continue;
        uchar2 wv = wl[j];
        double a = wv.x ? 0.0 : v.x * c;
        double b = wv.y ? 0.0 : v.y * c;
        a = a < 0.01 ? 0.0 : a;
        b = b < 0.01 ? 0.0 : b;
        if (clamp) { a = fmin(a, mx); b = fmin(b, mx); }
        ph[j] = make_double2(a, b);
    }
}

// Active-tile variant of the same pass: one warp per 16x16 tile, tiles without pheromone are skipped by their
// activity byte; a tile whose cells all decayed to zero is retired.
__global__ void __launch_bounds__(256) k_evaporate_tiles(Params p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t tiles_per_plane = (int64_t)p.tiles_x * p.tiles_y;
    const int64_t ntiles = (int64_t)p.E * p.P * tiles_per_plane;
    const double c = p.filt_center, mx = p.phero_max_val;
    const bool clamp = p.has_max_val && p.N > 0;
    unsigned long long processed = 0;
    // each warp scans 32 activity bytes at a time, interleaved across warps for balance
    for (int64_t g = warp_g * 32; g < ntiles; g += nwarps * 32) {
        int64_t t = g + lane;
        unsigned act = __ballot_sync(0xffffffffu, t < ntiles && p.tile_active[t] != 0);
        while (act) {
            int l = __ffs(act) - 1;
            act &= act - 1;
            int64_t tt = g + l;
            int64_t ep = tt / tiles_per_plane;
            int64_t tin = tt - ep * tiles_per_plane;
            int tx = (int)(tin / p.tiles_y), ty = (int)(tin - (int64_t)tx * p.tiles_y);
            int e = (int)(ep / p.P);
            // 16 rows x 16 cells: lane -> row = lane / 2 (+ 0), half = lane & 1 (8 cells = 4 double2)
            int row = tx * kTile + (lane >> 1);
            int col = ty * kTile + (lane & 1) * 8;
            bool any = false;
            if (row < p.W && col < p.Hp) {
                int64_t off = (int64_t)row * p.Hp + col;
                double2 *ph = reinterpret_cast<double2 *>(p.phero + ep * p.plane + off);
                const uchar2 *wl = reinterpret_cast<const uchar2 *>(p.walls + (int64_t)e * p.plane + off);
                double2 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = ph[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (v[k].x == 0.0 && v[k].y == 0.0) continue;
                    uchar2 wv = wl[k];
                    double a = wv.x ? 0.0 : v[k].x * c;
                    double b = wv.y ? 0.0 : v[k].y * c;
                    a = a < 0.01 ? 0.0 : a;
                    b = b < 0.01 ? 0.0 : b;
                    if (clamp) { a = fmin(a, mx); b = fmin(b, mx); }
                    ph[k] = make_double2(a, b);
                    any |= (a != 0.0) || (b != 0.0);
                }
            }
            if (!__any_sync(0xffffffffu, any) && lane == 0) p.tile_active[tt] = 0;
            ++processed;
        }
    }
    if (lane == 0 && processed) atomicAdd(p.tile_counter, processed);
}

// Diffusion stencil for DIFFUSE_FACTOR != 0 (pheromone.py:44, 3x3 zero-filled convolution) with the wall
// zeroing of walls.py:30 applied to the inputs, threshold and clamp; reads `phero`, writes `phero_alt`.
// Tile of 32 x 32 outputs per block, halo staged in shared memory.
constexpr int kStX = 32, kStY = 32;
__global__ void __launch_bounds__(256) k_diffuse_stencil(Params p, int nbx, int nby) {
    __shared__ double tile[kStX + 2][kStY + 2 + 1];
    const int64_t ep = blockIdx.x / (nbx * nby);
    const int brem = (int)(blockIdx.x - ep * (nbx * nby));
    const int e = (int)(ep / p.P);
    const int x0 = (brem / nby) * kStX, y0 = (brem % nby) * kStY;
    const double *src = p.phero + ep * p.plane;
    const uint8_t *wl = p.walls + (int64_t)e * p.plane;
    for (int t = threadIdx.x; t < (kStX + 2) * (kStY + 2); t += blockDim.x) {
        int lx = t / (kStY + 2), ly = t - lx * (kStY + 2);
        int gx = x0 + lx - 1, gy = y0 + ly - 1;
        double v = 0.0;
        if (gx >= 0 && gx < p.W && gy >= 0 && gy < p.H) {
            int64_t off = (int64_t)gx * p.Hp + gy;
            v = wl[off] ? 0.0 : src[off];
        }
        tile[lx][ly] = v;
    }
    __syncthreads();
    const double fc = p.filt_center, fr = p.filt_ring, mx = p.phero_max_val;
    const bool clamp = p.has_max_val && p.N > 0;
    double *dst = p.phero_alt + ep * p.plane;
    for (int t = threadIdx.x; t < kStX * kStY; t += blockDim.x) {
        int lx = t / kStY, ly = t - lx * kStY;
        int gx = x0 + lx, gy = y0 + ly;
        if (gx >= p.W || gy >= p.H) continue;
        double acc = 0.0;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
                acc += tile[lx + dx][ly + dy] * ((dx == 1 && dy == 1) ? fc : fr);
        acc = acc < 0.01 ? 0.0 : acc;
        if (clamp) acc = fmin(acc, mx);
        dst[(int64_t)gx * p.Hp + gy] = acc;
    }
}

// Ants.emit_pheromones -> Pheromone.add_pheromones (ants.py:98-100, pheromone.py:36-41): the owner of each cell
// adds its activation and clamps to max_val.
__global__ void __launch_bounds__(256) k_deposit_commit(Params p, uint32_t owner_stamp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.EN) return;
    int e = (int)(i / p.N);
    int a = (int)(i - (int64_t)e * p.N);
    int cx = cell_of(p.x[i], p.W), cy = cell_of(p.y[i], p.H);
    int64_t cell = (int64_t)cx * p.Hp + cy;
    if (p.owner[(int64_t)e * p.plane + cell] != (owner_stamp | (uint32_t)a)) return;
    for (int k = 0; k < p.P; ++k) {
        double av = p.act[(int64_t)k * p.EN + i];
        if (av == 0.0) continue;
        double *ph = p.phero + ((int64_t)e * p.P + k) * p.plane + cell;
        double v = *ph + av;
        if (p.has_max_val) v = fmin(v, p.phero_max_val);
        *ph = v;
        if (p.tile_active != nullptr)
            p.tile_active[((int64_t)e * p.P + k) * p.tiles_x * p.tiles_y + (int64_t)(cx / kTile) * p.tiles_y + cy / kTile] = 1;
    }
}

// Anthill.update (anthill.py:41-46) for the cells queued by k_food_commit.
__global__ void __launch_bounds__(256) k_absorb_list(Params p) {
    uint32_t n = *p.absorb_count;
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        uint32_t e = p.absorb_list[2 * k], c = p.absorb_list[2 * k + 1];
        unsigned long long *fp = reinterpret_cast<unsigned long long *>(p.food + (int64_t)e * p.plane + c);
        double v = __longlong_as_double((long long)atomicExch(fp, 0ull));      // qte -= qte * area
        if (v != 0.0) atomicAdd(p.hill_food + e, v);
    }
    __syncthreads();
    if (threadIdx.x == 0) *p.absorb_count = 0u;
}

// Anthill.update as a sweep over the disc's bounding box (first update after an import, quirk Q10).
__global__ void __launch_bounds__(256) k_absorb_sweep(Params p) {
    __shared__ double red[256];
    int e = blockIdx.x;
    const int32_t *hl = p.hill + 4 * e;
    int r = hl[2];
    int x0 = max(hl[0] - r, 0), x1 = min(hl[0] + r, p.W - 1);
    int y0 = max(hl[1] - r, 0), y1 = min(hl[1] + r, p.H - 1);
    int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    double acc = 0.0;
    if (bw > 0 && bh > 0) {
        for (int t = threadIdx.x; t < bw * bh; t += blockDim.x) {
            int cx = x0 + t / bh, cy = y0 + t % bh;
            if (!in_hill(hl, cx, cy)) continue;
            double *fp = p.food + (int64_t)e * p.plane + (int64_t)cx * p.Hp + cy;
            double v = *fp;
            if (v != 0.0) { acc += v; *fp = v - v; }
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && red[0] != 0.0) p.hill_food[e] += red[0];
}

// ------------------------------------------------------------------------------------------------ import / export helpers
__global__ void k_meta_from_explored(Params p, const uint8_t *__restrict__ explored_dense) {
    int64_t n = (int64_t)p.E * p.W * p.H;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t ex = j / p.H;
        int y = (int)(j - ex * p.H);
        p.meta[ex * p.Hp + y] = explored_dense[j] ? 0xFFFFu : 0u;
    }
}
__global__ void k_explored_from_meta(Params p, uint8_t *__restrict__ explored_dense) {
    int64_t n = (int64_t)p.E * p.W * p.H;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t ex = j / p.H;
        int y = (int)(j - ex * p.H);
        explored_dense[j] = (p.meta[ex * p.Hp + y] & 0xFFFFu) ? 1 : 0;
    }
}
// generation counters are 16 bit: before one wraps, fold every live stamp into the "long ago" value
__global__ void k_meta_renormalize(Params p, int fold_explored, int clear_occ) {
    int64_t n = (int64_t)p.E * p.plane;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        uint32_t m = p.meta[j];
        uint32_t lo = m & 0xFFFFu, hi = m >> 16;
        if (fold_explored && lo) lo = 0xFFFFu;
        if (clear_occ) hi = 0;
        p.meta[j] = (hi << 16) | lo;
    }
}
__global__ void k_rock_grid_build(Params p) {          // one block per env (after import)
    const int e = blockIdx.x;
    const int gcells = ((p.W + 31) >> kGridShift) * ((p.H + 31) >> kGridShift);
    unsigned long long *g = p.rock_grid + (int64_t)e * gcells;
    for (int k = threadIdx.x; k < gcells; k += blockDim.x) g[k] = 0ull;
    __syncthreads();
    for (int r = threadIdx.x; r < p.R; r += blockDim.x)
        rock_grid_mark(p, e, r, p.rock_c[((int64_t)e * p.R + r) * 2], p.rock_c[((int64_t)e * p.R + r) * 2 + 1],
                       p.rock_rad[(int64_t)e * p.R + r]);
}
__global__ void k_tiles_from_phero(Params p) {
    // mark every tile that holds a non-zero pheromone cell (after import)
    const int64_t tiles_per_plane = (int64_t)p.tiles_x * p.tiles_y;
    const int64_t ntiles = (int64_t)p.E * p.P * tiles_per_plane;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t ep = t / tiles_per_plane;
        int64_t tin = t - ep * tiles_per_plane;
        int tx = (int)(tin / p.tiles_y), ty = (int)(tin - (int64_t)tx * p.tiles_y);
        bool any = false;
        for (int dx = 0; dx < kTile && !any; ++dx) {
            int row = tx * kTile + dx;
            if (row >= p.W) break;
            const double *ph = p.phero + ep * p.plane + (int64_t)row * p.Hp + ty * kTile;
            for (int dy = 0; dy < kTile; ++dy)
                if (ty * kTile + dy < p.Hp && ph[dy] != 0.0) { any = true; break; }
        }
        p.tile_active[t] = any ? 1 : 0;
    }
}

}  // namespace ants
