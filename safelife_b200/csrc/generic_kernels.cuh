// generic_kernels.cuh -- CTA-per-board kernels for any board size H, W >= 3.
//
// These are the general path: one thread block per board, the board staged in
// shared memory, each thread owning a contiguous raster-order run of cells so
// that the reference's raster-order RNG draws (advance_board.c:115) can be
// reproduced with a block prefix sum of "qualifying cell" counts and a PCG64
// jump-ahead per thread.  The warp-per-env bit-plane kernels in
// warp_kernels.cuh are the fast path for boards up to 32x32.
#pragma once
#include "cells.cuh"
#include "../../include/safelife_b200.h"

namespace sl {

constexpr int GENERIC_THREADS = 128;

// Exclusive prefix sum of one int per thread across the block.
// scratch: int[33] shared.  Returns this thread's prefix; *total = block sum.
__device__ inline int block_exclusive_scan(int v, int *scratch, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int w = lane < nw ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < nw) scratch[lane] = winc - w;
        if (lane == 31) scratch[32] = winc;
    }
    __syncthreads();
    int res = scratch[warp] + inc - v;
    *total = scratch[32];
    __syncthreads();
    return res;
}

__device__ inline int block_sum(int v, int *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5, s = 0;
    for (int w = 0; w < nw; w++) s += scratch[w];
    __syncthreads();
    return s;
}

// One stencil step src -> dst (both shared, `cells` uint16 each) for the whole
// block.  rng_sh: uint64[4] in shared memory, advanced by the number of draws
// (thr == 0 or rng_sh == nullptr: every draw fails and no state is written).
// Ends with a __syncthreads().
__device__ inline void block_advance(const uint16_t *src, uint16_t *dst, int H, int W,
                                     uint64_t thr, uint64_t *rng_sh, int *scratch) {
    const int cells = H * W;
    const int chunk = (cells + blockDim.x - 1) / blockDim.x;
    const int begin = min(cells, (int)threadIdx.x * chunk);
    const int end = min(cells, begin + chunk);
    auto at = [src](int i) -> uint32_t { return src[i]; };
    int nq = 0;
    for (int i = begin; i < end; i++) {
        int y = i / W, x = i - y * W;
        Hood h = gather_hood(at, y, x, H, W);
        bool q;
        uint32_t r = apply_rule(src[i], h, &q);
        // a qualifying cell is dead and not frozen, so bit 15 (PULLABLE) of the
        // would-be spawned word is free to tag it until the draw is resolved
        dst[i] = (uint16_t)(q ? (r | 0x8000u) : r);
        nq += q;
    }
    int total;
    int prefix = block_exclusive_scan(nq, scratch, &total);
    if (total > 0) {
        if (nq > 0) {
            Pcg g;
            bool have = rng_sh != nullptr;
            if (have) {
                g = {{rng_sh[0], rng_sh[1]}, {rng_sh[2], rng_sh[3]}};
                g.state = pcg_jump(g.state, g.inc, (uint32_t)prefix);
            }
            for (int i = begin; i < end; i++) {
                uint32_t r = dst[i];
                if (!(src[i] & ALIVE) && (r & 0x8000u) && (r & ALIVE)) {
                    // tagged: a dead cell only comes out ALIVE with bit 15 set
                    // through the tag above (births and spawns never carry it)
                    bool ok = have && pcg_next_below(g, thr);
                    dst[i] = ok ? (uint16_t)(r & 0x7fffu) : src[i];
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && rng_sh != nullptr) {
            u128 s = pcg_jump({rng_sh[0], rng_sh[1]}, {rng_sh[2], rng_sh[3]}, (uint32_t)total);
            rng_sh[0] = s.hi;
            rng_sh[1] = s.lo;
        }
    }
    __syncthreads();
}

__host__ __device__ inline size_t generic_smem_bytes(int cells) {
    // three uint16 planes (16-byte aligned each) + scratch
    size_t plane = ((size_t)cells * 2 + 15) & ~(size_t)15;
    return 3 * plane + 64 * sizeof(int);
}

struct GenericSmem {
    uint16_t *a, *b, *c;
    int *scratch;       // int[33]
    uint64_t *rng;      // uint64[4]
    int *misc;          // int[8]
    __device__ GenericSmem(unsigned char *base, int cells) {
        size_t plane = ((size_t)cells * 2 + 15) & ~(size_t)15;
        a = (uint16_t *)base;
        b = (uint16_t *)(base + plane);
        c = (uint16_t *)(base + 2 * plane);
        scratch = (int *)(base + 3 * plane);
        rng = (uint64_t *)(scratch + 40);
        misc = scratch + 48;
    }
};

// ------------------------------------------------------------------------
// n-step advance, optionally accumulating life occupancy.
// counts == nullptr : out[b] = board after n_steps (module.c:20-49).
// counts != nullptr : counts[b][cell][colour] += occupancy after each step
//                     (module.c:52-81); `out` unused.
// ------------------------------------------------------------------------
__global__ void __launch_bounds__(GENERIC_THREADS)
k_advance_generic(const uint16_t *in, uint16_t *out,   // may alias
                  int32_t *__restrict__ counts, int n, int H, int W, uint64_t thr,
                  int n_steps, uint64_t *__restrict__ rng) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cells = H * W;
    GenericSmem sm(smem_raw, cells);
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const uint16_t *src = in + (size_t)b * cells;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) sm.a[i] = src[i];
        if (rng != nullptr && threadIdx.x < 4) sm.rng[threadIdx.x] = rng[4 * (size_t)b + threadIdx.x];
        __syncthreads();
        uint16_t *cur = sm.a, *nxt = sm.b;
        for (int s = 0; s < n_steps; s++) {
            block_advance(cur, nxt, H, W, thr, rng ? sm.rng : nullptr, sm.scratch);
            if (counts != nullptr) {
                int32_t *cnt = counts + (size_t)b * cells * 8;
                for (int i = threadIdx.x; i < cells; i += blockDim.x) {
                    uint32_t c = nxt[i];
                    if ((c & ALIVE) && !(c & (AGENT | EXIT | FROZEN)))
                        cnt[8 * i + ((c & COLORS) >> COLOR_SHIFT)] += 1;
                }
            }
            uint16_t *t = cur; cur = nxt; nxt = t;
        }
        if (counts == nullptr) {
            uint16_t *dst = out + (size_t)b * cells;
            for (int i = threadIdx.x; i < cells; i += blockDim.x) dst[i] = cur[i];
        }
        if (rng != nullptr && threadIdx.x < 2) rng[4 * (size_t)b + threadIdx.x] = sm.rng[threadIdx.x];
        __syncthreads();
    }
}

// Device part of side_effects.side_effect_score (side_effects.py:101-111), one
// CTA per board, any size: inaction board = init advanced num_steps[b] times;
// then num_samples more steps of it and of the final board with occupancy
// counted into counts[b][0] / counts[b][1]; all draws from rng[b] in that order.
__global__ void __launch_bounds__(GENERIC_THREADS)
k_side_effect_generic(const uint16_t *__restrict__ init_board,
                      const uint16_t *__restrict__ final_board, int32_t *__restrict__ counts,
                      int n, int H, int W, const uint64_t *__restrict__ thr,
                      const int32_t *__restrict__ num_steps, int num_samples,
                      uint64_t *__restrict__ rng) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cells = H * W;
    GenericSmem sm(smem_raw, cells);
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        if (rng != nullptr && threadIdx.x < 4) sm.rng[threadIdx.x] = rng[4 * (size_t)b + threadIdx.x];
        const uint64_t th = thr[b];
        for (int run = 0; run < 2; run++) {
            const uint16_t *src = (run == 0 ? init_board : final_board) + (size_t)b * cells;
            __syncthreads();
            for (int i = threadIdx.x; i < cells; i += blockDim.x) sm.a[i] = src[i];
            __syncthreads();
            uint16_t *cur = sm.a, *nxt = sm.b;
            const int lead = run == 0 ? num_steps[b] : 0;
            int32_t *cnt = counts + ((size_t)b * 2 + run) * cells * 8;
            for (int s = 0; s < lead + num_samples; s++) {
                block_advance(cur, nxt, H, W, th, rng ? sm.rng : nullptr, sm.scratch);
                if (s >= lead) {
                    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
                        uint32_t c = nxt[i];
                        if ((c & ALIVE) && !(c & (AGENT | EXIT | FROZEN)))
                            cnt[8 * i + ((c & COLORS) >> COLOR_SHIFT)] += 1;
                    }
                }
                uint16_t *t = cur; cur = nxt; nxt = t;
            }
        }
        __syncthreads();
        if (rng != nullptr && threadIdx.x < 2) rng[4 * (size_t)b + threadIdx.x] = sm.rng[threadIdx.x];
        __syncthreads();
    }
}

// out[b][72] histogram (advance_board.c:192-207).
__global__ void __launch_bounds__(GENERIC_THREADS)
k_alive_counts(const uint16_t *__restrict__ board, const uint16_t *__restrict__ goals,
               int64_t *__restrict__ out, int n, int cells) {
    __shared__ int hist[72];
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        if (threadIdx.x < 72) hist[threadIdx.x] = 0;
        __syncthreads();
        const uint16_t *bd = board + (size_t)b * cells, *gl = goals + (size_t)b * cells;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) {
            int bin = count_bin(bd[i], gl[i]);
            if (bin >= 0) atomicAdd(&hist[bin], 1);
        }
        __syncthreads();
        if (threadIdx.x < 72) out[(size_t)b * 72 + threadIdx.x] = hist[threadIdx.x];
        __syncthreads();
    }
}

// One thread per board; agents act in index order (advance_board.c:220).
__global__ void k_execute_actions(uint16_t *__restrict__ board, int64_t *__restrict__ locations,
                                  const int64_t *__restrict__ actions, int n, int H, int W,
                                  int n_agents, int action_stride) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    uint16_t *bd = board + (size_t)b * H * W;
    int64_t *loc = locations + (size_t)b * n_agents * 2;
    const int64_t *act = actions + (size_t)b * (action_stride ? n_agents : 1);
    for (int k = 0; k < n_agents; k++) {
        int64_t a = act[action_stride ? k : 0];
        if (a == 0) continue;
        int y0 = (int)(loc[2 * k] % H), x0 = (int)(loc[2 * k + 1] % W);
        int y = y0, x = x0;
        // action >= 5 toggles, direction (a-1)&3: fold large values into that form
        int act_small = (a >= 5) ? 5 + (int)((a - 1) & 3) : 1 + (int)((a - 1) & 3);
        agent_action(bd, H, W, y, x, act_small);
        if (y != y0 || x != x0) {
            loc[2 * k] = y;
            loc[2 * k + 1] = x;
        }
    }
}

// Sequential emulation of wrapped_label.c:5-55 (incl. the iset visiting order
// of iset.c:6-61), one thread per grid; work arrays in global scratch.
__global__ void k_wrapped_label(int32_t *__restrict__ data, int32_t *__restrict__ num_labels,
                                int32_t *__restrict__ scratch, int n, int H, int W) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int size = H * W;
    int32_t *d = data + (size_t)b * size;
    int32_t *order = scratch + (size_t)b * size * 3, *pos = order + size, *stack = pos + size;
    for (int i = 0; i < size; i++) order[i] = pos[i] = i;
    int nvis = 0, label = 0;
    while (nvis < size) {
        int seed = order[nvis++];
        if (!d[seed]) continue;
        d[seed] = ++label;
        int top = 0;
        stack[top++] = seed;
        while (top > 0) {
            int cur = stack[--top];
            int r = cur / W, c = cur - r * W;
            for (int dy = -1; dy <= 1; dy++) {
                int y = wrap_idx(r + dy, H);
                for (int dx = -1; dx <= 1; dx++) {
                    int x = wrap_idx(c + dx, W);
                    int j = y * W + x;
                    bool seen = pos[j] < nvis;
                    if (d[j] > 0 && !seen) {
                        d[j] = label;
                        stack[top++] = j;
                    }
                    if (!seen) {
                        int k = pos[j], other = order[nvis];
                        order[nvis] = j; pos[j] = nvis;
                        order[k] = other; pos[other] = k;
                        nvis++;
                    }
                }
            }
        }
    }
    num_labels[b] = label;
}

// Points contribution of one cell (alive_counts bin x points table).
// (table may live in global or shared memory).  dead_counts = false skips the
// lookup for dead cells when the table's "empty" column is all zero.
__device__ inline int cell_points(uint32_t b, uint32_t g, const int32_t *table,
                                  bool dead_counts = true) {
    if (!dead_counts && !(b & ALIVE)) return 0;
    int bin = count_bin(b, g);
    return bin < 0 ? 0 : table[bin];
}

// ------------------------------------------------------------------------
// Exit / points logic shared by step and reset (single thread per env).
// GameWithGoals.can_exit + update_exit_colors (safelife_game.py:537-552,
// 716-719) on the board in `bd` (shared or global); returns current_points
// (:684-687) AFTER the update.  count_points = sum(table * alive_counts) of
// the board as it is on entry; *delta_mem is what the exit-cell overwrite adds
// to that sum (0 unless something other than an exit sits on an exit cell).
// ------------------------------------------------------------------------
template <typename T>
__device__ inline int exit_update(T *bd, const T *gl, int W, int row, int col, int count_points,
                                  const SlEnvState &st, const int32_t *exits,
                                  const int32_t *table,
                                  bool *agent_alive, bool *exited, int *delta_mem) {
    const int init_points = st.init_points, required = st.required;
    bool can_exit = false;
    int idx = -1;
    if (row >= 0) {
        idx = row * W + col;
        uint32_t a = bd[idx];
        int earned = count_points - init_points + (((a & (AGENT | EXIT)) == EXIT) ? 1 : 0);
        earned = max(earned, 0);
        can_exit = (a & AGENT) && earned >= required;
        bd[idx] = (T)((a & ~(uint32_t)EXIT) | (can_exit ? (uint32_t)EXIT : 0u));
    }
    const uint32_t exit_type = LEVEL_EXIT | (can_exit ? (uint32_t)COLOR_R : 0u);
    int delta = 0;
    for (int k = 0; k < st.n_exits; k++) {
        // the first two exit cells travel in the env record, the rest stay in the pool
        int e = k == 0 ? st.exit0 : (k == 1 ? st.exit1 : exits[k]);
        uint32_t old = bd[e];
        if (old != exit_type) {
            if ((old & ~(uint32_t)COLORS) != LEVEL_EXIT) delta -= cell_points(old, gl[e], table);
            bd[e] = (T)exit_type;
        }
    }
    *delta_mem = delta;
    *agent_alive = false;
    *exited = false;
    if (idx < 0) return count_points;
    uint32_t a = bd[idx];
    *agent_alive = (a & AGENT) != 0;
    *exited = (a & (AGENT | EXIT)) == EXIT;
    return count_points + (*exited ? 1 : 0);
}

// Scalar tail of SafeLifeEnv.step (safelife_env.py:155-175) for one env.
struct StepOut {           // what finish_step wrote for the env, for an epilogue in the same kernel
    float reward;
    bool done, times_up;
};

__device__ inline bool finish_step(const SlEnvBatch &B, int e, SlEnvState &st, int row, int col,
                                   int value, bool agent_alive, bool exited, int static_goals,
                                   int count_points_mem, StepOut *out = nullptr,
                                   uint32_t host_tag = 0) {
    int steps = st.num_steps + 1;
    bool times_up = steps >= B.time_limit;
    bool active = st.active != 0;
    float reward;
    bool done, success;
    if (row >= 0) {
        reward = (float)((value - st.old_value) * (active ? 1 : 0));
        success = exited;
        done = !agent_alive || times_up;
    } else {   // no agent: safelife_env.py:163-166
        reward = 0.f; done = true; success = false;
    }
    st.num_steps = steps;
    st.old_value = value;
    // first step of an episode: totals restart (see SL_RESET_KEEP_STATS)
    st.ep_reward = (steps == 1 ? 0.f : st.ep_reward) + reward;
    st.ep_length = (steps == 1 ? 0 : st.ep_length) + (active ? 1 : 0);
    st.active = (active && !done) ? 1 : 0;
    st.static_goals = (uint8_t)static_goals;
    st.loc_row = row;
    st.loc_col = col;
    st.count_points = count_points_mem;
    B.reward[e] = reward;
    B.done[e] = done;
    B.success[e] = success;
    B.times_up[e] = times_up;
    if (B.host_out != nullptr)
        B.host_out[e] = (uint64_t)__float_as_uint(reward) | ((uint64_t)(done ? 1 : 0) << 32) |
                        ((uint64_t)(success ? 1 : 0) << 40) | ((uint64_t)(times_up ? 1 : 0) << 48) |
                        ((uint64_t)(host_tag & 0xffu) << 56);      // host sessions: the step that wrote it
    if (out != nullptr) *out = {reward, done, times_up};
    return done;
}

// sum(table * alive_counts) over a board/goals pair by the whole block.
__device__ inline int block_points(const uint16_t *bd, const uint16_t *gl, int cells,
                                   const int32_t *__restrict__ table, int *scratch) {
    int s = 0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        int bin = count_bin(bd[i], gl[i]);
        if (bin >= 0) s += __ldg(table + bin);
    }
    return block_sum(s, scratch);
}

// SafeLifeEnv.step, one CTA per env (safelife_env.py:148-201).
__global__ void __launch_bounds__(GENERIC_THREADS)
k_env_step_generic(SlEnvBatch B, const uint8_t *__restrict__ actions) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = B.H, W = B.W, cells = H * W;
    GenericSmem sm(smem_raw, cells);
    for (int e = blockIdx.x; e < B.n_envs; e += gridDim.x) {
        uint16_t *board = B.board + (size_t)e * cells;
        uint16_t *goals = B.goals + (size_t)e * cells;
        SlEnvState *stp = B.state + e;
        const int lvl = stp->level_id;
        const uint64_t thr = stp->spawn_thr;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) sm.a[i] = board[i];
        if (threadIdx.x < 4) sm.rng[threadIdx.x] = stp->rng[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            int row = stp->loc_row, col = stp->loc_col;
            if (row >= 0) agent_action(sm.a, H, W, row, col, (int)actions[e]);
            sm.misc[0] = row;
            sm.misc[1] = col;
        }
        __syncthreads();
        block_advance(sm.a, sm.b, H, W, thr, sm.rng, sm.scratch);   // board: a -> b
        // goals (safelife_game.py:753-761): advanced from the same stream
        // after the board unless known static.
        int sg = stp->static_goals;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) sm.a[i] = goals[i];
        __syncthreads();
        const uint16_t *gl = sm.a;
        if (sg != 1) {
            block_advance(sm.a, sm.c, H, W, thr, sm.rng, sm.scratch);   // goals: a -> c
            if (sg == 0) {
                int diff = 0;
                for (int i = threadIdx.x; i < cells; i += blockDim.x)
                    diff |= ((sm.c[i] & SPAWNING) || (sm.c[i] != sm.a[i])) ? 1 : 0;
                diff = block_sum(diff, sm.scratch);
                sg = diff ? 2 : 1;
            }
            for (int i = threadIdx.x; i < cells; i += blockDim.x) goals[i] = sm.c[i];
            gl = sm.c;
        }
        const int32_t *table = B.pool_table + (size_t)lvl * 72;
        int pts = block_points(sm.b, gl, cells, table, sm.scratch);
        if (threadIdx.x == 0) {
            bool agent_alive, exited;
            int row = sm.misc[0], col = sm.misc[1], delta;
            SlEnvState st = *stp;
            int value = exit_update(sm.b, gl, W, row, col, pts, st,
                                    B.pool_exits + (size_t)lvl * B.max_exits, table,
                                    &agent_alive, &exited, &delta);
            st.rng[0] = sm.rng[0];
            st.rng[1] = sm.rng[1];
            finish_step(B, e, st, row, col, value, agent_alive, exited, sg, pts + delta);
            *stp = st;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += blockDim.x) board[i] = sm.b[i];
        __syncthreads();
    }
}

// SafeLifeEnv.reset for masked envs (safelife_env.py:203-218), CTA per env.
__global__ void __launch_bounds__(GENERIC_THREADS)
k_env_reset(SlEnvBatch B, const uint8_t *__restrict__ mask, const int32_t *__restrict__ new_level,
            int flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = B.H, W = B.W, cells = H * W;
    GenericSmem sm(smem_raw, cells);
    for (int e = blockIdx.x; e < B.n_envs; e += gridDim.x) {
        if (mask != nullptr && !mask[e]) continue;
        SlEnvState *stp = B.state + e;
        int lvl = new_level != nullptr ? new_level[e] : stp->level_id;
        if (new_level == nullptr && (flags & SL_RESET_NEXT_LEVEL))
            lvl = (int)(((long long)lvl + B.level_stride) % B.n_levels);
        const uint16_t *pb = B.pool_board + (size_t)lvl * cells;
        const uint16_t *pg = B.pool_goals + (size_t)lvl * cells;
        uint16_t *board = B.board + (size_t)e * cells;
        uint16_t *goals = B.goals + (size_t)e * cells;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) {
            sm.a[i] = pb[i];
            uint16_t g = pg[i];
            sm.b[i] = g;
            goals[i] = g;
        }
        __syncthreads();
        const int32_t *table = B.pool_table + (size_t)lvl * 72;
        int pts = block_points(sm.a, sm.b, cells, table, sm.scratch);
        if (threadIdx.x == 0) {
            bool agent_alive, exited;
            int row = B.pool_loc[2 * lvl], col = B.pool_loc[2 * lvl + 1], delta;
            SlEnvState st = *stp;
            st.init_points = B.pool_init_points[lvl];
            st.required = B.pool_required[lvl];
            st.spawn_thr = B.pool_spawn_thr[lvl];
            bool empty = false;
            for (int g = 0; g < 8; g++) empty |= table[g * 9 + 8] != 0;
            st.empty_points = empty;
            const int32_t *exits = B.pool_exits + (size_t)lvl * B.max_exits;
            int n_exits = 0;
            while (n_exits < B.max_exits && exits[n_exits] >= 0) n_exits++;
            st.n_exits = n_exits;
            st.exit0 = n_exits > 0 ? exits[0] : -1;
            st.exit1 = n_exits > 1 ? exits[1] : -1;
            int value = exit_update(sm.a, sm.b, W, row, col, pts, st, exits, table,
                                    &agent_alive, &exited, &delta);
            st.level_id = lvl;
            st.loc_row = row;
            st.loc_col = col;
            st.num_steps = 0;
            st.old_value = value;
            st.count_points = pts + delta;
            if (!(flags & SL_RESET_KEEP_STATS)) {
                st.ep_length = 0;
                st.ep_reward = 0.f;
            }
            st.active = 1;
            st.static_goals = 0;
            *stp = st;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += blockDim.x) board[i] = sm.a[i];
        __syncthreads();
    }
}

// SimpleSideEffectPenalty cell count (env_wrappers.py:177-213).
__global__ void __launch_bounds__(GENERIC_THREADS)
k_side_effect_cells(SlEnvBatch B, const uint16_t *__restrict__ baseline, int32_t *__restrict__ out) {
    __shared__ int scratch[40];
    const int cells = B.H * B.W;
    const uint32_t keep = 0xffffu & ~(AGENT | DESTRUCTIBLE | FROZEN | PRESERVING | INHIBITING);
    for (int e = blockIdx.x; e < B.n_envs; e += gridDim.x) {
        const uint16_t *bd = B.board + (size_t)e * cells, *bl = baseline + (size_t)e * cells;
        const int32_t *exits = B.pool_exits + (size_t)B.state[e].level_id * B.max_exits;
        int s = 0;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) {
            bool is_exit = false;
            for (int k = 0; k < B.max_exits; k++) is_exit |= (exits[k] == i);
            if (!is_exit && ((bd[i] ^ bl[i]) & keep)) s++;
        }
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) out[e] = s;
    }
}

}  // namespace sl
