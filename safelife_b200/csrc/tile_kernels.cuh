// tile_kernels.cuh -- boards beyond 32x32: several warps per env.
//
// A board is cut into TR x TC tiles of at most 30 x 30 interior cells; one warp
// owns one tile.  Lane l of the warp holds board row r0 - 1 + l as ONE 32-bit
// word whose bit j is column c0 - 1 + j: lanes 1..hr / bits 1..wc are the tile's
// interior, lane 0 / lane hr + 1 and bit 0 / bit wc + 1 a halo ring read (with
// the torus wrap) from the neighbouring tiles' cells in the shared-memory slab.
// Inside a tile the 3x3 aggregate of advance_board.c:34-125 needs no wrap logic
// at all -- horizontal neighbours are plain shifts, vertical ones come from the
// adjacent lanes -- and no exchange with other warps: every warp transposes its
// own halo.  Compared with one warp per env on 64-bit rows (two rows per lane)
// this halves the registers per thread, spreads the per-changed-cell loops of a
// busy board (navigation: dozens of changes per step) over four times as many
// lanes, and lets the warps of one env run concurrently.
//
// What the warps of an env do share goes through a few bytes of shared memory
// and named barriers (one barrier id per env of the CTA):
//   * the reference's raster-order spawn draws (advance_board.c:115): every tile
//     publishes its candidate count per row; a candidate's draw index is the
//     number of candidates in earlier rows, plus those of the same row in tiles
//     to its left, plus its rank in its own word.  Every warp then runs the
//     (cheap) PCG64 leapfrog over ALL draws of the board and keeps the outcomes
//     of its own candidates, so no outcome has to travel;
//   * the change of sum(points_table * alive_counts), added up with one shared
//     atomic per warp;
//   * "did the goals change" for GameWithGoals._static_goals.
// The scalar tail of the step, the wrapper epilogue and a reset are run by the
// env's first warp (the shared code of warp_kernels.cuh); the observation is
// written by all of them, rows of the view dealt out round-robin.
#pragma once
#include <cuda_runtime.h>
#include "cells.cuh"
#include "generic_kernels.cuh"
#include "obs.cuh"
#include "wrap.cuh"
#include "warp_kernels.cuh"

namespace sl {

constexpr int TILE_DIM = 30;            // interior rows / columns of one tile
constexpr int TILE_MAX_WARPS = 16;      // warps of one env (4 x 4 tiles: boards up to 120 x 120)
constexpr int TILE_CTA_WARPS = 8;       // target CTA size in warps

// Shape of the tiling, computed once on the host (the kernels never divide).
struct TileParams {
    int32_t TR, TC, TW;      // tile rows, tile columns, warps per env
    int32_t epc;             // envs per CTA
    int16_t row_begin[6], col_begin[6];     // first interior row / column of each tile row / column, + end
    uint8_t env_of_warp[TILE_MAX_WARPS], tile_of_warp[TILE_MAX_WARPS];
    uint8_t tr_of_tile[TILE_MAX_WARPS], tc_of_tile[TILE_MAX_WARPS];
};
typedef TileParams TileShape;

inline TileParams tile_shape(int H, int W) {
    TileParams s = {};
    s.TR = (H + TILE_DIM - 1) / TILE_DIM;
    s.TC = (W + TILE_DIM - 1) / TILE_DIM;
    s.TW = s.TR * s.TC;
    s.epc = s.TW >= TILE_CTA_WARPS ? 1 : TILE_CTA_WARPS / s.TW;
    if (s.TR > 4 || s.TC > 4) return s;          // not supported (see tile_kernel_supports)
    for (int i = 0; i <= s.TR; i++) s.row_begin[i] = (int16_t)(i * (H / s.TR) + (i < H % s.TR ? i : H % s.TR));
    for (int i = 0; i <= s.TC; i++) s.col_begin[i] = (int16_t)(i * (W / s.TC) + (i < W % s.TC ? i : W % s.TC));
    for (int w = 0; w < TILE_MAX_WARPS; w++) {
        s.env_of_warp[w] = (uint8_t)(w / s.TW);
        s.tile_of_warp[w] = (uint8_t)(w % s.TW);
        s.tr_of_tile[w] = (uint8_t)(w / s.TC);
        s.tc_of_tile[w] = (uint8_t)(w % s.TC);
    }
    return s;
}

inline bool tile_kernel_supports(int H, int W) {
    if (H < 3 || W < 3) return false;
    const int TR = (H + TILE_DIM - 1) / TILE_DIM, TC = (W + TILE_DIM - 1) / TILE_DIM;
    return TR <= 4 && TC <= 4;                    // at most 16 warps per env, boards up to 120 x 120
}

// Planes of one byte of 32 cells (see planes32) whose first and last cell come from
// elsewhere: the halo columns of a tile that wrap around the torus.
template <unsigned SEL>
__device__ __forceinline__ void planes32_ends(const uint16_t *cells, const uint16_t *first,
                                              const uint16_t *last, uint32_t (&x)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t a = pack16(k == 0 ? *first : cells[k], cells[k + 16]);
        const uint32_t b = pack16(cells[k + 8], k == 7 ? *last : cells[k + 24]);
        x[k] = __byte_perm(a, b, SEL);
    }
    transpose8(x);
}

// One tile.  The 32-cell window of a lane starts `boff` cells left of the tile's
// first interior column: boff = 1 (left halo in bit 0) unless only the RIGHT halo
// column wraps around the torus, in which case the window is right-aligned (right
// halo in bit 31).  Either way a wrapping halo cell sits at a fixed end of the
// window and is fetched by address (planes32_ends); only a board of a single tile
// column (both halos wrap) patches its right halo bit by bit.
struct Tile {
    typedef uint32_t P;
    int H, W, lane;
    int r0, hr, c0, wc, boff;
    int y;                     // board row this lane holds (wrapped; any valid row for idle lanes)
    uint32_t imask;            // interior bits of this lane's word
    bool seam_l, seam_r, patch_r;

    __device__ __forceinline__ Tile(int H_, int W_, const TileParams &tp, int tile, int lane_)
        : H(H_), W(W_), lane(lane_) {
        const int tr = tp.tr_of_tile[tile], tc = tp.tc_of_tile[tile];
        r0 = tp.row_begin[tr];
        hr = tp.row_begin[tr + 1] - r0;
        c0 = tp.col_begin[tc];
        wc = tp.col_begin[tc + 1] - c0;
        seam_l = c0 == 0;
        seam_r = c0 + wc == W;
        boff = (seam_r && !seam_l) ? 31 - wc : 1;
        patch_r = seam_r && seam_l;
        int yy = r0 - 1 + lane;
        if (lane > hr + 1) yy = r0;
        y = yy < 0 ? yy + H : (yy >= H ? yy - H : yy);
        imask = (lane >= 1 && lane <= hr) ? (((1u << wc) - 1u) << boff) : 0u;
    }
    __device__ __forceinline__ P mask() const { return imask; }
    __device__ __forceinline__ P west(P x) const { return x << 1; }
    __device__ __forceinline__ P east(P x) const { return x >> 1; }
    __device__ __forceinline__ P north(P x) const { return __shfl_up_sync(FULL, x, 1); }
    __device__ __forceinline__ P south(P x) const { return __shfl_down_sync(FULL, x, 1); }
    // flat index of the cell behind bit j of this lane's word
    __device__ __forceinline__ int cell_base() const { return y * W + c0 - boff; }

    template <unsigned SEL>
    __device__ __forceinline__ void load(const uint16_t *slab, uint32_t (&x)[8], uint32_t *vr) const {
        const uint16_t *row = slab + y * W, *win = row + c0 - boff;
        planes32_ends<SEL>(win, seam_l ? row + W - 1 : win, (seam_r && !seam_l) ? row : win + 31, x);
        *vr = patch_r ? (uint32_t)row[0] : 0u;
    }
    // right halo bit of the plane of cell bit `bit` (single tile column only)
    __device__ __forceinline__ uint32_t fix(uint32_t x, uint32_t vr, int bit) const {
        if (patch_r) x = (x & ~(1u << (wc + 1))) | (((vr >> bit) & 1u) << (wc + 1));
        return x;
    }
    __device__ __forceinline__ LowPlanesT<P> load_low(const uint16_t *slab) const {
        uint32_t x[8], vr;
        load<0x6240>(slab, x, &vr);
        return {fix(x[0], vr, 0), fix(x[3], vr, 3), x[4], fix(x[5], vr, 5), fix(x[6], vr, 6),
                fix(x[7], vr, 7)};                 // FROZEN matters for the own cell only
    }
    __device__ __forceinline__ HighPlanesT<P> load_high(const uint16_t *slab) const {
        uint32_t x[8], vr;
        load<0x7351>(slab, x, &vr);
        return {fix(x[0], vr, 8), fix(x[1], vr, 9), fix(x[2], vr, 10), fix(x[3], vr, 11)};
    }
};

// ---- what the warps of one env share ------------------------------------------
constexpr int TILE_DRAW_ROUNDS = 8;     // rounds of 32 draws published at a time
struct TileSync {
    uint8_t cand[TILE_MAX_WARPS][32];   // spawn candidates per tile and lane (row)
    int delta;                          // change of sum(table * alive_counts)
    int goal_flags;                     // 1: some goal cell changed, 2: a spawner is left in the goals
    int special[2];                     // multi-step kernel: a P / I / S / E bit changed (by step parity)
    uint32_t ok[TILE_DRAW_ROUNDS];      // multi-step kernel: outcomes of 32 draws per word
    uint64_t state[2];                  //                    and the stream after them
};

__device__ __forceinline__ void env_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Draw index of the first candidate of this lane's word; *total = all candidates of
// the board.  Raster order: rows first, then the tiles of a row from left to right.
__device__ __forceinline__ int tile_draw_offsets(const TileSync *sy, const TileParams &tp, int tile,
                                                 int lane, int *total) {
    const int tr = tp.tr_of_tile[tile], tc = tp.tc_of_tile[tile];
    int row_tot = 0, left = 0, earlier = 0, all = 0;
    for (int t = 0; t < tp.TW; t++) {
        const int c = sy->cand[t][lane], tr2 = tp.tr_of_tile[t];
        all += c;
        if (tr2 == tr) {
            row_tot += c;
            if ((int)tp.tc_of_tile[t] < tc) left += c;
        } else if (tr2 < tr) {
            earlier += c;
        }
    }
    *total = __reduce_add_sync(FULL, all);
    if (*total == 0) return 0;
    int incl = row_tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    return __reduce_add_sync(FULL, earlier) + incl - row_tot + left;
}

// Outcomes of this warp's candidates out of the board's `total` draws (every warp
// of the env runs all rounds; see the header).  The state advances by `total`.
__device__ __forceinline__ uint32_t tile_spawn_draws(u128 &state, const SpawnCtx &ctx, uint64_t thr,
                                                     uint32_t cand, int next, int total) {
    uint32_t todo = cand, got = 0;
    for (int base = 0; base < total; base += 32) {
        const u128 sn = add128(mul128(ctx.la, state), ctx.ginc);   // state after lane + 1 more steps
        const uint32_t ok = __ballot_sync(FULL, (pcg_output(sn) >> 11) < thr);
        while (todo != 0 && next < base + 32) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            got |= ((ok >> (next - base)) & 1u) << j;
            next++;
        }
        const int last = min(31, total - base - 1);
        state.hi = __shfl_sync(FULL, sn.hi, last);
        state.lo = __shfl_sync(FULL, sn.lo, last);
    }
    return got;
}

// Write this tile's changed board cells back into the slab; returns what they add
// to sum(table * alive_counts), priced with the goals as they are in `goals`.  A
// cell that the rule changes is never FROZEN before or after, so it is always
// counted (advance_board.c:192-207) and only its colour column matters: cells that
// died leave column `colour`, new cells enter column `colour`, the other side of
// either change is the "empty" column (looked up only if it is not all zero).
__device__ __forceinline__ int tile_apply_board(const Tile &t, const RowDeltaT<uint32_t> &d,
                                                uint32_t spawned, uint16_t *board,
                                                const uint16_t *goals, const int32_t *table,
                                                bool dead_counts) {
    const uint32_t fresh = d.born | spawned, destr = (d.dn & d.born) | spawned;
    const int base = t.cell_base();
    int delta = 0;
    uint32_t m = d.die;
    while (m != 0) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const int g9 = (int)((goals[base + j] & COLORS) >> COLOR_SHIFT) * 9;
        delta -= table[g9 + ((board[base + j] & COLORS) >> COLOR_SHIFT)];
        if (dead_counts) delta += table[g9 + 8];
        board[base + j] = 0;
    }
    m = fresh;
    while (m != 0) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t colour = ((d.tR >> j) & 1u) | (((d.tG >> j) & 1u) << 1) | (((d.tB >> j) & 1u) << 2);
        const int g9 = (int)((goals[base + j] & COLORS) >> COLOR_SHIFT) * 9;
        delta += table[g9 + colour];
        if (dead_counts) delta -= table[g9 + 8];
        board[base + j] = (uint16_t)(ALIVE | (((destr >> j) & 1u) << 3) | (colour << COLOR_SHIFT));
    }
    return delta;
}

// The same for the goals plane; a goal cell whose colour changes re-prices the
// board cell on top of it (`board` is already the new board).
__device__ __forceinline__ int tile_apply_goals(const Tile &t, const RowDeltaT<uint32_t> &d,
                                                uint32_t spawned, uint16_t *goals,
                                                const uint16_t *board, const int32_t *table,
                                                bool dead_counts) {
    const uint32_t fresh = d.born | spawned, destr = (d.dn & d.born) | spawned;
    uint32_t m = d.die | fresh;
    const int base = t.cell_base();
    int delta = 0;
    while (m != 0) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t v = fresh_cell(fresh, destr, d.tR, d.tG, d.tB, j);
        const uint32_t old = goals[base + j];
        if ((old ^ v) & COLORS) {
            const uint32_t cell = board[base + j];
            delta += cell_points(cell, v, table, dead_counts) - cell_points(cell, old, table, dead_counts);
        }
        goals[base + j] = (uint16_t)v;
    }
    return delta;
}

// ---- shared memory of the tiled step kernel -------------------------------------
constexpr int TILE_SMEM_HEAD = 32;      // two mbarriers

__host__ __device__ inline int tile_group_stride(int epc, int cells) { return (epc * cells * 2 + 15) & ~15; }
__host__ __device__ inline size_t tile_step_smem_bytes(int epc, int cells, int obs_stage) {
    const size_t sync = ((size_t)epc * sizeof(TileSync) + 15) & ~(size_t)15;
    return TILE_SMEM_HEAD + WARP_SMEM_TABLE + sync + (size_t)epc * 96 +
           2 * (size_t)tile_group_stride(epc, cells) + 256 + (size_t)epc * obs_stage;
}

// Observation of one env by the TW warps of the env (obs.cuh): blocks of view rows
// for the gather, then, after an env barrier, rounds of output bytes.
__device__ __noinline__ void tile_write_obs(const SlEnvBatch &B, const ObsParams &O, int e,
                                            const SlEnvState *st, const uint16_t *board,
                                            const uint16_t *goals, unsigned char *stage, int part,
                                            int nparts, int bar_id, int env_threads, int lane) {
    warp_write_obs_any<true, true>(B, O, e, st, board, goals, stage, lane, part, nparts,
                                   [&] { env_barrier(bar_id, env_threads); });
}

// SafeLifeEnv.step (safelife_env.py:148-201), tiled: TW warps per env, epc envs per CTA.
__global__ void __launch_bounds__(TILE_MAX_WARPS * 32, 2)      // 64 registers: four 8-warp CTAs per SM
k_env_step_tile(const __grid_constant__ SlEnvBatch B, const __grid_constant__ ObsParams O,
                const __grid_constant__ SlWrapSpec Wp, const __grid_constant__ TileParams sh,
                const uint8_t *__restrict__ actions, int bulk_ok, int flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = B.H, W = B.W, cells = H * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int es = sh.env_of_warp[warp], tile = sh.tile_of_warp[warp];   // env slot of the CTA, tile of the env
    const int e0 = blockIdx.x * sh.epc;
    const int nenv = min(sh.epc, B.n_envs - e0);
    const int e = e0 + es;
    const int group_bytes = nenv * cells * 2, state_bytes = nenv * 96;
    const int env_threads = sh.TW * 32, bar_id = 1 + es;

    uint64_t *bar = (uint64_t *)smem_raw;                   // [0] state + boards, [1] goals
    int32_t *stable = (int32_t *)(smem_raw + TILE_SMEM_HEAD);
    TileSync *sync_all = (TileSync *)(smem_raw + TILE_SMEM_HEAD + WARP_SMEM_TABLE);
    const size_t sync_bytes = ((size_t)sh.epc * sizeof(TileSync) + 15) & ~(size_t)15;
    SlEnvState *sstate = (SlEnvState *)((unsigned char *)sync_all + sync_bytes);
    uint16_t *sboard = (uint16_t *)((unsigned char *)sstate + sh.epc * 96);
    uint16_t *sgoals = sboard + tile_group_stride(sh.epc, cells) / 2;
    uint16_t *gboard = B.board + (size_t)e0 * cells;
    uint16_t *ggoals = B.goals + (size_t)e0 * cells;
    SlEnvState *gstate = B.state + e0;
    const bool bulk = bulk_ok && (group_bytes & 15) == 0;

    pdl_launch_dependents();
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            pdl_wait();
            mbar_expect_tx(bar, (uint32_t)(group_bytes + state_bytes));
            bulk_load(sstate, gstate, state_bytes, bar);
            bulk_load(sboard, gboard, group_bytes, bar);
            mbar_expect_tx(bar + 1, (uint32_t)group_bytes);
            bulk_load(sgoals, ggoals, group_bytes, bar + 1);
        }
    } else {
        pdl_wait();
        copy_slabs(sstate, gstate, state_bytes);
        copy_slabs(sboard, gboard, group_bytes);
        copy_slabs(sgoals, ggoals, group_bytes);
    }
    CtaShared cs;
    cs.scratch = nullptr;
    cs.lock = nullptr;
    pdl_wait();
    cs.table = stage_table(B, stable);
    if (tile == 0 && lane == 0 && es < sh.epc) {
        sync_all[es].delta = 0;
        sync_all[es].goal_flags = 0;
    }
    __syncthreads();                      // mbarrier init / fallback copy / table / sync records
    if (bulk) mbar_wait(bar, 0);

    bool goals_dirty = false;
    if (es < nenv) {
        TileSync *sy = sync_all + es;
        SlEnvState *st = sstate + es;
        uint16_t *board = sboard + es * cells, *goals = sgoals + es * cells;
        const Tile t(H, W, sh, tile, lane);
        const uint64_t thr = st->spawn_thr;
        u128 state = {st->rng[0], st->rng[1]};
        const u128 inc = {st->rng[2], st->rng[3]};
        const uint64_t state_lo0 = state.lo;
        int sg = st->static_goals;
        const int32_t *table = cs.table ? cs.table : B.pool_table + (size_t)st->level_id * 72;
        const bool dead_counts = st->empty_points != 0;
        int delta = 0;

        // execute_actions (advance_board.c:217-300) by the env's first warp: lane 0
        // acts, lanes 0-3 account for the (at most four) cells it can touch
        uint32_t before = 0, after = 0;
        int idx = 0;
        if (tile == 0) {
            int row = st->loc_row, col = st->loc_col;
            const int action = actions[e];
            if (row >= 0 && action != 0) {
                const ActionCells ac = action_cells(H, W, row, col, action);
                idx = lane == 0 ? ac.i0 : (lane == 1 ? ac.i1 : (lane == 2 ? ac.i2 : ac.i3));
                const bool mine = lane < 4 && !(lane == 3 && ac.i3 == ac.i2);
                if (mine) before = board[idx];
                __syncwarp();
                if (lane == 0) {
                    agent_action_at(board, ac, row, col, action);
                    st->loc_row = row;
                    st->loc_col = col;
                }
                __syncwarp();
                after = mine ? board[idx] : 0u;
            }
        }
        env_barrier(bar_id, env_threads);              // the board as the agent left it

        // advance_board on the tile
        {
            const LowPlanesT<uint32_t> p = t.load_low(board);
            const RowDeltaT<uint32_t> d = stencil_planes(t, p, [&] { return t.load_high(board); });
            sy->cand[tile][lane] = (uint8_t)__popc(d.cand);
            if (bulk) mbar_wait(bar + 1, 0);            // goal slabs: needed for pricing from here on
            env_barrier(bar_id, env_threads);          // counts published; every warp has read the slab
            int total;
            const int next = tile_draw_offsets(sy, sh, tile, lane, &total);
            const uint32_t spawned =
                total ? tile_spawn_draws(state, make_spawn_ctx(lane, inc), thr, d.cand, next, total) : 0u;
            delta += tile_apply_board(t, d, spawned, board, goals, table, dead_counts);
        }
        if (after != before)
            delta += cell_points(after, goals[idx], table, dead_counts) -
                     cell_points(before, goals[idx], table, dead_counts);
        if (sg != 1) {   // goals advance from the same stream (safelife_game.py:753-761)
            const LowPlanesT<uint32_t> gp = t.load_low(goals);
            const RowDeltaT<uint32_t> gd = stencil_planes(t, gp, [&] { return t.load_high(goals); });
            env_barrier(bar_id, env_threads);          // the draws of the board have been consumed
            sy->cand[tile][lane] = (uint8_t)__popc(gd.cand);
            env_barrier(bar_id, env_threads);          // goal planes read and board priced everywhere
            int total;
            const int next = tile_draw_offsets(sy, sh, tile, lane, &total);
            const uint32_t gspawned =
                total ? tile_spawn_draws(state, make_spawn_ctx(lane, inc), thr, gd.cand, next, total) : 0u;
            delta += tile_apply_goals(t, gd, gspawned, goals, board, table, dead_counts);
            const uint32_t gchanged = gd.die | gd.born | gspawned;
            const bool any_changed = __any_sync(FULL, gchanged != 0);
            const bool spawner_left = __any_sync(FULL, (gp.S & t.imask & ~gchanged) != 0);
            if (lane == 0 && (any_changed || spawner_left))
                atomicOr(&sy->goal_flags, (any_changed ? 1 : 0) | (spawner_left ? 2 : 0));
        }
        delta = __reduce_add_sync(FULL, delta);
        if (lane == 0 && delta != 0) atomicAdd(&sy->delta, delta);
        env_barrier(bar_id, env_threads);              // every tile applied and accounted for

        if (tile == 0) {
            // scalar tail, wrapper epilogue and reset by the env's first warp
            if (lane == 0) {
                st->count_points += sy->delta;
                if (sg != 1) {
                    // static iff nothing changed and no spawner is left (safelife_game.py:756-760)
                    if (sg == 0) sg = sy->goal_flags ? 2 : 1;
                    st->static_goals = (uint8_t)sg;
                }
                if (state.lo != state_lo0) {   // the LCG is a bijection: lo changes iff draws happened
                    st->rng[0] = state.hi;
                    st->rng[1] = state.lo;
                }
            }
            goals_dirty = (sg != 1) && (sy->goal_flags & 1);
        }
        if (Wp.wstate == nullptr) {
            if (tile == 0) goals_dirty |= step_tail_env(B, Wp, cs, e, st, board, goals, flags, lane);
        } else {
            // with the wrapper epilogue: scalar tail on the first warp, the side-effect
            // count on all of them, then reward and reset on the first warp again
            StepOut so = {0.f, false, false};
            int done = 0;
            if (tile == 0) {
                done = step_tail_scalar(B, cs, e, st, board, goals, lane, &so);
                if (lane == 0) sy->delta = 0;          // reused as the count
            }
            env_barrier(bar_id, env_threads);          // exit colours written, counter cleared
            if (Wp.flags & SL_WRAP_SIDE_EFFECT) {
                const int part = warp_side_effect_count(B, Wp, st->level_id, board, goals,
                                                        Wp.baseline + (size_t)e * cells, lane, tile, sh.TW);
                if (lane == 0 && part != 0) atomicAdd(&sy->delta, part);
            }
            env_barrier(bar_id, env_threads);
            if (tile == 0) {
                if (lane == 0) wrap_finish(B, Wp, e, st, so.reward, so.done, so.times_up, sy->delta);
                __syncwarp();
                goals_dirty |= step_tail_reset(B, Wp, e, st, board, goals, flags, done, lane);
            }
        }
        if (O.obs != nullptr) {
            env_barrier(bar_id, env_threads);          // the board as the step leaves it
            // the view rows / output chunks of the env dealt out over its warps
            const int stage_bytes = obs_stage_bytes(O.vh, O.vw, O.n_channels);
            unsigned char *stage = smem_raw + tile_step_smem_bytes(sh.epc, cells, 0) + (size_t)es * stage_bytes;
            tile_write_obs(B, O, e, st, board, goals, stage, tile, sh.TW, bar_id, env_threads, lane);
        }
    } else if (bulk) {
        mbar_wait(bar + 1, 0);
    }
    fence_async_smem();                   // generic-proxy writes -> visible to the bulk engine
    const int any_goals_dirty = __syncthreads_or(goals_dirty);
    if (bulk) {
        if (threadIdx.x == 0) {
            bulk_store(gboard, sboard, group_bytes);
            if (any_goals_dirty) bulk_store(ggoals, sgoals, group_bytes);
            bulk_store(gstate, sstate, state_bytes);
            bulk_commit_and_wait_read();
        }
    } else {
        copy_slabs(gboard, sboard, group_bytes);
        if (any_goals_dirty) copy_slabs(ggoals, sgoals, group_bytes);
        copy_slabs(gstate, sstate, state_bytes);
    }
}


// ---------------------------------------------------------------------------
// n-step advance / life occupancy / the simulation part of side_effect_score on
// tiles: advance_board_nstep (advance_board.c:128-149), life_occupancy
// (:153-189) and side_effects.py:101-111 for boards beyond 32x32.  As in the
// one-warp-per-board kernels the planes stay in registers for all the steps; what
// changes is that a board is spread over TW warps, so after every step the warps
// refresh their halo ring from their neighbours through shared memory: first
// the halo rows (a neighbour's whole boundary word), then, with the halo rows in
// place, the halo columns (one ballot per plane and side, which brings the
// corners along).  A small batch of large boards (the finished episodes of a
// step) therefore occupies TW times as many warps as before.
// ---------------------------------------------------------------------------
struct TileHalo {
    uint32_t row[2][9];     // [top / bottom interior row][plane]
    uint32_t col[2][9];     // [first / last interior column][plane], bit l = lane l
};

struct TileBoard {
    LowPlanesT<uint32_t> lp;
    HighPlanesT<uint32_t> hp;
    uint32_t agent, kept;
};

__device__ __forceinline__ TileBoard tile_board_load(const Tile &t, const uint16_t *slab) {
    TileBoard b;
    uint32_t x[8], vr;
    t.load<0x6240>(slab, x, &vr);
    b.lp = {t.fix(x[0], vr, 0), t.fix(x[3], vr, 3), x[4], t.fix(x[5], vr, 5), t.fix(x[6], vr, 6),
            t.fix(x[7], vr, 7)};
    b.agent = x[1] & t.imask;
    b.lp.F &= t.imask;
    b.hp = t.load_high(slab);
    b.kept = t.imask;
    return b;
}

// The nine planes whose neighbours the rule reads, by index.
__device__ __forceinline__ uint32_t &tile_plane(TileBoard &b, int p) {
    switch (p) {
        case 0: return b.lp.A;
        case 1: return b.lp.D;
        case 2: return b.lp.P_;
        case 3: return b.lp.I;
        case 4: return b.lp.S;
        case 5: return b.hp.E;
        case 6: return b.hp.R;
        case 7: return b.hp.G;
        default: return b.hp.B;
    }
}

// Halo refresh after a step.  `mine`: this tile's exchange record, up / down / left /
// right: its neighbours' (possibly itself on a single tile row / column).  Planes
// 2-5 (PRESERVING, INHIBITING, SPAWNING, EXIT) only change when such a cell dies,
// which is rare: `special` (this warp saw one die) is OR-ed over the env's warps
// through sy->special[parity] and those planes are exchanged only then.
__device__ __forceinline__ void tile_halo_exchange(const Tile &t, TileBoard &b, TileSync *sy,
                                                   int parity, bool special, TileHalo *mine,
                                                   const TileHalo *up, const TileHalo *down,
                                                   const TileHalo *left, const TileHalo *right,
                                                   int bar_id, int env_threads) {
    const int lane = t.lane;
#pragma unroll
    for (int p = 0; p < 9; p++) {
        const uint32_t x = tile_plane(b, p);
        if (lane == 1) mine->row[0][p] = x;
        if (lane == t.hr) mine->row[1][p] = x;
    }
    if (special && lane == 0) atomicOr(&sy->special[parity], 1);
    env_barrier(bar_id, env_threads);
    const bool all = sy->special[parity] != 0;
#pragma unroll
    for (int p = 0; p < 9; p++) {
        if (p >= 2 && p <= 5 && !all) continue;
        uint32_t &x = tile_plane(b, p);
        if (lane == 0) x = up->row[1][p];
        if (lane == t.hr + 1) x = down->row[0][p];
    }
    const int first = t.boff, last = t.boff + t.wc - 1;
#pragma unroll
    for (int p = 0; p < 9; p++) {
        if (p >= 2 && p <= 5 && !all) continue;
        const uint32_t x = tile_plane(b, p);
        const uint32_t c0 = __ballot_sync(FULL, (x >> first) & 1u), c1 = __ballot_sync(FULL, (x >> last) & 1u);
        if (lane == 0) {
            mine->col[0][p] = c0;
            mine->col[1][p] = c1;
        }
    }
    if (lane == 0) sy->special[parity ^ 1] = 0;     // the next step's flag (every warp writes the same 0)
    env_barrier(bar_id, env_threads);
    const uint32_t keep = ~((1u << (first - 1)) | (1u << (last + 1)));
#pragma unroll
    for (int p = 0; p < 9; p++) {
        if (p >= 2 && p <= 5 && !all) continue;
        uint32_t &x = tile_plane(b, p);
        const uint32_t l = (left->col[1][p] >> lane) & 1u, r = (right->col[0][p] >> lane) & 1u;
        x = (x & keep) | (l << (first - 1)) | (r << (last + 1));
    }
}

// One advance_board step on the register image of a tile (halos valid on entry and
// on exit).  `draws`: the board may draw at all (an RNG was given and it has a spawner).
__device__ __forceinline__ void tile_board_step(const Tile &t, TileBoard &b, u128 &state,
                                                const u128 &inc, uint64_t thr, bool draws, int parity,
                                                TileSync *sy, const TileParams &tp, int tile,
                                                TileHalo *mine, const TileHalo *up,
                                                const TileHalo *down, const TileHalo *left,
                                                const TileHalo *right, int bar_id, int env_threads) {
    LowPlanesT<uint32_t> &lp = b.lp;
    HighPlanesT<uint32_t> &hp = b.hp;
    const RowDeltaT<uint32_t> d = stencil_planes(t, lp, [&] { return hp; });
    uint32_t spawned = 0u;
    if (draws) {
        sy->cand[tile][t.lane] = (uint8_t)__popc(d.cand);
        env_barrier(bar_id, env_threads);
        int total;
        int next = tile_draw_offsets(sy, tp, tile, t.lane, &total);
        // The draws of the whole board, TILE_DRAW_ROUNDS x 32 at a time: the env's first warp
        // runs the PCG64 leapfrog and publishes the outcome words and the stream, every warp
        // picks the outcomes of its own candidates (raster order, see tile_draw_offsets).
        uint32_t todo = d.cand;
        for (int base = 0; base < total; base += 32 * TILE_DRAW_ROUNDS) {
            const int rounds = min(TILE_DRAW_ROUNDS, (total - base + 31) >> 5);
            if (tile == 0) {
                const SpawnCtx ctx = make_spawn_ctx(t.lane, inc);
                for (int r = 0; r < rounds; r++) {
                    const u128 sn = add128(mul128(ctx.la, state), ctx.ginc);
                    const uint32_t ok = __ballot_sync(FULL, (pcg_output(sn) >> 11) < thr);
                    const int last = min(31, total - base - 32 * r - 1);
                    state.hi = __shfl_sync(FULL, sn.hi, last);
                    state.lo = __shfl_sync(FULL, sn.lo, last);
                    if (t.lane == 0) sy->ok[r] = ok;
                }
                if (t.lane == 0) {
                    sy->state[0] = state.hi;
                    sy->state[1] = state.lo;
                }
            }
            env_barrier(bar_id, env_threads);
            for (int r = 0; r < rounds; r++) {
                const uint32_t ok = sy->ok[r];
                const int lo = base + 32 * r;
                while (todo != 0 && next < lo + 32) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1;
                    spawned |= ((ok >> (next - lo)) & 1u) << j;
                    next++;
                }
            }
            state = {sy->state[0], sy->state[1]};
            if (base + 32 * TILE_DRAW_ROUNDS < total) env_barrier(bar_id, env_threads);   // words consumed
        }
    }
    const uint32_t fresh = d.born | spawned, changed = d.die | fresh, keep = ~changed;
    const bool special = __any_sync(FULL, (changed & (lp.P_ | lp.I | lp.S | hp.E)) != 0);
    lp.A = (lp.A & keep) | fresh;
    lp.D = (lp.D & keep) | (d.dn & d.born) | spawned;
    hp.R = (hp.R & keep) | (d.tR & fresh);
    hp.G = (hp.G & keep) | (d.tG & fresh);
    hp.B = (hp.B & keep) | (d.tB & fresh);
    lp.P_ &= keep; lp.I &= keep; lp.S &= keep; hp.E &= keep; b.agent &= keep;
    b.kept &= keep;                     // (F is already 0 on every changed cell)
    tile_halo_exchange(t, b, sy, parity, special, mine, up, down, left, right, bar_id, env_threads);
}

// Bit-sliced occupancy counters of the tile's interior (see OccCounters).
struct TileOcc {
    uint32_t cnt[8][OCC_BITS];
    int pending;

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
            for (int k = 0; k < OCC_BITS; k++) cnt[c][k] = 0;
        pending = 0;
    }
    __device__ __forceinline__ void flush(const Tile &t, int32_t *board_counts) {
        if (t.imask != 0) {
            int32_t *dst = board_counts + (size_t)t.cell_base() * 8;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t nz = 0;
#pragma unroll
                for (int k = 0; k < OCC_BITS; k++) nz |= cnt[c][k];
                while (nz != 0) {
                    const int j = __ffs(nz) - 1;
                    nz &= nz - 1;
                    int v = 0;
#pragma unroll
                    for (int k = 0; k < OCC_BITS; k++) v |= ((cnt[c][k] >> j) & 1u) << k;
                    dst[j * 8 + c] += v;
                }
#pragma unroll
                for (int k = 0; k < OCC_BITS; k++) cnt[c][k] = 0;
            }
        }
        pending = 0;
    }
    // +1 for cells ALIVE and none of AGENT | EXIT | FROZEN, by colour (advance_board.c:153-162)
    __device__ __forceinline__ void add(const Tile &t, const TileBoard &b, int32_t *board_counts) {
        const uint32_t occ = b.lp.A & ~(b.agent | b.hp.E | b.lp.F) & t.imask;
        if (__any_sync(FULL, occ != 0)) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t carry = occ & ((c & 1) ? b.hp.R : ~b.hp.R) & ((c & 2) ? b.hp.G : ~b.hp.G) &
                                 ((c & 4) ? b.hp.B : ~b.hp.B);
                if (__any_sync(FULL, carry != 0)) {
#pragma unroll
                    for (int k = 0; k < OCC_BITS; k++) {
                        const uint32_t tmp = cnt[c][k] & carry;
                        cnt[c][k] ^= carry;
                        carry = tmp;
                    }
                }
            }
        }
        if (++pending == (1 << OCC_BITS) - 1) flush(t, board_counts);
    }
};

__host__ __device__ inline size_t tile_multistep_smem_bytes(int epc, int TW, int cells) {
    const size_t sync = ((size_t)epc * sizeof(TileSync) + 15) & ~(size_t)15;
    return sync + (size_t)epc * TW * sizeof(TileHalo) + 64 + (size_t)epc * (((size_t)cells * 2 + 15) & ~(size_t)15) + 256;
}

// mode 0: out[b] = board after n_steps (advance_board_nstep); mode 1: counts[b] +=
// life_occupancy over n_steps; mode 2: the three runs of side_effect_score
// (counts[b][0..1], first run preceded by lead_steps[b] uncounted steps).
__global__ void __launch_bounds__(TILE_MAX_WARPS * 32)
k_multistep_tile(const __grid_constant__ TileParams sh, int mode, const uint16_t *in,
                 const uint16_t *in2, uint16_t *out, int32_t *__restrict__ counts, int n, int H,
                 int W, uint64_t thr_all, const uint64_t *__restrict__ thr_each, int n_steps,
                 const int32_t *__restrict__ lead_steps, uint64_t *__restrict__ rng) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cells = H * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int es = sh.env_of_warp[warp], tile = sh.tile_of_warp[warp];
    const int env_threads = sh.TW * 32, bar_id = 1 + es;
    const bool counting = mode != 0;
    const size_t sync_bytes = ((size_t)sh.epc * sizeof(TileSync) + 15) & ~(size_t)15;
    TileSync *sy = (TileSync *)smem_raw + es;
    TileHalo *halos = (TileHalo *)(smem_raw + sync_bytes) + es * sh.TW;
    const size_t slab_bytes = ((size_t)cells * 2 + 15) & ~(size_t)15;
    uint16_t *slab = (uint16_t *)(smem_raw + sync_bytes + (size_t)sh.epc * sh.TW * sizeof(TileHalo) + 64 +
                                  es * slab_bytes);
    const Tile t(H, W, sh, tile, lane);
    // neighbours on the torus of tiles
    const int tr = sh.tr_of_tile[tile], tc = sh.tc_of_tile[tile];
    const int tr_up = tr == 0 ? sh.TR - 1 : tr - 1, tr_dn = tr == sh.TR - 1 ? 0 : tr + 1;
    const int tc_l = tc == 0 ? sh.TC - 1 : tc - 1, tc_r = tc == sh.TC - 1 ? 0 : tc + 1;
    TileHalo *mine = halos + tile;
    const TileHalo *up = halos + tr_up * sh.TC + tc, *down = halos + tr_dn * sh.TC + tc;
    const TileHalo *left = halos + tr * sh.TC + tc_l, *right = halos + tr * sh.TC + tc_r;

    for (int b0 = blockIdx.x * sh.epc; b0 < n; b0 += gridDim.x * sh.epc) {
        const int b = b0 + es;
        if (b >= n) continue;                       // (whole envs drop out together: no barrier is split)
        u128 state = {0, 0}, inc = {0, 0};
        if (rng != nullptr) {
            state = {rng[4 * (size_t)b], rng[4 * (size_t)b + 1]};
            inc = {rng[4 * (size_t)b + 2], rng[4 * (size_t)b + 3]};
        }
        const uint64_t th = thr_each != nullptr ? thr_each[b] : thr_all;
        const int runs = mode == 2 ? 2 : 1;
        TileOcc occ;
        for (int run = 0; run < runs; run++) {
            const uint16_t *src = (run == 0 ? in : in2) + (size_t)b * cells;
            env_barrier(bar_id, env_threads);       // the slab is free
            for (int i = tile * 32 + lane; i < cells; i += env_threads) slab[i] = src[i];
            if (tile == 0 && lane == 0) {
                sy->goal_flags = 0;
                sy->special[0] = sy->special[1] = 0;
            }
            env_barrier(bar_id, env_threads);
            TileBoard bd = tile_board_load(t, slab);
            int step_no = 0;
            // a board without a spawner never draws (spawners are not created)
            if (__any_sync(FULL, (bd.lp.S & t.imask) != 0) && lane == 0) atomicOr(&sy->goal_flags, 1);
            env_barrier(bar_id, env_threads);
            const bool draws = rng != nullptr && sy->goal_flags != 0;
            if (mode == 2 && run == 0) {
                const int lead = lead_steps[b];
                for (int s = 0; s < lead; s++, step_no++)
                    tile_board_step(t, bd, state, inc, th, draws, step_no & 1, sy, sh, tile, mine, up, down,
                                    left, right, bar_id, env_threads);
            }
            if (!counting) {
                for (int s = 0; s < n_steps; s++, step_no++)
                    tile_board_step(t, bd, state, inc, th, draws, step_no & 1, sy, sh, tile, mine, up, down,
                                    left, right, bar_id, env_threads);
                // cells that changed at least once are fully described by the planes
                uint32_t m = t.imask & ~bd.kept;
                const int base = t.cell_base();
                while (m != 0) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    slab[base + j] = (uint16_t)(((bd.lp.A >> j) & 1u) | (((bd.lp.D >> j) & 1u) << 3) |
                                                (((bd.hp.R >> j) & 1u) << 9) | (((bd.hp.G >> j) & 1u) << 10) |
                                                (((bd.hp.B >> j) & 1u) << 11));
                }
                env_barrier(bar_id, env_threads);
                uint16_t *dst = out + (size_t)b * cells;
                for (int i = tile * 32 + lane; i < cells; i += env_threads) dst[i] = slab[i];
            } else {
                int32_t *dst = counts + ((size_t)b * runs + run) * cells * 8;
                occ.clear();
                for (int s = 0; s < n_steps; s++, step_no++) {
                    tile_board_step(t, bd, state, inc, th, draws, step_no & 1, sy, sh, tile, mine, up, down,
                                    left, right, bar_id, env_threads);
                    occ.add(t, bd, dst);
                }
                if (occ.pending) occ.flush(t, dst);
            }
        }
        if (rng != nullptr && tile == 0 && lane == 0) {
            rng[4 * (size_t)b] = state.hi;
            rng[4 * (size_t)b + 1] = state.lo;
        }
    }
}

}  // namespace sl
