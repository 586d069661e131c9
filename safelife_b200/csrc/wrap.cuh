// wrap.cuh -- the reference's training wrappers (safelife/env_wrappers.py) as an
// epilogue of the step: per env and per step they are three small reductions
// over state the step kernel already holds, so they run in the same launch.
//
//   MovementBonusWrapper    (:32-99)    L1 distance to the agent position `period`
//                                       steps ago (a ring in the wrapper record)
//   ExtraExitBonus          (:121-128)  bonus * episode_reward on a non-timeout done
//   SimpleSideEffectPenalty (:148-213)  cells that differ from the baseline board,
//                                       player bits and exit cells ignored
//   MinPerformanceScheduler (:131-145)  required_points rescaled after every reset
//
// in the order training/env_factory.py:274-283 stacks them.  The wrapped reward
// is float64, as the reference's becomes once a wrapper adds a Python float to
// the float32 base reward; each term is added with one IEEE operation in the
// reference's order (no contraction), so the result is the same double.
#pragma once
#include <stdint.h>
#include "cells.cuh"
#include "../../include/safelife_b200.h"

namespace sl {

constexpr uint32_t WRAP_KEEP_BITS = 0xffffu & ~(AGENT | DESTRUCTIBLE | FROZEN | PRESERVING | INHIBITING);

// One cell of SimpleSideEffectPenalty.step (env_wrappers.py:186-208): does it count
// as a side effect?  b, bl: board / baseline cell; g: goal cell (ignore_reward_cells only).
__device__ __forceinline__ bool side_effect_cell(uint32_t b, uint32_t bl, uint32_t g, bool ignore_reward) {
    b &= WRAP_KEEP_BITS;
    bl &= WRAP_KEEP_BITS;
    if (b == bl) return false;
    if (!ignore_reward) return true;
    const uint32_t red_life = ALIVE | COLOR_R;
    const bool start_red = (bl & red_life) == red_life, end_red = (b & red_life) == red_life;
    const bool goal_cell = (g & COLORS) == (4u << COLOR_SHIFT), end_alive = (b & red_life) == ALIVE;
    return !((start_red && !end_red) || (goal_cell && end_alive));
}

// number of non-zero 16-bit lanes of x
__device__ __forceinline__ int nonzero_u16x4(uint64_t x) {
    const uint64_t m = 0x7fff7fff7fff7fffull;
    return __popcll((((x & m) + m) | x) & 0x8000800080008000ull);
}

// Side-effect cell count of one env: the share of warp `part` of `nparts` (the sum
// over the parts is the count).  board / goals: the env's planes (shared or global
// memory), baseline: its baseline board (HBM).  Exit cells are excluded
// (`board[i1,i2] = baseline_board[i1,i2]`, :190-191).  Four cells per load where
// the planes allow it.
__device__ __forceinline__ int warp_side_effect_count(const SlEnvBatch &B, const SlWrapSpec &Wp,
                                                      int level_id, const uint16_t *board,
                                                      const uint16_t *goals,
                                                      const uint16_t *baseline, int lane,
                                                      int part = 0, int nparts = 1) {
    const int cells = B.H * B.W;
    const bool ignore = (Wp.flags & SL_WRAP_IGNORE_REWARD_CELLS) != 0;
    int n = 0;
    if (!ignore && (cells & 3) == 0 && ((((uintptr_t)board) | ((uintptr_t)baseline)) & 7) == 0) {
        const uint64_t keep4 = (uint64_t)WRAP_KEEP_BITS * 0x0001000100010001ull;
        const uint64_t *b8 = (const uint64_t *)board, *bl8 = (const uint64_t *)baseline;
        for (int i = part * 32 + lane; i < (cells >> 2); i += 32 * nparts)
            n += nonzero_u16x4((b8[i] ^ bl8[i]) & keep4);
    } else {
        for (int i = part * 32 + lane; i < cells; i += 32 * nparts)
            n += side_effect_cell(board[i], baseline[i], ignore ? goals[i] : 0u, ignore) ? 1 : 0;
    }
    if (lane == 0 && part == 0) {
        const int32_t *exits = B.pool_exits + (size_t)level_id * B.max_exits;
        for (int k = 0; k < B.max_exits; k++) {
            const int idx = exits[k];
            if (idx >= 0 && side_effect_cell(board[idx], baseline[idx], ignore ? goals[idx] : 0u, ignore))
                n--;
        }
    }
    return __reduce_add_sync(0xffffffffu, n);
}

// The wrapped reward of one env for the step that just finished (one thread).
// row, col: the agent's location after the step (row < 0: no agent);
// side_effect: the current cell count (ignored unless SL_WRAP_SIDE_EFFECT).
__device__ __forceinline__ double wrap_reward(const SlEnvBatch &B, const SlWrapSpec &Wp,
                                              SlWrapState &ws, int row, int col, float reward,
                                              bool done, bool times_up, float ep_reward,
                                              int side_effect) {
    double r = (double)reward;
    if (Wp.flags & SL_WRAP_MOVEMENT) {
        const int n = Wp.movement_period;
        int dist = 0;
        if (row >= 0) {   // no agent: an empty agent_locs array sums to speed 0 (:74-87)
            // the deque holds at most n positions; [-n] of a full one and [0] of a
            // shorter one are both its oldest entry
            dist = abs(row - (int)ws.hist[0][0]) + abs(col - (int)ws.hist[0][1]) + (n - ws.hist_len);
            if (ws.hist_len == n) {
                for (int k = 0; k + 1 < n; k++) {
                    ws.hist[k][0] = ws.hist[k + 1][0];
                    ws.hist[k][1] = ws.hist[k + 1][1];
                }
                ws.hist_len = n - 1;
            }
            ws.hist[ws.hist_len][0] = (int16_t)row;
            ws.hist[ws.hist_len][1] = (int16_t)col;
            ws.hist_len++;
        }
        r = __dadd_rn(r, Wp.movement_table[dist]);           // reward += bonus * speed ** power
        if (Wp.flags & SL_WRAP_MOVEMENT_PENALTY) r = __dsub_rn(r, Wp.movement_bonus);
    }
    if ((Wp.flags & SL_WRAP_EXIT_BONUS) && !times_up)         // :124-127
        r = __dadd_rn(r, __dmul_rn(__dmul_rn(done ? 1.0 : 0.0, Wp.exit_bonus), (double)ep_reward));
    if (Wp.flags & SL_WRAP_SIDE_EFFECT) {                     // :210-212
        const int delta = side_effect - ws.last_side_effect;
        r = __dsub_rn(r, __dmul_rn((double)delta, Wp.penalty_coef));
        ws.last_side_effect = side_effect;
    }
    return r;
}

// Wrapper state of one env after a reset (the wrappers' own reset methods,
// :95-99, :141-145, :171-175), by one warp: position ring = [start], no side
// effect yet, baseline board = the fresh board, required points rescaled.
// `board`: the env's board after the reset (shared or global memory).
__device__ __forceinline__ void warp_wrap_reset(const SlEnvBatch &B, const SlWrapSpec &Wp, int e,
                                                SlEnvState *st, const uint16_t *board, int lane) {
    if (lane == 0) {
        SlWrapState ws = {};
        ws.hist[0][0] = (int16_t)st->loc_row;
        ws.hist[0][1] = (int16_t)st->loc_col;
        ws.hist_len = 1;
        ws.last_side_effect = 0;
        Wp.wstate[e] = ws;
        if (Wp.pool_required != nullptr) st->required = Wp.pool_required[st->level_id];
    }
    if (Wp.flags & SL_WRAP_SIDE_EFFECT) {
        const int cells = B.H * B.W;
        uint16_t *bl = Wp.baseline + (size_t)e * cells;
        for (int i = lane; i < cells; i += 32) bl[i] = board[i];
    }
    __syncwarp();
}

// The wrapped reward and the wrapper outputs of env e, given its side-effect count
// (lane 0 of one warp).  reward / done / times_up: what finish_step just wrote.
__device__ __forceinline__ void wrap_finish(const SlEnvBatch &B, const SlWrapSpec &Wp, int e,
                                            const SlEnvState *st, float reward, bool done,
                                            bool times_up, int count) {
    SlWrapState ws = Wp.wstate[e];
    Wp.reward[e] = wrap_reward(B, Wp, ws, st->loc_row, st->loc_col, reward, done, times_up,
                               st->ep_reward, count);
    Wp.wstate[e] = ws;
    if (Wp.side_effect != nullptr) Wp.side_effect[e] = count;
}

// The epilogue of one env inside a step kernel, between the scalar tail and the
// reset, by one warp: count, reward, outputs.
__device__ __noinline__ void warp_wrap_step(const SlEnvBatch &B, const SlWrapSpec &Wp, int e,
                                            const SlEnvState *st, const uint16_t *board,
                                            const uint16_t *goals, float reward, bool done,
                                            bool times_up, int lane) {
    int count = 0;
    if (Wp.flags & SL_WRAP_SIDE_EFFECT)
        count = warp_side_effect_count(B, Wp, st->level_id, board, goals,
                                       Wp.baseline + (size_t)e * B.H * B.W, lane);
    if (lane == 0) wrap_finish(B, Wp, e, st, reward, done, times_up, count);
    __syncwarp();
}

// Stand-alone forms for the generic (CTA-per-env) path and for host-driven resets.
__global__ void __launch_bounds__(256)
k_wrap_step(const __grid_constant__ SlEnvBatch B, const __grid_constant__ SlWrapSpec Wp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cells = B.H * B.W;
    for (int e = blockIdx.x * 8 + warp; e < B.n_envs; e += gridDim.x * 8)
        warp_wrap_step(B, Wp, e, B.state + e, B.board + (size_t)e * cells, B.goals + (size_t)e * cells,
                       B.reward[e], B.done[e] != 0, B.times_up[e] != 0, lane);
}

__global__ void __launch_bounds__(256)
k_wrap_reset(const __grid_constant__ SlEnvBatch B, const __grid_constant__ SlWrapSpec Wp,
             const uint8_t *__restrict__ mask) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cells = B.H * B.W;
    for (int e = blockIdx.x * 8 + warp; e < B.n_envs; e += gridDim.x * 8) {
        if (mask != nullptr && !mask[e]) continue;
        warp_wrap_reset(B, Wp, e, B.state + e, B.board + (size_t)e * cells, lane);
    }
}

}  // namespace sl
