// obs.cuh -- SafeLifeEnv.get_obs by one warp per env.
//
// safelife_env.py:105-146 + helper_utils.recenter_view (helper_utils.py:42-75):
//     full = uint32(board) | (goal colour, white removed) << 16
//     view[r][c] = full[(y0 - vh/2 + r) mod H][(x0 - vw/2 + c) mod W]      (agent 0 at y0, x0)
//     every exit cell (iy, ix) of the level: view[clip(...)][clip(...)] = full[iy][ix]
//     obs[r][c][k] = (view[r][c] >> channels[k]) & 1          (or the raw uint32 view)
//
// The writer is shared by the fused epilogue of the step kernels (board slab in
// shared memory, goals in shared memory or HBM) and by the stand-alone
// k_get_obs_warp (both in HBM).  Lane c gathers column c of the view row by row
// -- the source column is a per-lane constant, the source row a warp-uniform
// counter, so there is no per-cell index arithmetic -- and reduces each cell to
// its C-bit channel mask in a per-warp staging array; then every lane expands
// 16 consecutive output bytes (= 16 consecutive bits of the cell-major mask
// stream) per 16-byte store.  The channel list is applied as runs of consecutive
// bits (the training list 0-4, 7-11, 25-27, 12-13 is four shift-and-mask steps).
#pragma once
#include <stdint.h>
#include "cells.cuh"
#include "../../include/safelife_b200.h"

namespace sl {

constexpr int OBS_MAX_RUNS = 32;

// Kernel-side form of SlObsSpec (built on the host by make_obs_params).
struct ObsParams {
    void *obs;                 // uint8 [n][vh][vw][C] or uint32 [n][vh][vw] (C == 0)
    int32_t vh, vw, n_channels, remove_white, n_runs;
    uint8_t run_src[OBS_MAX_RUNS], run_dst[OBS_MAX_RUNS];
    uint32_t run_mask[OBS_MAX_RUNS];
};

inline ObsParams make_obs_params(const SlObsSpec &s) {
    ObsParams p = {};
    p.obs = s.obs;
    p.vh = s.vh;
    p.vw = s.vw;
    p.n_channels = s.n_channels;
    p.remove_white = s.remove_white_goals ? 1 : 0;
    int nr = 0;
    for (int k = 0; k < s.n_channels;) {
        int len = 1;
        const int c0 = s.channels[k];
        while (k + len < s.n_channels && s.channels[k + len] == c0 + len) len++;
        p.run_src[nr] = (uint8_t)c0;
        p.run_dst[nr] = (uint8_t)k;
        p.run_mask[nr] = len >= 32 ? 0xffffffffu : ((1u << len) - 1u);
        nr++;
        k += len;
    }
    p.n_runs = nr;
    return p;
}

// Staging bytes of one warp: one mask per view cell (uint16 up to 16 channels,
// else uint32) plus zero padding that the expansion may read past the last cell.
__host__ __device__ inline int obs_stage_bytes(int vh, int vw, int n_channels) {
    if (n_channels == 0) return 0;
    const int elt = n_channels <= 16 ? 2 : 4;
    return ((vh * vw + 34) * elt + 15) & ~15;
}

// spread 16 bits into 16 bytes (bit i -> byte i = 0 / 1)
__device__ __forceinline__ uint4 expand16(uint32_t bits) {
    uint4 w;
    w.x = ((bits & 15u) * 0x00204081u) & 0x01010101u;
    w.y = (((bits >> 4) & 15u) * 0x00204081u) & 0x01010101u;
    w.z = (((bits >> 8) & 15u) * 0x00204081u) & 0x01010101u;
    w.w = (((bits >> 12) & 15u) * 0x00204081u) & 0x01010101u;
    return w;
}

// a << n with n >= 32 giving 0 (one funnel shift with a clamped amount)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t a, int n) { return __funnelshift_lc(0u, a, n); }

// Observation of env `e` by one warp.  `board` / `goals`: the env's planes, in
// shared (BSH / GSH) or global memory -- plain loads, so that planes rewritten by
// this warp's reset are seen.  `stage`: this warp's staging array in shared memory
// (ST[vh * vw + 34]); unused when n_channels == 0.  (row0, col0): agent location,
// row0 < 0 if the level has none.
// Several warps may share one env (the tiled kernels): warp `part` of `nparts`
// gathers a block of the view's rows and expands every nparts-th round of output
// bytes; `sync()` is a barrier over all of them (__syncwarp for a single warp).
template <typename ST, bool BSH, bool GSH, class SyncFn>
__device__ __forceinline__ void warp_write_obs(const SlEnvBatch &B, const ObsParams &O, int e,
                                               int row0, int col0, int level_id,
                                               const uint16_t *board, const uint16_t *goals,
                                               ST *stage, int lane, int part, int nparts,
                                               SyncFn sync) {
    // everything the loops need, read once (O and B live in parameter space)
    const int H = B.H, W = B.W, vh = O.vh, vw = O.vw, vcells = vh * vw, C = O.n_channels;
    const int n_runs = O.n_runs;
    const bool raw = C == 0;
    const int y0 = row0 < 0 ? 0 : row0, x0 = row0 < 0 ? 0 : col0;
    const uint32_t white = O.remove_white ? COLORS : 0xffffffffu;   // goal colour to blank
    // address spaces the caller vouches for: lets the compiler use LDS / STS
    if (BSH) __builtin_assume(__isShared(board));
    if (GSH) __builtin_assume(__isShared(goals));
    __builtin_assume(__isShared(stage));
    auto full_at = [&](int idx) -> uint32_t {
        uint32_t g = goals[idx] & COLORS;
        g = g == white ? 0u : g;
        return __byte_perm(board[idx], g, 0x5410);   // board | g << 16
    };
    // channel runs: bits [src, src + len) of the cell go to bits [dst, dst + len) of the
    // mask -- one rotate and one and-or each; the first four runs live in registers
    uint32_t rot[4], sel[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool on = i < n_runs;
        rot[i] = on ? (uint32_t)(O.run_dst[i] - O.run_src[i]) & 31u : 0u;
        sel[i] = on ? O.run_mask[i] << O.run_dst[i] : 0u;
    }
    auto mask4 = [&](uint32_t v) -> uint32_t {
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) m |= __funnelshift_l(v, v, rot[i]) & sel[i];
        return m;
    };
    auto mask_of = [&](uint32_t v) -> uint32_t {     // any number of runs (the rare case)
        uint32_t m = mask4(v);
        for (int r = 4; r < n_runs; r++)
            m |= __funnelshift_l(v, v, (uint32_t)(O.run_dst[r] - O.run_src[r]) & 31u) &
                 (O.run_mask[r] << O.run_dst[r]);
        return m;
    };
    auto stage_st = [&](int i, uint32_t m) { stage[i] = (ST)m; };
    auto stage_ld = [&](int i) -> uint32_t { return stage[i]; };
    uint32_t *out32 = (uint32_t *)O.obs + (size_t)e * vcells;
    // this warp's block of view rows [rb, re)
    const int rb = (int)((long long)vh * part / nparts), re = (int)((long long)vh * (part + 1) / nparts);
    int sy0 = (y0 - vh / 2 + rb) % H;
    if (sy0 < 0) sy0 += H;
    for (int c0 = 0; c0 < vw; c0 += 32) {
        const int c = c0 + lane;
        int sx = (x0 - vw / 2 + c) % W;
        if (sx < 0) sx += W;
        if (c < vw) {
            // rows in runs that end where the torus wraps: no per-row wrap test
            int idx = sy0 * W + sx, o = rb * vw + c, r = rb, run = min(re - rb, H - sy0);
            while (r < re) {
                if (raw) {
                    for (int i = 0; i < run; i++, o += vw, idx += W) out32[o] = full_at(idx);
                } else if (n_runs <= 4) {
#pragma unroll 5
                    for (int i = 0; i < run; i++, o += vw, idx += W) stage_st(o, mask4(full_at(idx)));
                } else {
                    for (int i = 0; i < run; i++, o += vw, idx += W) stage_st(o, mask_of(full_at(idx)));
                }
                r += run;
                idx = sx;                       // wrapped to source row 0
                run = min(re - r, H);
            }
        }
    }
    if (!raw && part == 0) stage_st(vcells + lane, 0u);     // zero padding the expansion may touch
    sync();
    // exit cells outside the view are clamped onto its perimeter; in index order,
    // so that the last one wins where two collide (helper_utils.py:64-74)
    if (lane == 0 && part == 0) {
        const int32_t *exits = B.pool_exits + (size_t)level_id * B.max_exits;
        const int max_exits = B.max_exits;
        for (int k = 0; k < max_exits; k++) {
            const int idx = exits[k];
            if (idx < 0) continue;
            const int iy = idx / W, ix = idx - iy * W;
            int jy = ((iy - y0 + H / 2) % H + H) % H - H / 2 + vh / 2;
            int jx = ((ix - x0 + W / 2) % W + W) % W - W / 2 + vw / 2;
            jy = min(max(jy, 0), vh - 1);
            jx = min(max(jx, 0), vw - 1);
            const uint32_t v = full_at(idx);
            if (raw) out32[jy * vw + jx] = v;
            else stage_st(jy * vw + jx, mask_of(v));
        }
    }
    sync();
    if (raw) return;
    // expansion: output byte b of the env = bit (b % C) of cell (b / C)
    const int total = vcells * C;
    uint8_t *o = (uint8_t *)O.obs + (size_t)e * total;
    const int head = min(total, (int)((16u - (unsigned)((uintptr_t)o & 15u)) & 15u));
    // every lane expands 32 consecutive stream bits into two 16-byte stores per round
    const int head32 = head;
    const int nchunks = (total - head32) >> 5;
    auto bit_at = [&](int b) -> uint8_t {
        const int cell = b / C;
        return (uint8_t)((stage_ld(cell) >> (b - cell * C)) & 1u);
    };
    if (part == 0) {
        if (lane < head32) o[lane] = bit_at(lane);
        const int b = head32 + (nchunks << 5) + lane;      // fewer than 32 bytes are left over
        if (b < total) o[b] = bit_at(b);
    }
    uint4 *o16 = (uint4 *)(o + head32);
    // 32 lanes x 32 bytes per round, rounds dealt out over the parts
    const int round_bytes = 1024 * nparts;
    const int dcell = round_bytes / C, dk = round_bytes - dcell * C;
    const int first = head32 + 32 * (32 * part + lane);
    int cell = first / C, k = first - cell * C;
    // Each lane gathers the 32 stream bits of its own 32-byte span; two shuffles then
    // hand the 16-bit halves out so that each of the round's two store instructions
    // covers 512 contiguous bytes (a lane storing its own two halves would leave every
    // 32-byte sector half written by each instruction).
    const int n16 = 2 * nchunks;
    for (int q0 = 32 * part; q0 < nchunks; q0 += 32 * nparts) {
        uint32_t bits = 0;
        if (q0 + lane < nchunks) {
            if (C >= 11) {     // 32 stream bits span at most four cells
                const uint32_t a0 = stage_ld(cell), a1 = stage_ld(cell + 1), a2 = stage_ld(cell + 2),
                               a3 = stage_ld(cell + 3);
                const int nb = C - k;
                bits = (a0 >> k) | shl_clamp(a1, nb) | shl_clamp(a2, nb + C) | shl_clamp(a3, nb + 2 * C);
            } else {
                bits = stage_ld(cell) >> k;
                int nb = C - k, cc = cell + 1;
                while (nb < 32) {
                    bits |= stage_ld(cc) << nb;
                    nb += C;
                    cc++;
                }
            }
            cell += dcell;
            k += dk;
            if (k >= C) { k -= C; cell++; }
        }
        const uint32_t v1 = __shfl_sync(0xffffffffu, bits, lane >> 1);
        const uint32_t v2 = __shfl_sync(0xffffffffu, bits, 16 + (lane >> 1));
        const int sh = (lane & 1) * 16, c1 = 2 * q0 + lane, c2 = c1 + 32;
        if (c1 < n16) o16[c1] = expand16(v1 >> sh);
        if (c2 < n16) o16[c2] = expand16(v2 >> sh);
    }
}

// The writer for the plane placement at hand, staging type by channel count.
template <bool BSH, bool GSH, class SyncFn>
__device__ __forceinline__ void warp_write_obs_any(const SlEnvBatch &B, const ObsParams &O, int e,
                                                   const SlEnvState *st, const uint16_t *board,
                                                   const uint16_t *goals, unsigned char *stage,
                                                   int lane, int part, int nparts, SyncFn sync) {
    const int row = st->loc_row, col = st->loc_col, lvl = st->level_id;
    if (O.n_channels <= 16)
        warp_write_obs<uint16_t, BSH, GSH>(B, O, e, row, col, lvl, board, goals, (uint16_t *)stage,
                                           lane, part, nparts, sync);
    else
        warp_write_obs<uint32_t, BSH, GSH>(B, O, e, row, col, lvl, board, goals, (uint32_t *)stage,
                                           lane, part, nparts, sync);
}

// Stand-alone SafeLifeEnv.get_obs: one warp per env, planes read from HBM.
constexpr int OBS_WARPS = 8;

__global__ void __launch_bounds__(OBS_WARPS * 32)
k_get_obs_warp(const __grid_constant__ SlEnvBatch B, const __grid_constant__ ObsParams O) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = B.H * B.W;
    unsigned char *stage = smem_raw + (size_t)warp * obs_stage_bytes(O.vh, O.vw, O.n_channels);
    for (int e = blockIdx.x * OBS_WARPS + warp; e < B.n_envs; e += gridDim.x * OBS_WARPS) {
        warp_write_obs_any<false, false>(B, O, e, B.state + e, B.board + (size_t)e * cells,
                                         B.goals + (size_t)e * cells, stage, lane, 0, 1,
                                         [] { __syncwarp(); });
        __syncwarp();
    }
}

}  // namespace sl
