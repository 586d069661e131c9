// cells.cuh -- cell word layout, PCG64 arithmetic and small device helpers
// shared by every kernel of libsafelife_b200.
//
// Cell word: safelife/speedups_src/constants.h:4-33 (reference checkout),
// mirrored by CellTypes in safelife/safelife_game.py:75-123.
#pragma once
#include <stdint.h>
#include "pcg_tables.h"

namespace sl {

enum : uint32_t {
    ALIVE = 1u << 0,
    AGENT = 1u << 1,
    PUSHABLE = 1u << 2,
    DESTRUCTIBLE = 1u << 3,
    FROZEN = 1u << 4,
    PRESERVING = 1u << 5,
    INHIBITING = 1u << 6,
    SPAWNING = 1u << 7,
    EXIT = 1u << 8,
    COLORS = 7u << 9,
    COLOR_R = 1u << 9,
    ORIENT = 3u << 12,
    PULLABLE = 1u << 15,
    LEVEL_EXIT = EXIT | FROZEN,   // CellTypes.level_exit = 272
};
constexpr int COLOR_SHIFT = 9;
constexpr int ORIENT_SHIFT = 12;

// ---------------------------------------------------------------- 128-bit
struct u128 {
    uint64_t hi, lo;
};

__host__ __device__ inline u128 mul128(u128 a, u128 b) {
#ifdef __CUDA_ARCH__
    uint64_t hi = __umul64hi(a.lo, b.lo);
#else
    uint64_t hi = (uint64_t)(((unsigned __int128)a.lo * b.lo) >> 64);
#endif
    return {hi + a.hi * b.lo + a.lo * b.hi, a.lo * b.lo};
}

__host__ __device__ inline u128 add128(u128 a, u128 b) {
    uint64_t lo = a.lo + b.lo;
    return {a.hi + b.hi + (lo < a.lo ? 1ull : 0ull), lo};
}

// PCG64 (numpy.random.PCG64; speedups_src/random.c:76-83 calls its
// next_double): XSL-RR output of the state AFTER the LCG step.
struct Pcg {
    u128 state, inc;
};

__device__ inline Pcg pcg_load(const uint64_t *p) {
    return {{p[0], p[1]}, {p[2], p[3]}};
}

__device__ inline uint64_t pcg_output(u128 s) {
    uint64_t x = s.hi ^ s.lo;
    unsigned rot = (unsigned)(s.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

__device__ inline void pcg_step(Pcg &g) {
    g.state = add128(mul128(g.state, {SL_PCG_MULT_HI, SL_PCG_MULT_LO}), g.inc);
}

// draw < spawn_prob  <=>  (u64 >> 11) < thr, thr = ceil((double)(float)p * 2^53)
__device__ inline bool pcg_next_below(Pcg &g, uint64_t thr) {
    pcg_step(g);
    return (pcg_output(g.state) >> 11) < thr;
}

__constant__ uint64_t c_pow2_a[32][2] = SL_PCG_POW2_A;
__constant__ uint64_t c_pow2_g[32][2] = SL_PCG_POW2_G;
__constant__ uint64_t c_lane_a[32][2] = SL_PCG_LANE_A;
__constant__ uint64_t c_lane_g[32][2] = SL_PCG_LANE_G;

// state after k more steps: A^k * s + G_k * inc, by binary decomposition of k.
__device__ inline u128 pcg_jump(u128 state, u128 inc, uint32_t k) {
    for (int j = 0; k != 0; j++, k >>= 1) {
        if (k & 1u) {
            u128 a = {c_pow2_a[j][0], c_pow2_a[j][1]};
            u128 g = {c_pow2_g[j][0], c_pow2_g[j][1]};
            state = add128(mul128(a, state), mul128(g, inc));
        }
    }
    return state;
}

// ------------------------------------------------------------- cell rule
// Aggregate of the 3x3 neighbourhood (centre included) that the rule needs;
// declarative form of advance_board.c:12-87.
struct Hood {
    int n;             // alive cells
    uint32_t flags;    // OR of PRESERVING | INHIBITING | SPAWNING
    uint32_t inherit;  // colour bits: >=2 alive carriers, or any spawner carrying it
    bool destructible; // >=2 alive cells that are DESTRUCTIBLE or EXIT
};

__device__ inline int wrap_idx(int v, int n) {
    return v < 0 ? v + n : (v >= n ? v - n : v);
}

template <typename Board>
__device__ inline Hood gather_hood(const Board &cell_at, int y, int x, int H, int W) {
    int n = 0, cr = 0, cg = 0, cb = 0, cd = 0;
    uint32_t flags = 0, spawn_colors = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
        int yy = wrap_idx(y + dy, H);
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            int xx = wrap_idx(x + dx, W);
            uint32_t c = cell_at(yy * W + xx);
            flags |= c;
            if (c & SPAWNING) spawn_colors |= c;
            uint32_t a = c & ALIVE;
            n += a;
            cr += a & (c >> 9);
            cg += a & (c >> 10);
            cb += a & (c >> 11);
            cd += a & ((c >> 3) | (c >> 8));
        }
    }
    Hood h;
    h.n = n;
    h.flags = flags & (PRESERVING | INHIBITING | SPAWNING);
    h.inherit = (spawn_colors & COLORS) | (cr >= 2 ? 1u << 9 : 0) | (cg >= 2 ? 1u << 10 : 0) |
                (cb >= 2 ? 1u << 11 : 0);
    h.destructible = cd >= 2;
    return h;
}

// Result of the rule for one cell given its neighbourhood.  When the cell
// qualifies for a spawn draw, *draws is set and the returned value is the
// cell as it would be spawned; the caller keeps `self` if the draw fails.
// advance_board.c:92-124.
__device__ inline uint32_t apply_rule(uint32_t self, const Hood &h, bool *draws) {
    *draws = false;
    if (self & ALIVE) {
        bool keep = (self & FROZEN) || (h.flags & PRESERVING) || h.n == 3 || h.n == 4;
        return keep ? self : 0u;
    }
    if ((self & FROZEN) || (h.flags & INHIBITING)) return self;
    if (h.n == 3) return ALIVE | h.inherit | (h.destructible ? (uint32_t)DESTRUCTIBLE : 0u);
    if (h.flags & SPAWNING) {
        *draws = true;
        return ALIVE | DESTRUCTIBLE | h.inherit;
    }
    return self;
}

// alive_counts bin of one cell, or -1 if the cell is not counted
// (advance_board.c:192-207).
__device__ inline int count_bin(uint32_t b, uint32_t g) {
    if (!((b & (DESTRUCTIBLE | PUSHABLE | PULLABLE)) || !(b & FROZEN))) return -1;
    int gc = (g & COLORS) >> COLOR_SHIFT;
    int col = (b & ALIVE) ? (int)((b & COLORS) >> COLOR_SHIFT) : 8;
    return gc * 9 + col;
}

// One agent's action on a board held in (shared or global) memory.
// advance_board.c:217-300; (y0, x0) is updated when the agent moves.
// The cells an action can touch: the agent's (i0), the one ahead (i1, at
// (y1, x1)), two ahead (i2) and the one behind (i3), on the torus.
struct ActionCells {
    int dir, i0, i1, i2, i3, y1, x1;
};

__device__ inline ActionCells action_cells(int H, int W, int y0, int x0, int action) {
    ActionCells c;
    c.dir = (action - 1) & 3;   // 0 up, 1 right, 2 down, 3 left
    // all four cells lie on one axis: walk it once
    const bool vertical = (c.dir & 1) == 0;
    const int step = (c.dir == 1 || c.dir == 2) ? 1 : -1;
    const int n = vertical ? H : W, pos = vertical ? y0 : x0;
    const int stride = vertical ? W : 1, base = vertical ? x0 : y0 * W;
    int p1 = pos + step;
    p1 = p1 < 0 ? n - 1 : (p1 == n ? 0 : p1);
    int p2 = p1 + step;
    p2 = p2 < 0 ? n - 1 : (p2 == n ? 0 : p2);
    int p3 = pos - step;
    p3 = p3 < 0 ? n - 1 : (p3 == n ? 0 : p3);
    c.i0 = y0 * W + x0;
    c.i1 = base + p1 * stride;
    c.i2 = base + p2 * stride;
    c.i3 = base + p3 * stride;
    c.y1 = vertical ? p1 : y0;
    c.x1 = vertical ? x0 : p1;
    return c;
}

template <typename T>
__device__ inline void agent_action(T *board, int H, int W, int &y0, int &x0, int action);

template <typename T>
__device__ inline void agent_action_at(T *board, const ActionCells &c, int &y0, int &x0, int action) {
    const int dir = c.dir, i0 = c.i0, i1 = c.i1, i2 = c.i2, i3 = c.i3, y1 = c.y1, x1 = c.x1;
    uint32_t me = board[i0];
    if (!(me & AGENT)) return;
    me = (me & ~(uint32_t)ORIENT) | ((uint32_t)dir << ORIENT_SHIFT);
    board[i0] = (T)me;
    uint32_t front = board[i1];
    bool can_push = !(me & PUSHABLE) && (front & PUSHABLE);
    if (action >= 5) {   // toggle
        if (front == 0) {
            board[i1] = (T)(ALIVE | DESTRUCTIBLE | (me & COLORS));
        } else if (front & DESTRUCTIBLE) {
            board[i1] = (front & AGENT) ? (T)((front ^ (AGENT | DESTRUCTIBLE)) | FROZEN) : (T)0;
        } else if (can_push) {
            uint32_t beyond = board[i2];
            if (beyond == 0) {
                board[i2] = (T)front;
                board[i1] = 0;
            } else if (beyond & EXIT) {
                board[i1] = 0;
            }
        }
        return;
    }
    bool moved = false, onto_exit = false;
    if (can_push) {
        uint32_t beyond = board[i2];
        if (beyond == 0) {
            board[i2] = (T)front;
            moved = true;
        } else if (beyond & EXIT) {
            moved = true;
        }
    } else if (front == 0) {
        moved = true;
    } else if ((me & front & EXIT) && !(front & AGENT)) {
        onto_exit = true;
    }
    if (!moved && !onto_exit) return;
    if (moved) board[i1] = (T)me;
    y0 = y1;
    x0 = x1;
    uint32_t behind = board[i3];
    if (!(me & PULLABLE) && (behind & PULLABLE)) {
        board[i0] = (T)behind;
        board[i3] = 0;
    } else {
        board[i0] = 0;
    }
}

template <typename T>
__device__ inline void agent_action(T *board, int H, int W, int &y0, int &x0, int action) {
    if (action == 0) return;
    agent_action_at(board, action_cells(H, W, y0, x0, action), y0, x0, action);
}

}  // namespace sl
