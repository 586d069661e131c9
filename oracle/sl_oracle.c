/*
 * sl_oracle.c -- CPU restatement of the SafeLife step path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under safelife_b200/ may link, import or
 * call this file.  It exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py can check the CUDA path against an independent
 * scalar implementation.  Parity of this restatement is pinned against the
 * compiled reference C extension (oracle/_ref, built by oracle/Makefile from
 * the sources under /root/reference) by tests/golden/make_golden.py and
 * tests/test_oracle_golden.py -- see DESIGN.md "Oracle".
 *
 * Every function states the rule declaratively (direct 3x3 gather, no
 * separable accumulator trick) and cites the reference lines it follows.
 * Paths are relative to the reference checkout (safelife/...).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Cell word bit layout: speedups_src/constants.h:4-33, safelife_game.py:75-123 */
#define C_ALIVE        0x0001u
#define C_AGENT        0x0002u
#define C_PUSHABLE     0x0004u
#define C_DESTRUCTIBLE 0x0008u
#define C_FROZEN       0x0010u
#define C_PRESERVING   0x0020u
#define C_INHIBITING   0x0040u
#define C_SPAWNING     0x0080u
#define C_EXIT         0x0100u
#define C_COLORS       0x0E00u
#define C_ORIENT       0x3000u
#define C_PULLABLE     0x8000u
#define COLOR_SHIFT    9
#define ORIENT_SHIFT   12

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------
 * PCG64 (numpy.random.PCG64, the stream behind safelife_game.py:175-180 and
 * speedups_src/random.c:76-83): 128-bit LCG, XSL-RR output of the post-step
 * state, double = (out >> 11) * 2^-53.
 * rng[4] = {state_hi, state_lo, inc_hi, inc_lo}.
 * ---------------------------------------------------------------------- */
static const uint64_t PCG_MULT_HI = 0x2360ED051FC65DA4ull;
static const uint64_t PCG_MULT_LO = 0x4385DF649FCCF645ull;

static inline uint64_t pcg64_next_u64(uint64_t *rng) {
    u128 state = ((u128)rng[0] << 64) | rng[1];
    u128 inc = ((u128)rng[2] << 64) | rng[3];
    u128 mult = ((u128)PCG_MULT_HI << 64) | PCG_MULT_LO;
    state = state * mult + inc;
    rng[0] = (uint64_t)(state >> 64);
    rng[1] = (uint64_t)state;
    uint64_t x = rng[0] ^ rng[1];
    unsigned rot = (unsigned)(rng[0] >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}

static inline double pcg64_next_double(uint64_t *rng) {
    return (double)(pcg64_next_u64(rng) >> 11) * (1.0 / 9007199254740992.0);
}

uint64_t orc_pcg64_next_u64(uint64_t *rng) { return pcg64_next_u64(rng); }
double orc_pcg64_next_double(uint64_t *rng) { return pcg64_next_double(rng); }

static inline int wrap(int v, int n) {
    if (v < 0) return v + n;
    if (v >= n) return v - n;
    return v;
}

/* ------------------------------------------------------------------------
 * One Life-variant step on a torus.  advance_board.c:34-125.
 * `in` is not modified; `out` receives the new board.  One RNG draw per
 * cell that is dead, not frozen, not inhibited, has n != 3 and a spawner in
 * its 3x3, in raster order (advance_board.c:115, short-circuit &&).
 * spawn_prob is a C float compared against a double draw (module.c:23).
 * Returns the number of draws taken.
 * ---------------------------------------------------------------------- */
int orc_advance_board(const uint16_t *in, uint16_t *out, int H, int W,
                      float spawn_prob, uint64_t *rng) {
    int draws = 0;
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            uint16_t self = in[y * W + x];
            int n = 0;                 /* alive cells in the 3x3, centre included */
            unsigned any = 0;          /* OR of PRESERVING|INHIBITING|SPAWNING */
            int ccount[3] = {0, 0, 0}; /* alive cells carrying each colour bit */
            int dcount = 0;            /* alive cells that are DESTRUCTIBLE or EXIT (:46) */
            unsigned spawn_colors = 0; /* colours of spawners in the 3x3 (:22) */
            for (int dy = -1; dy <= 1; dy++) {
                int yy = wrap(y + dy, H);
                for (int dx = -1; dx <= 1; dx++) {
                    int xx = wrap(x + dx, W);
                    uint16_t c = in[yy * W + xx];
                    any |= c & (C_PRESERVING | C_INHIBITING | C_SPAWNING);
                    if (c & C_SPAWNING) spawn_colors |= c & C_COLORS;
                    if (c & C_ALIVE) {
                        n++;
                        for (int k = 0; k < 3; k++)
                            if (c & (1u << (COLOR_SHIFT + k))) ccount[k]++;
                        if (c & (C_DESTRUCTIBLE | C_EXIT)) dcount++;
                    }
                }
            }
            unsigned inherit = spawn_colors;
            for (int k = 0; k < 3; k++)
                if (ccount[k] >= 2) inherit |= 1u << (COLOR_SHIFT + k);
            uint16_t res = self;
            if (self & C_ALIVE) {
                /* :96-105 */
                if (!((self & C_FROZEN) || (any & C_PRESERVING) || n == 3 || n == 4))
                    res = 0;
            } else if ((self & C_FROZEN) || (any & C_INHIBITING)) {
                /* :108-109 unchanged */
            } else if (n == 3) {
                /* :110-114 */
                res = (uint16_t)(C_ALIVE | inherit | (dcount >= 2 ? C_DESTRUCTIBLE : 0));
            } else if (any & C_SPAWNING) {
                /* :115-118 */
                draws++;
                if (pcg64_next_double(rng) < (double)spawn_prob)
                    res = (uint16_t)(C_ALIVE | C_DESTRUCTIBLE | inherit);
            }
            out[y * W + x] = res;
        }
    }
    return draws;
}

/* n >= 1 successive steps (advance_board.c:128-149; n=0 is a reference quirk
 * that returns scratch contents and is not part of the contract). */
int orc_advance_board_nstep(const uint16_t *in, uint16_t *out, int H, int W,
                            float spawn_prob, int n_steps, uint64_t *rng) {
    if (n_steps < 1) return -1;
    size_t bytes = (size_t)H * W * sizeof(uint16_t);
    uint16_t *a = malloc(bytes), *b = malloc(bytes);
    memcpy(a, in, bytes);
    for (int s = 0; s < n_steps; s++) {
        orc_advance_board(a, b, H, W, spawn_prob, rng);
        uint16_t *t = a; a = b; b = t;
    }
    memcpy(out, a, bytes);
    free(a); free(b);
    return 0;
}

/* advance_board.c:153-189: advance n steps; after each, count cells that are
 * ALIVE and none of AGENT|EXIT|FROZEN into counts[y][x][colour]. */
int orc_life_occupancy(const uint16_t *in, int32_t *counts, int H, int W,
                       float spawn_prob, int n_steps, uint64_t *rng) {
    if (n_steps < 1) return -1;
    size_t cells = (size_t)H * W;
    uint16_t *a = malloc(cells * 2), *b = malloc(cells * 2);
    memcpy(a, in, cells * 2);
    for (int s = 0; s < n_steps; s++) {
        orc_advance_board(a, b, H, W, spawn_prob, rng);
        for (size_t i = 0; i < cells; i++) {
            uint16_t c = b[i];
            if ((c & C_ALIVE) && !(c & (C_AGENT | C_EXIT | C_FROZEN)))
                counts[8 * i + ((c & C_COLORS) >> COLOR_SHIFT)]++;
        }
        uint16_t *t = a; a = b; b = t;
    }
    free(a); free(b);
    return 0;
}

/* advance_board.c:192-207: out[goal_colour][cell_colour or 8] histogram over
 * cells an agent could change (movable, or not frozen).  `out` is added to. */
void orc_alive_counts(const uint16_t *board, const uint16_t *goals, int n,
                      int64_t *out) {
    for (int i = 0; i < n; i++) {
        uint16_t b = board[i];
        if (!((b & (C_DESTRUCTIBLE | C_PUSHABLE | C_PULLABLE)) || !(b & C_FROZEN)))
            continue;
        int g = (goals[i] & C_COLORS) >> COLOR_SHIFT;
        int col = (b & C_ALIVE) ? ((b & C_COLORS) >> COLOR_SHIFT) : 8;
        out[g * 9 + col] += 1;
    }
}

/* advance_board.c:217-300.  locations are (row, col) pairs, updated in place;
 * agents act in index order on the shared board. action_stride 0 broadcasts
 * one action (module.c:198). */
void orc_execute_actions(uint16_t *board, int H, int W, int64_t *locations,
                         const int64_t *actions, int n_agents, int action_stride) {
    for (int k = 0; k < n_agents; k++) {
        int64_t action = actions[(size_t)k * action_stride];
        if (action == 0) continue;
        int dir = (int)((action - 1) & 3);  /* 0 up, 1 right, 2 down, 3 left */
        int dy = (dir == 0) ? -1 : (dir == 2) ? 1 : 0;
        int dx = (dir == 1) ? 1 : (dir == 3) ? -1 : 0;
        int y0 = (int)(locations[2 * k] % H);
        int x0 = (int)(locations[2 * k + 1] % W);
        int i0 = y0 * W + x0;
        int y1 = wrap(y0 + dy, H), x1 = wrap(x0 + dx, W);
        int i1 = y1 * W + x1;
        int i2 = wrap(y1 + dy, H) * W + wrap(x1 + dx, W);
        int i3 = wrap(y0 - dy, H) * W + wrap(x0 - dx, W);
        if (!(board[i0] & C_AGENT)) continue;                       /* :241 */
        board[i0] = (uint16_t)((board[i0] & ~C_ORIENT) | (dir << ORIENT_SHIFT));
        int agent_can_push = !(board[i0] & C_PUSHABLE);
        if (action >= 5) {                                           /* toggle */
            uint16_t front = board[i1];
            if (front == 0) {
                board[i1] = (uint16_t)(C_ALIVE | C_DESTRUCTIBLE | (board[i0] & C_COLORS));
            } else if (front & C_DESTRUCTIBLE) {
                if (front & C_AGENT)
                    board[i1] = (uint16_t)((front ^ (C_AGENT | C_DESTRUCTIBLE)) | C_FROZEN);
                else
                    board[i1] = 0;
            } else if (agent_can_push && (front & C_PUSHABLE)) {
                if (board[i2] == 0) { board[i2] = front; board[i1] = 0; }
                else if (board[i2] & C_EXIT) board[i1] = 0;
            }
            continue;
        }
        /* move */
        int moved = 0, onto_exit = 0;
        uint16_t front = board[i1];
        if (agent_can_push && (front & C_PUSHABLE)) {
            if (board[i2] == 0) { board[i2] = front; moved = 1; }
            else if (board[i2] & C_EXIT) moved = 1;
        } else if (front == 0) {
            moved = 1;
        } else if ((board[i0] & front & C_EXIT) && !(front & C_AGENT)) {
            onto_exit = 1;
        }
        if (!moved && !onto_exit) continue;
        if (moved) board[i1] = board[i0];
        locations[2 * k] = y1;
        locations[2 * k + 1] = x1;
        if (!(board[i0] & C_PULLABLE) && (board[i3] & C_PULLABLE)) {  /* :292 */
            board[i0] = board[i3];
            board[i3] = 0;
        } else {
            board[i0] = 0;
        }
    }
}

/* ------------------------------------------------------------------------
 * wrapped_label.c:5-55 with the index-set of iset.c:6-61: 8-connected
 * toroidal labelling whose label NUMBERING follows the order in which the
 * "visited" index set hands out unvisited cells.  That order depends on the
 * swap-to-front bookkeeping, so the structure is emulated: order[] is a
 * permutation of cell indices, the first `nvis` entries are visited, pos[] is
 * its inverse; marking a cell visited swaps it with order[nvis].
 * ---------------------------------------------------------------------- */
int orc_wrapped_label(int32_t *data, int H, int W) {
    int size = H * W, nvis = 0, label = 0, top;
    int *order = malloc(sizeof(int) * size), *pos = malloc(sizeof(int) * size);
    int *stack = malloc(sizeof(int) * size);
    for (int i = 0; i < size; i++) order[i] = pos[i] = i;
    while (nvis < size) {
        int seed = order[nvis++];          /* wrapped_label.c:15 */
        if (!data[seed]) continue;
        data[seed] = ++label;
        top = 0;
        stack[top++] = seed;
        while (top > 0) {
            int cur = stack[--top];
            int r = cur / W, c = cur % W;
            for (int dy = -1; dy <= 1; dy++) {
                int y = wrap(r + dy, H);
                for (int dx = -1; dx <= 1; dx++) {
                    int x = wrap(c + dx, W);
                    int j = y * W + x;
                    int seen = pos[j] < nvis;
                    if (data[j] > 0 && !seen) {
                        data[j] = label;
                        stack[top++] = j;
                    }
                    if (!seen) {           /* iset_add(&visited, j) */
                        int k = pos[j], other = order[nvis];
                        order[nvis] = j; pos[j] = nvis;
                        order[k] = other; pos[other] = k;
                        nvis++;
                    }
                }
            }
        }
    }
    free(order); free(pos); free(stack);
    return label;
}

/* ------------------------------------------------------------------------
 * Batched single-agent env step: SafeLifeEnv.step (safelife_env.py:148-201)
 * over SafeLifeGame (safelife_game.py:380-389, 505-552, 657-719, 746-761).
 *
 * Per-env state (all arrays indexed by env e):
 *   board, goals  uint16[H*W]
 *   loc           int32[2]   (row, col) of agent 0; has_agent=0 if the level has none
 *   rng           uint64[4]
 *   table         int32[72]  points_table[0] (rows: goal colour, cols: cell colour / empty)
 *   exits         int32[max_exits] flat cell indices of static exit cells, -1 padded
 *   init_points   int32  sum(table * initial_counts)
 *   required      int32  required_points()
 *   static_goals  uint8  0 unknown (None), 1 static, 2 dynamic
 *   num_steps, old_value, active, ep_length, ep_reward
 * Outputs: reward f32, done u8, success u8, times_up u8.
 * ---------------------------------------------------------------------- */
typedef struct {
    int n_envs, H, W, max_exits, time_limit;
    uint16_t *board, *goals;
    int32_t *loc;
    uint8_t *has_agent;
    uint64_t *rng;
    const int32_t *table;
    const int32_t *exits;
    const int32_t *init_points;
    const int32_t *required;
    const float *spawn_prob;
    uint8_t *static_goals;
    int32_t *num_steps;
    int32_t *old_value;
    uint8_t *active;
    int32_t *ep_length;
    float *ep_reward;
    float *reward;
    uint8_t *done, *success, *times_up;
} orc_env_batch;

static int32_t table_points(const uint16_t *board, const uint16_t *goals, int n,
                            const int32_t *table) {
    int64_t counts[72];
    memset(counts, 0, sizeof counts);
    orc_alive_counts(board, goals, n, counts);
    int64_t s = 0;
    for (int i = 0; i < 72; i++) s += counts[i] * table[i];
    return (int32_t)s;
}

/* GameWithGoals.update_exit_colors + can_exit (safelife_game.py:537-552,
 * 716-719).  count_points = sum(table*alive_counts) of the cached counts. */
static void update_exit_colors(orc_env_batch *B, int e, int32_t count_points) {
    int cells = B->H * B->W;
    uint16_t *board = B->board + (size_t)e * cells;
    int can_exit = 0;
    if (B->has_agent[e]) {
        int idx = B->loc[2 * e] * B->W + B->loc[2 * e + 1];
        uint16_t a = board[idx];
        int exited = (a & (C_AGENT | C_EXIT)) == C_EXIT;
        int32_t earned = count_points - B->init_points[e] + exited;
        if (earned < 0) earned = 0;
        can_exit = (a & C_AGENT) && earned >= B->required[e];
        board[idx] = (uint16_t)((a & ~C_EXIT) | (can_exit ? C_EXIT : 0));
    }
    uint16_t exit_type = (uint16_t)(C_EXIT | C_FROZEN | (can_exit ? (1u << COLOR_SHIFT) : 0));
    for (int k = 0; k < B->max_exits; k++) {
        int idx = B->exits[(size_t)e * B->max_exits + k];
        if (idx >= 0) board[idx] = exit_type;
    }
}

static int32_t current_points(orc_env_batch *B, int e, int32_t count_points) {
    if (!B->has_agent[e]) return count_points;
    int cells = B->H * B->W;
    uint16_t a = B->board[(size_t)e * cells + B->loc[2 * e] * B->W + B->loc[2 * e + 1]];
    return count_points + (((a & (C_AGENT | C_EXIT)) == C_EXIT) ? 1 : 0);
}

/* SafeLifeEnv.reset after the level data is in place (safelife_env.py:203-218):
 * counts are taken BEFORE update_exit_colors (cached property, :657-663). */
void orc_env_reset(orc_env_batch *B, int e) {
    int cells = B->H * B->W;
    int32_t cp = table_points(B->board + (size_t)e * cells, B->goals + (size_t)e * cells,
                              cells, B->table + (size_t)e * 72);
    update_exit_colors(B, e, cp);
    B->old_value[e] = current_points(B, e, cp);
    B->active[e] = 1;
    B->ep_length[e] = 0;
    B->ep_reward[e] = 0.0f;
    B->num_steps[e] = 0;
    B->static_goals[e] = 0;
}

void orc_env_step(orc_env_batch *B, const int64_t *actions) {
    int cells = B->H * B->W;
    uint16_t *tmp = malloc(sizeof(uint16_t) * cells);
    for (int e = 0; e < B->n_envs; e++) {
        uint16_t *board = B->board + (size_t)e * cells;
        uint16_t *goals = B->goals + (size_t)e * cells;
        uint64_t *rng = B->rng + 4 * (size_t)e;
        /* safelife_env.py:151 */
        if (B->has_agent[e]) {
            int64_t loc[2] = {B->loc[2 * e], B->loc[2 * e + 1]};
            orc_execute_actions(board, B->H, B->W, loc, actions + e, 1, 1);
            B->loc[2 * e] = (int32_t)loc[0];
            B->loc[2 * e + 1] = (int32_t)loc[1];
        }
        /* :152 -> safelife_game.py:746-761 */
        B->num_steps[e] += 1;
        orc_advance_board(board, tmp, B->H, B->W, B->spawn_prob[e], rng);
        memcpy(board, tmp, sizeof(uint16_t) * cells);
        if (B->static_goals[e] != 1) {
            orc_advance_board(goals, tmp, B->H, B->W, B->spawn_prob[e], rng);
            if (B->static_goals[e] == 0) {
                int is_static = 1;
                for (int i = 0; i < cells; i++)
                    if ((tmp[i] & C_SPAWNING) || tmp[i] != goals[i]) { is_static = 0; break; }
                B->static_goals[e] = is_static ? 1 : 2;
            }
            memcpy(goals, tmp, sizeof(uint16_t) * cells);
        }
        /* :153 */
        int32_t cp = table_points(board, goals, cells, B->table + (size_t)e * 72);
        update_exit_colors(B, e, cp);
        /* :155-160 */
        int times_up = B->num_steps[e] >= B->time_limit;
        int32_t new_value = current_points(B, e, cp);
        float reward;
        int done, success;
        if (B->has_agent[e]) {
            uint16_t a = board[B->loc[2 * e] * B->W + B->loc[2 * e + 1]];
            reward = (float)((new_value - B->old_value[e]) * (B->active[e] ? 1 : 0));
            success = (a & (C_AGENT | C_EXIT)) == C_EXIT;
            done = !(a & C_AGENT) || times_up;
        } else {   /* :163-166 */
            reward = 0.0f; done = 1; success = 0;
        }
        B->old_value[e] = new_value;
        B->ep_reward[e] += reward;
        B->ep_length[e] += B->active[e] ? 1 : 0;
        B->active[e] = B->active[e] && !done;
        B->reward[e] = reward;
        B->done[e] = (uint8_t)done;
        B->success[e] = (uint8_t)success;
        B->times_up[e] = (uint8_t)times_up;
    }
    free(tmp);
}

/* ------------------------------------------------------------------------
 * SafeLifeEnv.get_obs (safelife_env.py:105-146) + recenter_view
 * (helper_utils.py:42-75) for agent 0 (origin when the level has no agent).
 * out: uint32[vh*vw] packed view (board | goal colours << 16).
 * ---------------------------------------------------------------------- */
void orc_get_view(const uint16_t *board, const uint16_t *goals, int H, int W,
                  int y0, int x0, const int32_t *exits, int n_exits,
                  int vh, int vw, int remove_white_goals, uint32_t *out) {
    for (int r = 0; r < vh; r++) {
        int y = ((y0 - vh / 2 + r) % H + H) % H;
        for (int c = 0; c < vw; c++) {
            int x = ((x0 - vw / 2 + c) % W + W) % W;
            uint32_t g = goals[y * W + x] & C_COLORS;
            if (remove_white_goals && g == C_COLORS) g = 0;
            out[r * vw + c] = (uint32_t)board[y * W + x] | (g << 16);
        }
    }
    for (int k = 0; k < n_exits; k++) {
        int idx = exits[k];
        if (idx < 0) continue;
        int iy = idx / W, ix = idx % W;
        int jy = ((iy - y0 + H / 2) % H + H) % H - H / 2 + vh / 2;
        int jx = ((ix - x0 + W / 2) % W + W) % W - W / 2 + vw / 2;
        if (jy < 0) jy = 0; if (jy > vh - 1) jy = vh - 1;
        if (jx < 0) jx = 0; if (jx > vw - 1) jx = vw - 1;
        uint32_t g = goals[idx] & C_COLORS;
        if (remove_white_goals && g == C_COLORS) g = 0;
        out[jy * vw + jx] = (uint32_t)board[idx] | (g << 16);
    }
}

/* SimpleSideEffectPenalty.step with ignore_reward_cells=False
 * (env_wrappers.py:177-213): number of cells where the board differs from the
 * baseline once player bits (122) are masked out, exit cells excluded. */
int orc_side_effect_cells(const uint16_t *board, const uint16_t *baseline, int cells,
                          const int32_t *exits, int n_exits) {
    const uint16_t keep = (uint16_t)~(C_AGENT | C_DESTRUCTIBLE | C_FROZEN | C_PRESERVING | C_INHIBITING);
    int n = 0;
    for (int i = 0; i < cells; i++) {
        int is_exit = 0;
        for (int k = 0; k < n_exits; k++) if (exits[k] == i) is_exit = 1;
        if (!is_exit && ((board[i] ^ baseline[i]) & keep)) n++;
    }
    return n;
}
