/*
 * safelife_b200.h -- C ABI of libsafelife_b200.so, the B200 (sm_100a) batched
 * SafeLife step engine.
 *
 * Every entry point is plain C: raw pointers, sizes, an opaque cudaStream_t
 * (passed as void*, NULL = the legacy default stream), an int status return
 * (SL_OK = 0).  No function throws, allocates caller-visible memory or keeps
 * global simulation state; sl_last_error() returns a thread-local message for
 * the last non-zero status.
 *
 * Each declaration cites the reference interface it replaces; paths are
 * relative to the PartnershipOnAI/safelife checkout (v1.2.2).  The reference
 * binding (safelife/speedups_src/module.c) works on one numpy board per call;
 * the entry points here take a batch of `n` boards laid out [n, H, W] uint16,
 * C-contiguous, row 0 = top -- the SafeLifeGame board layout
 * (safelife/safelife_game.py:75-123, speedups_src/constants.h:4-33) with a
 * leading batch axis.
 *
 * Pointers named *_dev / documented "device" must be device memory of the
 * current CUDA device; the `_host` variants take host memory, stage it through
 * the device and synchronise before returning.
 *
 * RNG: where the reference reads a process-global numpy bit generator
 * (speedups_src/random.c:21-39, 76-83) these functions take the PCG64 state
 * explicitly as uint64[n][4] = {state_hi, state_lo, inc_hi, inc_lo} per board
 * and advance it in place by exactly the number of doubles the reference
 * would have drawn (one per qualifying cell in raster order,
 * speedups_src/advance_board.c:115).  rng may be NULL only if no board of the
 * batch can draw (no SPAWNING cell); a NULL rng with spawners present leaves
 * every draw "not spawned".
 */
#ifndef SAFELIFE_B200_H
#define SAFELIFE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SL_OK            0
#define SL_ERR_ARG       1   /* invalid argument (the reference raises ValueError) */
#define SL_ERR_CUDA      2   /* CUDA runtime error, see sl_last_error() */
#define SL_ERR_NO_DEVICE 3   /* no usable CUDA device: there is no CPU fallback */

#define SL_ABI_VERSION   3

/* ---- library plumbing ------------------------------------------------ */
int sl_abi_version(void);
const char *sl_last_error(void);
/* SM count / L2 bytes of the current device (for grid sizing by callers). */
int sl_device_info(int *sm_count, int *l2_bytes, int *cc_major, int *cc_minor);
/* 1 if the warp-per-env bit-plane kernels handle H x W boards, else 0
 * (larger boards run on the generic CTA-per-env kernels). */
int sl_warp_kernel_supported(int H, int W);

/* ---- board operators (device pointers) ------------------------------- */

/* n_steps >= 1 Life-variant steps on each of n toroidal boards.
 * Replaces speedups.advance_board (module.c:20-49 -> advance_board.c:34-149).
 * in and out may alias.  H, W >= 3.  spawn_prob is the C float the reference
 * parses (module.c:23); it is compared with a double draw. */
int sl_advance_board(const uint16_t *in, uint16_t *out, int n, int H, int W,
                     float spawn_prob, int n_steps, uint64_t *rng, void *stream);

/* Advance n_steps and accumulate, after every step, +1 into
 * counts[b][y][x][colour] for cells ALIVE and none of AGENT|EXIT|FROZEN.
 * counts is int32[n][H][W][8] and is ADDED to (the reference returns a fresh
 * zero-initialised array: callers zero it first).  The boards are not
 * modified.  Replaces speedups.life_occupancy (module.c:52-81 ->
 * advance_board.c:153-189). */
int sl_life_occupancy(const uint16_t *board, int32_t *counts, int n, int H, int W,
                      float spawn_prob, int n_steps, uint64_t *rng, void *stream);

/* The simulation part of side_effects.side_effect_score for n finished
 * episodes at once (side_effects.py:101-111, num_runs = 1):
 *     b1 = advance_board(init_board[b], spawn_prob, num_steps[b])
 *     counts[b][0] += life_occupancy(b1, spawn_prob, num_samples)
 *     counts[b][1] += life_occupancy(final_board[b], spawn_prob, num_samples)
 * with the three calls drawing from rng[b] in this order (what the reference
 * does under `with set_rng(...)`).  counts is int32[n][2][H][W][8], added to;
 * spawn_thr[b] = ceil((double)(float)spawn_prob * 2^53) (SlEnvState.spawn_thr);
 * num_steps[b] >= 0 is GameState.num_steps of the episode (0 leaves the initial
 * board as it is; the reference's n_step = 0 quirk, advance_board.c:128-149,
 * is not reproduced).  The earth-mover distance over these histograms stays on
 * the host (sl_emd_host). */
int sl_side_effect_occupancy(const uint16_t *init_board, const uint16_t *final_board,
                             int32_t *counts, int n, int H, int W,
                             const uint64_t *spawn_thr, const int32_t *num_steps,
                             int num_samples, uint64_t *rng, void *stream);

/* out[b][goal_colour][cell_colour | 8=empty] histogram, int64[n][8][9],
 * overwritten.  Replaces speedups.alive_counts (module.c:99-132 ->
 * advance_board.c:192-207). */
int sl_alive_counts(const uint16_t *board, const uint16_t *goals, int64_t *out,
                    int n, int cells, void *stream);

/* Agents act in index order on their board, in place.  locations is
 * int64[n][n_agents][2] (row, col), updated in place; actions is
 * int64[n][n_agents] (action_stride 1) or int64[n][1] (action_stride 0, one
 * action broadcast to all agents of the board).  Replaces
 * speedups.execute_actions (module.c:155-202 -> advance_board.c:217-300). */
int sl_execute_actions(uint16_t *board, int64_t *locations, const int64_t *actions,
                       int n, int H, int W, int n_agents, int action_stride,
                       void *stream);

/* 8-connected toroidal component labelling of the non-zero cells of each
 * int32[H][W] grid, in place; num_labels[b] receives the label count.  Label
 * numbering follows the reference's index-set visiting order.  Replaces
 * speedups.wrapped_label (module.c:223-249 -> wrapped_label.c:5-55). */
int sl_wrapped_label(int32_t *data, int32_t *num_labels, int n, int H, int W,
                     void *stream);

/* ---- the same operators on host buffers ------------------------------ */
/* What a maintainer of module.c binds in place of the CPU loops: same
 * arguments, host memory, one board batch per call, synchronous. */
int sl_advance_board_host(const uint16_t *in, uint16_t *out, int n, int H, int W,
                          float spawn_prob, int n_steps, uint64_t *rng);
int sl_life_occupancy_host(const uint16_t *board, int32_t *counts, int n, int H, int W,
                           float spawn_prob, int n_steps, uint64_t *rng);
int sl_alive_counts_host(const uint16_t *board, const uint16_t *goals, int64_t *out,
                         int n, int cells);
int sl_execute_actions_host(uint16_t *board, int64_t *locations, const int64_t *actions,
                            int n, int H, int W, int n_agents, int action_stride);
int sl_wrapped_label_host(int32_t *data, int32_t *num_labels, int n, int H, int W);

/* Earth mover's distance between two histograms over the same k points, host
 * memory, no GPU involved: *out = min-cost flow of min(sum a, sum b) mass under
 * dist[k][k] + extra_mass_penalty * |sum a - sum b| (extra_mass_penalty < 0:
 * max(dist)).  dist must be a (quasi-)metric -- dist[i][i] = 0 and the directed
 * triangle inequality -- as for pyemd's emd_hat_gd_metric: mass common to a[i]
 * and b[i] stays in place.  Stands in for pyemd.emd as called by
 * side_effects.earth_mover_distance (side_effects.py:57); exact in double
 * precision, where pyemd/FastEMD quantises to 1e-6 of the maxima. */
int sl_emd_host(const double *a, const double *b, const double *dist, int k,
                double extra_mass_penalty, double *out);
/* n independent problems (k[i] points each) on up to n_threads host threads
 * (n_threads <= 0: one per hardware thread): the per-cell-type distances of all
 * the episodes that finished in one step. */
int sl_emd_host_batch(const double *const *a, const double *const *b,
                      const double *const *dist, const int *k, int n,
                      double extra_mass_penalty, double *out, int n_threads);

/* ---- batched environment (SafeLifeEnv.step / reset) ------------------- */

/*
 * Per-env scalar state, one 96-byte record per env (device memory).  Packed so
 * that a warp reads the whole record with six 16-byte loads.
 */
typedef struct SlEnvState {
    uint64_t rng[4];       /* PCG64 of the game: state_hi, state_lo, inc_hi, inc_lo
                              (safelife_game.py:175-180) */
    int32_t  loc_row;      /* agent 0 (row, col); row < 0: the level has no agent */
    int32_t  loc_col;
    int32_t  level_id;     /* index into the level pool */
    int32_t  num_steps;    /* GameState.num_steps */
    int32_t  old_value;    /* SafeLifeEnv._old_game_value */
    int32_t  ep_length;    /* SafeLifeEnv.episode_length */
    float    ep_reward;    /* SafeLifeEnv.episode_reward */
    uint8_t  active;       /* SafeLifeEnv._is_active */
    uint8_t  static_goals; /* 0 = unknown (None), 1 = static, 2 = dynamic
                              (GameWithGoals._static_goals, safelife_game.py:753-761) */
    uint8_t  empty_points; /* the level's points table has a non-zero "empty" column */
    uint8_t  reserved0;
    /* copied from the level pool by sl_env_reset */
    uint64_t spawn_thr;    /* ceil((double)(float)spawn_prob * 2^53): a draw (u64 >> 11)
                              spawns iff it is below this */
    int32_t  init_points;  /* sum(points_table * initial_counts) (safelife_game.py:689-694) */
    int32_t  required;     /* required_points() (:711-714) */
    int32_t  count_points; /* sum(points_table * alive_counts) of the board in memory */
    int32_t  n_exits;      /* number of exit cells of the level */
    int32_t  exit0, exit1; /* the first two of them (flat indices, -1: none); levels with more
                              keep the full list in pool_exits */
} SlEnvState;

/* Promises a caller may make about its level pool; they only select faster
 * kernels and never change results.
 * SL_POOL_UNIFORM_TABLE  every level's points table equals pool_table[0]: the
 *     step kernels keep it in shared memory.
 * SL_POOL_STATIC_GOALS   every level's goals are a fixed point of advance_board
 *     and hold no SPAWNING cell, so GameWithGoals._static_goals becomes True on
 *     the first step of every episode (safelife_game.py:753-761; true of the
 *     append / prune -still and -spawn tasks): the step kernels leave the goal
 *     planes in HBM and read only the goal cells under changed board cells.
 *     If a level breaks the promise its envs still step correctly, slowly. */
#define SL_POOL_UNIFORM_TABLE 1
#define SL_POOL_STATIC_GOALS  2

/*
 * Device-resident state of n_envs single-agent SafeLife environments plus the
 * read-only level pool they reset from.  All pointers are device pointers
 * owned by the caller (torch tensors in the Python host layer).
 */
typedef struct SlEnvBatch {
    int32_t n_envs, H, W;
    int32_t max_exits;     /* width of pool_exits */
    int32_t time_limit;    /* SafeLifeEnv.time_limit, safelife_env.py:65 */
    int32_t n_levels;
    int32_t level_stride;  /* auto-reset rule: level_id <- (level_id + level_stride) % n_levels */
    int32_t pool_flags;    /* SL_POOL_* promises about the level pool (0 is always valid) */

    /* per-env state */
    uint16_t   *board;     /* [n_envs][H][W]  SafeLifeGame.board */
    uint16_t   *goals;     /* [n_envs][H][W]  SafeLifeGame.goals */
    SlEnvState *state;     /* [n_envs] */
    /* per-step outputs */
    float    *reward;      /* [n_envs] float32 reward (safelife_env.py:157,172) */
    uint8_t  *done;        /* [n_envs] */
    uint8_t  *success;     /* [n_envs] has_exited (safelife_env.py:159) */
    uint8_t  *times_up;    /* [n_envs] info['times_up'] */

    /* level pool, read-only */
    const uint16_t *pool_board;       /* [n_levels][H][W] as deserialised, before update_exit_colors */
    const uint16_t *pool_goals;       /* [n_levels][H][W] */
    const int32_t  *pool_loc;         /* [n_levels][2], row < 0: no agent */
    const int32_t  *pool_table;       /* [n_levels][72] points_table[0] (safelife_game.py:595-605) */
    const int32_t  *pool_exits;       /* [n_levels][max_exits] flat indices of exit cells, -1 padded,
                                         raster order (GameState.exit_locs, safelife_game.py:533-535) */
    const int32_t  *pool_init_points; /* [n_levels] */
    const int32_t  *pool_required;    /* [n_levels] */
    const uint64_t *pool_spawn_thr;   /* [n_levels] */

    /* optional (NULL: none): the step's outputs once more, packed 8 bytes per env
     * { float reward; uint8 done, success, times_up, tag } -- for host-driven loops that
     * point this at pinned host memory: the envs of a CTA then cross PCIe as ONE write
     * of up to 64 bytes instead of four small ones.  tag is 0, or inside a host session
     * the low byte of the step's sequence number. */
    uint64_t *host_out;               /* [n_envs] */
} SlEnvBatch;

/* One SafeLifeEnv.step for every env (safelife_env.py:148-201, single_agent):
 * execute_actions -> advance_board (board, then goals unless static) ->
 * update_exit_colors -> reward / done / episode counters.  actions is
 * uint8[n_envs] (device), values as the reference's Discrete(9).
 * kernel: 0 = pick automatically, 1 = generic CTA-per-env kernel,
 * 2 = the fast kernels (warp per env up to 32 x 32, tiles up to 120 x 120) or an error,
 * 3 = as 2, with the persistent pipelined kernel wherever it applies (by default it
 *     runs only where it measured faster: large batches without the observation).
 * flags: SL_STEP_AUTO_RESET resets every env that finished in this step in
 * the same call, exactly as sl_env_reset(batch, done, NULL,
 * SL_RESET_NEXT_LEVEL | SL_RESET_KEEP_STATS) would (the batched form of
 * BaseAlgo.act_on_envs resetting on done, training/base_algo.py:231-236);
 * reward / done / success / times_up still describe the finished step. */
#define SL_STEP_AUTO_RESET 1
int sl_env_step(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                void *stream);

/* The observation a step returns (SafeLifeEnv.step ends in `return self.get_obs(),
 * ...`, safelife_env.py:195; get_obs :105-146 with helper_utils.recenter_view,
 * helper_utils.py:42-75), described as for sl_get_obs but with the channel list
 * by value.  obs: device uint8[n_envs][vh][vw][n_channels], or uint32[n_envs][vh][vw]
 * when n_channels == 0 (output_channels=None). */
typedef struct SlObsSpec {
    void   *obs;
    int32_t vh, vw;            /* SafeLifeEnv.view_shape */
    int32_t n_channels;        /* 0 .. 32 */
    int32_t remove_white_goals;
    int32_t channels[32];      /* SafeLifeEnv.output_channels: bit positions 0 .. 31 */
} SlObsSpec;

/* sl_env_step with the observation written by the same launch, from the board
 * slabs the step kernel already holds in shared memory: what SafeLifeEnv.step
 * returns, one launch per gym step.  The observation describes the env as the
 * call leaves it -- with SL_STEP_AUTO_RESET a finished env shows the first
 * observation of its next episode, which is what the reference's rollout loop
 * feeds the policy next (training/base_algo.py:231-238; the terminal observation
 * of a finished env is never used there, training/ppo.py:117-121).  obs == NULL
 * is sl_env_step.  On the generic path (kernel 1, or boards beyond the warp
 * kernels) the observation is a separate launch of the same writer. */
int sl_env_step_obs(const SlEnvBatch *batch, const uint8_t *actions, const SlObsSpec *obs,
                    int kernel, int flags, void *stream);

/* ---- training wrappers as an epilogue of the step (safelife/env_wrappers.py) ---- */

/* Per-env wrapper state, 48 bytes (device memory). */
typedef struct SlWrapState {
    int16_t hist[8][2];        /* MovementBonusWrapper._prior_positions, oldest first (row, col) */
    int32_t hist_len;          /* 1 .. movement_period after a reset */
    int32_t last_side_effect;  /* SimpleSideEffectPenalty.last_side_effect */
    int32_t reserved[2];
} SlWrapState;

#define SL_WRAP_MOVEMENT            1   /* MovementBonusWrapper, env_wrappers.py:32-99 */
#define SL_WRAP_MOVEMENT_PENALTY    2   /*   as_penalty=True: reward -= movement_bonus */
#define SL_WRAP_EXIT_BONUS          4   /* ExtraExitBonus, :121-128 */
#define SL_WRAP_SIDE_EFFECT         8   /* SimpleSideEffectPenalty, :148-213 */
#define SL_WRAP_IGNORE_REWARD_CELLS 16  /*   ignore_reward_cells=True */

/* The wrapper stack of training/env_factory.py:274-283, applied in that order
 * (movement bonus, exit bonus, side-effect penalty; min-performance rescaling at
 * reset).  All pointers are device pointers owned by the caller. */
typedef struct SlWrapSpec {
    int32_t flags;                  /* SL_WRAP_* */
    int32_t movement_period;        /* 1 .. 8 */
    double  movement_bonus;
    const double *movement_table;   /* [H + W + movement_period]: movement_bonus *
                                       (d / movement_period) ** movement_bonus_power by L1
                                       distance d, evaluated by the caller in float64 */
    double  exit_bonus;             /* call(ExtraExitBonus.bonus) for this step */
    double  penalty_coef;           /* call(SimpleSideEffectPenalty.penalty_coef) for this step */
    uint16_t *baseline;             /* [n_envs][H][W] SimpleSideEffectPenalty.baseline_board */
    SlWrapState *wstate;            /* [n_envs] */
    double  *reward;                /* [n_envs] out: the wrapped reward (float64, as the
                                       reference's becomes) */
    int32_t *side_effect;           /* [n_envs] out, optional: the current cell count */
    const int32_t *pool_required;   /* [n_levels] required points under
                                       MinPerformanceScheduler.min_performance_fraction, written
                                       into the env record after every reset; NULL: none */
} SlWrapSpec;

/* sl_env_step_obs with the wrapper epilogue in the same launch: after the scalar
 * tail of an env (and before its auto-reset) the side-effect cells are counted
 * against wrap->baseline, the wrapped reward is written to wrap->reward, and a
 * reset re-initialises the wrapper state of the env (position ring, baseline
 * board = the fresh board, last_side_effect = 0, rescaled required points).
 * batch->reward still receives the base env's float32 reward.  obs may be NULL.
 * With baseline="inaction" the caller advances wrap->baseline (sl_advance_board)
 * before this call, as env_wrappers.py:179-180 does before comparing. */
int sl_env_step_wrapped(const SlEnvBatch *batch, const uint8_t *actions, const SlObsSpec *obs,
                        const SlWrapSpec *wrap, int kernel, int flags, void *stream);

/* The wrappers' own reset() for the envs whose mask byte is set (NULL: all), to
 * follow an sl_env_reset of the same envs. */
int sl_wrap_reset(const SlEnvBatch *batch, const SlWrapSpec *wrap, const uint8_t *mask,
                  void *stream);

/* Which kernel(s) sl_env_step(batch, ..., kernel, ...) launches on the current
 * device, as text (for benchmarks and profiles), and how many launches that is
 * per step without SL_STEP_AUTO_RESET on the generic path.  Launches nothing. */
int sl_env_step_plan(const SlEnvBatch *batch, int kernel, char *text, int text_len,
                     int *n_launches);
/* The same for sl_env_step_obs (obs may be NULL). */
int sl_env_step_obs_plan(const SlEnvBatch *batch, const SlObsSpec *obs, int kernel, char *text,
                         int text_len, int *n_launches);

/* An empty kernel with the step kernel's launch shape (the single-pass kernel's grid and
 * CTA size for this batch, the same parameter block): what a benchmark's event pair
 * measures around a step that does no work. */
int sl_launch_floor(const SlEnvBatch *batch, void *stream);

/* sl_env_step followed by a wait for the stream: for host-driven loops.  The
 * batch's actions / reward / done / success / times_up pointers may be pinned
 * host memory (device-addressable under UVA); the kernel then reads and writes
 * them in place and the results are valid on return. */
int sl_env_step_sync(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                     void *stream);
/* The same; if timing_us is not NULL it receives the host-side cost of the call
 * in microseconds: [0] validation + launch, [1] wait until the results are
 * visible in host memory (cudaStreamSynchronize; with SAFELIFE_B200_HOST_WAIT=1
 * a spin on a word of pinned memory that a stream memory operation writes after
 * the kernel, which measured 1.2 us slower). */
int sl_env_step_host(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                     void *stream, double *timing_us);
/* The two halves of sl_env_step_host for callers that keep several batches in
 * flight (one stream each) and overlap the host's work on one batch with the
 * device's on another: submit launches the step and queues, in stream order, a
 * write of `seq` into the pinned host word `flag`; wait spins until the word
 * holds `seq`, at which point the step's outputs in pinned memory are valid. */
int sl_env_step_submit(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                       void *stream, uint32_t *flag, uint32_t seq);
int sl_host_wait(const uint32_t *flag, uint32_t seq, void *stream);

/* ---- host sessions: many host-driven steps, one resident kernel --------------
 * For a policy that lives on the host (actions in, reward / done out, once per
 * host round trip: the CPU form of BaseAlgo.act_on_envs, training/base_algo.py:
 * 231-236) a launch per step is mostly launch and completion latency.  A session
 * keeps one kernel resident: the envs stay in shared memory, the host posts a
 * step by writing tagged actions into pinned memory, the kernel answers by
 * writing the packed result records (SlEnvBatch.host_out), each tagged with the
 * step's sequence number in its top byte, into pinned memory.  Boards up to 32 x 32, batches whose CTAs are all resident
 * at once (about 4700 envs per GPU); physics + auto-reset, no observation.
 * The caller allocates the mailbox:
 *   host_actions  pinned, n_pad = (n_envs + 15) & ~15 bytes, zeroed
 *   dev_actions   device, 2 * n_pad bytes
 *   dev_ctrl      device, SL_SESSION_CTRL_WORDS x uint32 (words 8..23: globaltimer
 *                 timeline of the last step; then the per-SM mailbox lines)
 *   host_ctrl     pinned, 4 x uint32: [0] last completed step, [1] alive,
 *                 [2] written by the host: 1 = end the session
 * While a session is open the batch's state in HBM is stale and no other call
 * may touch the batch; other CUDA work on the device waits for the session to
 * end (it ends by itself after idle_timeout_ms without a step). */
#define SL_SESSION_CTRL_WORDS (32 + 32 * 256)
typedef struct SlHostSession {
    uint8_t *host_actions;
    uint8_t *dev_actions;
    uint32_t *dev_ctrl;
    uint32_t *host_ctrl;
    uint64_t *host_out;     /* library-owned: the batch's result records (SlEnvBatch.host_out) */
    uint32_t idle_timeout_ms;
    uint32_t seq;           /* steps posted in this session (library-owned) */
    int32_t n_envs;         /* library-owned */
    int32_t reserved;
} SlHostSession;
#define SL_SESSION_ENDED 4  /* sl_host_session_step: the session has ended (idle timeout); begin again */
/* Allocates the mailbox for n_envs envs on the current device (host_actions
 * write-combined if asked: the device then reads it without snooping the CPU's
 * caches, which is several times faster for lines the CPU has just written);
 * sl_host_session_free releases it.  A caller may also bring its own buffers. */
int sl_host_session_alloc(SlHostSession *session, int n_envs, int write_combined);
int sl_host_session_free(SlHostSession *session);
/* Diagnostic: the eight globaltimer stamps (ns) the session kernel left for its last
 * step (csrc/session.cuh); call between sessions. */
int sl_host_session_timeline(const SlHostSession *session, uint64_t *stamps8);
/* Starts the session kernel on `stream` (the batch must have host_out set and all
 * its arrays 16-byte aligned).  flags as for sl_env_step. */
int sl_host_session_begin(const SlEnvBatch *batch, SlHostSession *session, int flags, void *stream);
/* One step: posts actions[n_envs] (host memory, values 0..8) and returns when the
 * result records of the step are in host_out.  No CUDA call is made.  timing_us
 * (may be NULL): [0] posting, [1] waiting.  SL_SESSION_ENDED: nothing was stepped. */
int sl_host_session_step(SlHostSession *session, const uint8_t *actions, double *timing_us);
/* Ends the session and waits for the kernel: the batch's state is back in HBM. */
int sl_host_session_end(SlHostSession *session, void *stream);

/* SafeLifeEnv.reset (safelife_env.py:203-218) for every env whose mask byte is
 * non-zero (mask NULL = all envs): copy a level from the pool, then
 * update_exit_colors and the episode counters.  The level is new_level[e]
 * (device int32) if given, else level_id[e], advanced by level_stride when
 * SL_RESET_NEXT_LEVEL is set.  With SL_RESET_KEEP_STATS ep_length / ep_reward
 * keep the finished episode's totals until the env's next step (what
 * info['episode'] reports after an auto-reset).  The RNG state is left as it
 * is; callers that want the reference's per-level SeedSequence stream write
 * rng[e] themselves. */
#define SL_RESET_NEXT_LEVEL 1
#define SL_RESET_KEEP_STATS 2
int sl_env_reset(const SlEnvBatch *batch, const uint8_t *mask, const int32_t *new_level,
                 int flags, void *stream);

/* SafeLifeEnv.get_obs (safelife_env.py:105-146) with recenter_view
 * (helper_utils.py:42-75): agent-centred toroidal view of board | goal
 * colours << 16, exit cells outside the view clamped to its perimeter.
 * 0 < n_channels <= 32: obs is uint8[n_envs][vh][vw][n_channels] with
 * obs[...,k] = (view >> channels[k]) & 1 (channels: device int32, each 0..31);
 * n_channels == 0: obs is uint32[n_envs][vh][vw] (output_channels=None). */
int sl_get_obs(const SlEnvBatch *batch, void *obs, int vh, int vw,
               const int32_t *channels, int n_channels, int remove_white_goals,
               void *stream);
/* The same with the view described by an SlObsSpec (channel list in host memory,
 * no copy, no synchronisation): what SafeLifeEnv.reset returns. */
int sl_get_obs_spec(const SlEnvBatch *batch, const SlObsSpec *obs, void *stream);

/* SimpleSideEffectPenalty.step cell count (env_wrappers.py:177-213,
 * ignore_reward_cells=False): out[e] = number of cells whose non-player bits
 * differ from baseline[e], exit cells excluded.  baseline: [n_envs][H][W]. */
int sl_side_effect_cells(const SlEnvBatch *batch, const uint16_t *baseline, int32_t *out,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SAFELIFE_B200_H */
