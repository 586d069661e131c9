"""
Batched forms of the reference's training wrappers (safelife/env_wrappers.py),
stacked in the order of training/env_factory.py:274-283:

    env = BatchedSafeLifeEnv(levels, N, ...)
    env = MovementBonusWrapper(env, as_penalty=True)
    env = ExtraExitBonus(env)
    env = SimpleSideEffectPenalty(env, baseline="starting-state", penalty_coef=c)
    env = MinPerformanceScheduler(env, min_performance_fraction=f)
    obs = env.reset()
    obs, reward, done, info = env.step(actions)      # reward: float64 [N]

Same class names, attribute names and arithmetic as the reference, but the
wrappers do no work of their own: each one registers its parameters with the
BatchedSafeLifeEnv underneath, and the env's ONE launch per step
(sl_env_step_wrapped, csrc/wrap.cuh) runs the whole stack as an epilogue of the
step kernel -- the position ring, the side-effect cell count against the
baseline board and the reward arithmetic, before the in-kernel reset of
finished envs re-initialises the wrapper state.  The kernel applies the
wrappers in the reference's stacking order (movement bonus, exit bonus,
side-effect penalty), whatever order they were attached in.

The reference's rewards become float64 once a wrapper adds a Python float to the
float32 base reward; the epilogue computes in float64 with one IEEE operation
per reference operation.  Parameters may be callables (schedules), evaluated
once per step like the reference's ``call(...)``.

With ``auto_reset=True`` finished envs restart inside the same launch, as
``BaseAlgo.act_on_envs`` restarts them right after the step; with
``auto_reset=False`` the caller drives ``reset_done()``.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .env import BatchedSafeLifeEnv, seed_rng_states, _as_i64

PLAYER_BITS = 2 | 8 | 16 | 32 | 64      # CellTypes.player without colour / orientation


def call(x):
    return x() if callable(x) else x


class WrapRuntime(object):
    """
    Device state and launch parameters of the wrapper stack of one env
    (struct SlWrapSpec / SlWrapState, include/safelife_b200.h).
    """

    def __init__(self, base):
        self.base = base
        N, dev = base.num_envs, base.device
        self.movement = self.exit_bonus = self.side_effect = self.min_performance = None
        self.wstate = torch.zeros((N, _lib.WRAP_STATE_WORDS), dtype=torch.int32, device=dev)
        self.rewards = [torch.zeros(N, dtype=torch.float64, device=dev) for _ in range(2)]
        self.reward = self.rewards[0]
        self.count = torch.zeros(N, dtype=torch.int32, device=dev)
        self.baseline = None
        self._table = self._table_key = None
        self._required = self._required_frac = None
        self.spec = _lib.SlWrapSpec()
        self._baseline_rng = None

    # ---- parameters -> struct SlWrapSpec, once per launch
    def _movement_table(self, m):
        base = self.base
        key = (float(m.movement_bonus), float(m.movement_bonus_power), int(m.movement_bonus_period))
        if key != self._table_key:
            n = key[2]
            if not 1 <= n <= 8:
                raise ValueError("movement_bonus_period must be 1..8")
            dist = np.arange(base.pool.H + base.pool.W + n, dtype=np.int64)
            speed = dist / n                                   # env_wrappers.py:87
            table = key[0] * speed ** key[1]                   # :90
            self._table = torch.from_numpy(np.asarray(table, dtype=np.float64)).to(base.device)
            self._table_key = key
        return self._table

    def _scaled_required(self, frac):
        """required_points() of every level under min_performance * frac (safelife_game.py:711-714)."""
        if frac != self._required_frac:
            pool = self.base.pool
            avail = np.array([_available_points(pool, i) for i in range(len(pool))], dtype=np.float64)
            mp = np.array([lv["min_performance"] for lv in pool.levels], dtype=np.float64) * frac
            req = np.maximum(0, np.ceil(mp * avail).astype(np.int64)).astype(np.int32)
            if self._required is None:
                self._required = torch.from_numpy(req).to(self.base.device)
            else:
                self._required.copy_(torch.from_numpy(req), non_blocking=False)
            self._required_frac = frac
        return self._required

    def fill_spec(self, flip=0):
        s, base = self.spec, self.base
        flags = 0
        m = self.movement
        if m is not None:
            flags |= _lib.SL_WRAP_MOVEMENT | (_lib.SL_WRAP_MOVEMENT_PENALTY if m.as_penalty else 0)
            s.movement_table = self._movement_table(m).data_ptr()
            s.movement_period = int(m.movement_bonus_period)
            s.movement_bonus = float(m.movement_bonus)
        if self.exit_bonus is not None:
            flags |= _lib.SL_WRAP_EXIT_BONUS
            s.exit_bonus = float(call(self.exit_bonus.bonus))
        se = self.side_effect
        if se is not None:
            flags |= _lib.SL_WRAP_SIDE_EFFECT
            if se.ignore_reward_cells:
                flags |= _lib.SL_WRAP_IGNORE_REWARD_CELLS
            s.penalty_coef = float(call(se.penalty_coef))
            if self.baseline is None:
                self.baseline = torch.zeros_like(base._t["board"])
            s.baseline = self.baseline.data_ptr()
        mp = self.min_performance
        s.pool_required = (self._scaled_required(float(call(mp.min_performance_fraction))).data_ptr()
                           if mp is not None else None)
        s.flags = flags
        s.wstate = self.wstate.data_ptr()
        self.reward = self.rewards[flip]
        s.reward = self.reward.data_ptr()
        s.side_effect = self.count.data_ptr()
        return ctypes.byref(s)

    # ---- SimpleSideEffectPenalty(baseline="inaction"): the counterfactual board
    def advance_baseline(self):
        """
        ``self.baseline_board = advance_board(self.baseline_board, self.game.spawn_prob)``
        (env_wrappers.py:179-180) for every env, before the step's comparison.  Its
        draws come from per-env PCG64 streams of their own, seeded independently of
        the games': the reference takes them from the process-global generator,
        which has no batched counterpart.
        """
        base = self.base
        if self._baseline_rng is None:
            seed = np.random.SeedSequence([0xBA5E11E, base.num_envs, base.env_offset])
            self._baseline_rng = _as_i64(seed_rng_states(seed, base.num_envs)).to(base.device)
            probs = np.array([lv["spawn_prob"] for lv in base.pool.levels], dtype=np.float64)
            self._uniform_spawn = float(probs[0]) if (probs == probs[0]).all() else None
            self._pool_spawn = torch.from_numpy(probs).to(base.device)
        bb, rng = self.baseline, self._baseline_rng
        H, W = base.pool.H, base.pool.W
        with torch.cuda.device(base.device):
            if self._uniform_spawn is not None:
                _lib.check(base._lib.sl_advance_board(
                    bb.data_ptr(), bb.data_ptr(), base.num_envs, H, W, self._uniform_spawn, 1,
                    rng.data_ptr(), base._stream()))
                return
            per_env = self._pool_spawn[base._t["level_id"].long()]
            for p in torch.unique(per_env).tolist():          # mixed pools: one launch per value
                idx = torch.nonzero(per_env == p).flatten()
                sub, sub_rng = bb[idx].contiguous(), rng[idx].contiguous()
                _lib.check(base._lib.sl_advance_board(
                    sub.data_ptr(), sub.data_ptr(), idx.numel(), H, W, float(p), 1,
                    sub_rng.data_ptr(), base._stream()))
                bb[idx] = sub
                rng[idx] = sub_rng

    def before_step(self):
        se = self.side_effect
        if se is not None and se.baseline == "inaction":
            if self.baseline is None:
                raise RuntimeError("reset() the env before stepping it")
            self.advance_baseline()
        elif se is not None and se.baseline != "starting-state":
            raise ValueError("baseline must be 'starting-state' or 'inaction'")


class BaseWrapper(object):
    """env_wrappers.py:16-29: attribute forwarding + kwargs as attributes."""
    _slot = None

    def __init__(self, env, **kwargs):
        self.env = env
        base = env
        while isinstance(base, BaseWrapper):
            base = base.env
        if not isinstance(base, BatchedSafeLifeEnv):
            raise ValueError("wrappers need a BatchedSafeLifeEnv underneath")
        self.unwrapped = base
        for key, val in kwargs.items():
            if key.startswith("_") or not hasattr(type(self), key):
                raise ValueError("Unrecognized parameter: '%s'" % (key,))
            setattr(self, key, val)
        if self._slot is not None:
            if base._wrap is None:
                base._wrap = WrapRuntime(base)
            if getattr(base._wrap, self._slot) is not None:
                raise ValueError("%s is already part of this env's wrapper stack" % type(self).__name__)
            setattr(base._wrap, self._slot, self)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, *args, **kwargs):
        return self.env.reset(*args, **kwargs)

    def reset_done(self, mask=None, out=None):
        return self.env.reset_done(mask, out)

    def step(self, action, obs_out=None):
        """The env's launch ran the whole stack: every wrapper returns the wrapped reward."""
        obs, reward, done, info = self.env.step(action, obs_out=obs_out)
        if self._slot is not None:
            reward = self.unwrapped._wrap.reward
        return obs, reward, done, info


class MovementBonusWrapper(BaseWrapper):
    """env_wrappers.py:32-99."""
    _slot = "movement"
    movement_bonus = 0.1
    movement_bonus_power = 1e-100
    movement_bonus_period = 4
    as_penalty = True


class ContinuingEnv(BaseWrapper):
    """env_wrappers.py:102-118: only a time-out ends the episode."""

    def __init__(self, env, **kwargs):
        super().__init__(env, **kwargs)
        if self.unwrapped.auto_reset:
            raise ValueError("ContinuingEnv restarts finished episodes itself: use auto_reset=False")

    def step(self, action, obs_out=None):
        obs, reward, done, info = self.env.step(action, obs_out=obs_out)
        restart = done & ~info["times_up"]
        if bool(restart.any()):
            obs = self.reset_done(restart, obs_out)
        return obs, reward, done & info["times_up"], info


class ExtraExitBonus(BaseWrapper):
    """env_wrappers.py:121-128."""
    _slot = "exit_bonus"
    bonus = 0.5


class MinPerformanceScheduler(BaseWrapper):
    """
    env_wrappers.py:131-145: after each reset min_performance is scaled, which
    changes required_points from the first step on (the exit colours written by
    the reset itself still use the unscaled value, as in the reference).
    """
    _slot = "min_performance"
    min_performance_fraction = 1


def _available_points(pool, i):
    """GameWithGoals.initial_available_points (safelife_game.py:696-709) of pool level i."""
    from .levels import AGENT, ALIVE, COLORS, SPAWNING, alive_counts_np
    cache = pool.__dict__.setdefault("_available", {})
    key = (i, id(pool.levels[i]))
    if key not in cache:
        table = pool.table[i].reshape(8, 9).astype(np.int64)
        counts = alive_counts_np(pool.board[i], pool.goals[i])
        colors = np.zeros(9, dtype=bool)
        gen = pool.board[i][(pool.board[i] & (AGENT | ALIVE | SPAWNING)) > 0] & COLORS
        colors[np.unique(gen) >> 9] = True
        colors[8] = True
        cache[key] = int(((table * colors).max(axis=1) * counts.sum(axis=1)).sum()) - int(pool.init_points[i])
    return cache[key]


class SimpleSideEffectPenalty(BaseWrapper):
    """env_wrappers.py:148-213."""
    _slot = "side_effect"
    penalty_coef = 0.0
    baseline = "starting-state"     # or "inaction"
    ignore_reward_cells = False

    @property
    def baseline_board(self):
        """int16 [N,H,W] view of the baseline boards (the reference's ``baseline_board``)."""
        return self.unwrapped._wrap.baseline

    @property
    def last_side_effect(self):
        return self.unwrapped._wrap.wstate[:, 9]


def training_stack(env, exit_difficulty=1.0, se_baseline="starting-state", se_penalty=0.0,
                   ignore_reward_cells=False):
    """The wrapper chain of training/env_factory.py:274-283 (without the logger)."""
    env = MovementBonusWrapper(env, as_penalty=True)
    env = ExtraExitBonus(env)
    env = SimpleSideEffectPenalty(env, baseline=se_baseline, penalty_coef=se_penalty,
                                  ignore_reward_cells=ignore_reward_cells)
    env = MinPerformanceScheduler(env, min_performance_fraction=exit_difficulty)
    return env
