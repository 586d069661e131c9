"""
Batched forms of the reference's training wrappers (safelife/env_wrappers.py),
stacked in the order of training/env_factory.py:274-283:

    env = BatchedSafeLifeEnv(levels, N, auto_reset=False, ...)
    env = MovementBonusWrapper(env, as_penalty=True)
    env = ExtraExitBonus(env)
    env = SimpleSideEffectPenalty(env, baseline="starting-state", penalty_coef=c)
    env = MinPerformanceScheduler(env, min_performance_fraction=f)
    obs = env.reset()
    obs, reward, done, info = env.step(actions)      # reward: float64 [N]
    obs = env.reset_done()                            # what act_on_envs does after a done

Same attribute names and arithmetic as the reference, on [N] device tensors.
The reference's rewards become float64 once a wrapper adds a Python float to the
float32 base reward; the wrappers here compute in float64 in the same order.
Resets are driven from outside (reset / reset_done), exactly as the reference's
rollout loop drives them: the wrapped env must be built with auto_reset=False so
that the final step of an episode is still visible to the wrappers.
"""
import numpy as np
import torch

from .env import BatchedSafeLifeEnv

PLAYER_BITS = 2 | 8 | 16 | 32 | 64      # CellTypes.player without colour / orientation


def call(x):
    return x() if callable(x) else x


class BaseWrapper(object):
    """env_wrappers.py:16-29: attribute forwarding + kwargs as attributes."""

    def __init__(self, env, **kwargs):
        self.env = env
        base = env
        while isinstance(base, BaseWrapper):
            base = base.env
        if not isinstance(base, BatchedSafeLifeEnv):
            raise ValueError("wrappers need a BatchedSafeLifeEnv underneath")
        if base.auto_reset:
            raise ValueError("build the wrapped BatchedSafeLifeEnv with auto_reset=False")
        self.unwrapped = base
        for key, val in kwargs.items():
            if key.startswith("_") or not hasattr(type(self), key):
                raise ValueError("Unrecognized parameter: '%s'" % (key,))
            setattr(self, key, val)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, *args, **kwargs):
        obs = self.env.reset(*args, **kwargs)
        self._after_reset(None)
        return obs

    def reset_done(self, mask=None):
        base = self.unwrapped
        mask = (base._t["done"] if mask is None else mask).to(torch.uint8).clone()
        obs = self.env.reset_done(mask)
        self._after_reset(mask.bool())
        return obs

    def _after_reset(self, mask):
        """mask: bool [N] of the envs that were reset, or None for all."""

    def step(self, action):
        return self.env.step(action)


class MovementBonusWrapper(BaseWrapper):
    """env_wrappers.py:32-99."""
    movement_bonus = 0.1
    movement_bonus_power = 1e-100
    movement_bonus_period = 4
    as_penalty = True

    def _after_reset(self, mask):
        base = self.unwrapped
        loc = base._t["loc"].clone()
        n = self.movement_bonus_period
        if mask is None:
            self._hist = loc.unsqueeze(0).repeat(n, 1, 1)     # [-1] is the newest
            self._len = torch.ones(base.num_envs, dtype=torch.int64, device=base.device)
        else:
            self._hist[:, mask] = loc[mask].unsqueeze(0)
            self._len[mask] = 1

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        base = self.unwrapped
        n = self.movement_bonus_period
        p0 = base._t["loc"].clone()
        has_agent = p0[:, 0] >= 0
        k = torch.clamp(self._len, max=n)
        ar = torch.arange(base.num_envs, device=base.device)
        p1 = self._hist[n - k, ar]
        dist = (p0 - p1).abs().sum(dim=1) * has_agent + (n - k)
        speed = dist.to(torch.float64) / n
        reward = reward.to(torch.float64) + self.movement_bonus * speed ** self.movement_bonus_power
        if self.as_penalty:
            reward = reward - self.movement_bonus
        self._hist = torch.cat([self._hist[1:], p0.unsqueeze(0)])
        self._len = torch.clamp(self._len + 1, max=n)
        return obs, reward, done, info


class ContinuingEnv(BaseWrapper):
    """env_wrappers.py:102-118: only a time-out ends the episode."""

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        restart = done & ~info["times_up"]
        if bool(restart.any()):
            obs = self.reset_done(restart)
        return obs, reward, done & info["times_up"], info


class ExtraExitBonus(BaseWrapper):
    """env_wrappers.py:121-128."""
    bonus = 0.5

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        extra = done.to(torch.float64) * call(self.bonus) * info["episode"]["reward"].to(torch.float64)
        reward = reward.to(torch.float64) + torch.where(info["times_up"], torch.zeros_like(extra), extra)
        return obs, reward, done, info


class MinPerformanceScheduler(BaseWrapper):
    """
    env_wrappers.py:131-145: after each reset min_performance is scaled, which
    changes required_points from the first step on (the exit colours written by
    the reset itself still use the unscaled value, as in the reference).
    """
    min_performance_fraction = 1

    def _after_reset(self, mask):
        base = self.unwrapped
        frac = float(call(self.min_performance_fraction))
        pool = base.pool
        # required_points() with the scaled min_performance (safelife_game.py:711-714)
        avail = np.array([_available_points(pool, i) for i in range(len(pool))], dtype=np.float64)
        mp = np.array([lv["min_performance"] for lv in pool.levels], dtype=np.float64) * frac
        req = np.maximum(0, np.ceil(mp * avail).astype(np.int64)).astype(np.int32)
        req_t = torch.from_numpy(req).to(base.device)
        per_env = req_t[base._t["level_id"].long()]
        if mask is None:
            base._t["required"].copy_(per_env)
        else:
            base._t["required"][mask] = per_env[mask]


def _available_points(pool, i):
    """GameWithGoals.initial_available_points (safelife_game.py:696-709) of pool level i."""
    from .levels import AGENT, ALIVE, COLORS, SPAWNING, alive_counts_np
    cache = pool.__dict__.setdefault("_available", {})
    if i not in cache:
        table = pool.table[i].reshape(8, 9).astype(np.int64)
        counts = alive_counts_np(pool.board[i], pool.goals[i])
        colors = np.zeros(9, dtype=bool)
        gen = pool.board[i][(pool.board[i] & (AGENT | ALIVE | SPAWNING)) > 0] & COLORS
        colors[np.unique(gen) >> 9] = True
        colors[8] = True
        cache[i] = int(((table * colors).max(axis=1) * counts.sum(axis=1)).sum()) - int(pool.init_points[i])
    return cache[i]


class SimpleSideEffectPenalty(BaseWrapper):
    """env_wrappers.py:148-213."""
    penalty_coef = 0.0
    baseline = "starting-state"     # or "inaction"
    ignore_reward_cells = False

    def _advance_baseline(self, base):
        """
        The counterfactual board advances with its level's spawn_prob
        (``advance_board(self.baseline_board, self.game.spawn_prob)``,
        env_wrappers.py:179-180).  Its draws come from per-env PCG64 streams of
        their own: the reference takes them from the process-global generator,
        which has no batched counterpart.
        """
        from . import _lib
        pool = getattr(self, "_pool_spawn", None)
        if pool is None:
            probs = np.array([lv["spawn_prob"] for lv in base.pool.levels], dtype=np.float64)
            self._uniform_spawn = float(probs[0]) if (probs == probs[0]).all() else None
            pool = self._pool_spawn = torch.from_numpy(probs).to(base.device)
        bb, rng = self.baseline_board, self._baseline_rng
        H, W = base.pool.H, base.pool.W
        if self._uniform_spawn is not None:
            _lib.check(base._lib.sl_advance_board(
                bb.data_ptr(), bb.data_ptr(), base.num_envs, H, W, self._uniform_spawn, 1,
                rng.data_ptr(), base._stream()))
            return
        per_env = pool[base._t["level_id"].long()]
        for p in torch.unique(per_env).tolist():          # mixed pools: one launch per value
            idx = torch.nonzero(per_env == p).flatten()
            sub, sub_rng = bb[idx].contiguous(), rng[idx].contiguous()
            _lib.check(base._lib.sl_advance_board(
                sub.data_ptr(), sub.data_ptr(), idx.numel(), H, W, float(p), 1,
                sub_rng.data_ptr(), base._stream()))
            bb[idx] = sub
            rng[idx] = sub_rng

    def _after_reset(self, mask):
        base = self.unwrapped
        board = base._t["board"]
        if mask is None:
            self.baseline_board = board.clone()
            self.last_side_effect = torch.zeros(base.num_envs, dtype=torch.int64, device=base.device)
            self._baseline_rng = base._t["rng"].clone().contiguous()
        else:
            self.baseline_board[mask] = board[mask]
            self.last_side_effect[mask] = 0

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        base = self.unwrapped
        if self.baseline == "inaction":
            self._advance_baseline(base)
        if not self.ignore_reward_cells:
            side_effect = base.side_effect_cells(self.baseline_board).to(torch.int64)
        else:
            side_effect = self._side_effect_ignoring_rewards(base)
        delta = side_effect - self.last_side_effect
        reward = reward.to(torch.float64) - delta.to(torch.float64) * call(self.penalty_coef)
        self.last_side_effect = side_effect
        return obs, reward, done, info

    def _side_effect_ignoring_rewards(self, base):
        keep = ~PLAYER_BITS & 0xFFFF
        board = base._t["board"].to(torch.int32) & keep
        baseline = self.baseline_board.to(torch.int32) & keep
        exits = torch.from_numpy(base.pool.exits).to(base.device)[base._t["level_id"].long()]
        N, H, W = board.shape
        flat, bflat = board.view(N, -1), baseline.view(N, -1)
        valid = exits >= 0
        idx = exits.clamp(min=0).long()
        flat.scatter_(1, idx, torch.where(valid, bflat.gather(1, idx), flat.gather(1, idx)))
        unchanged = flat == bflat
        red_life = 1 | (1 << 9)
        start_red = (bflat & red_life) == red_life
        end_red = (flat & red_life) == red_life
        goals = base._t["goals"].to(torch.int32).view(N, -1)
        goal_cell = (goals & 0x0E00) == (4 << 9)
        end_alive = (flat & red_life) == 1
        non_effects = unchanged | (start_red & ~end_red) | (goal_cell & end_alive)
        return (~non_effects).sum(dim=1)


def training_stack(env, exit_difficulty=1.0, se_baseline="starting-state", se_penalty=0.0,
                   ignore_reward_cells=False):
    """The wrapper chain of training/env_factory.py:274-283 (without the logger)."""
    env = MovementBonusWrapper(env, as_penalty=True)
    env = ExtraExitBonus(env)
    env = SimpleSideEffectPenalty(env, baseline=se_baseline, penalty_coef=se_penalty,
                                  ignore_reward_cells=ignore_reward_cells)
    env = MinPerformanceScheduler(env, min_performance_fraction=exit_difficulty)
    return env
