"""
Golden fixture for the rollout adapter: the UNMODIFIED reference's
``PPO.gen_training_batch`` (training/ppo.py:75-143, through
``BaseAlgo.obs_for_envs`` / ``act_on_envs``, training/base_algo.py:152-244) run
in this container over reference ``SafeLifeEnv`` instances under the training
wrapper stack of training/env_factory.py:274-283, with a fixed deterministic
policy.  Writes tests/golden/ppo_batch.npz.

    python tests/golden/make_golden_ppo.py

The policy is an integer function of the observation (so that numpy / torch /
CUDA all agree on it bit for bit): counts = per-channel sums of the observation,
action = argmax(counts @ WP), value = (counts @ WV) / 64; the reference samples
``choice(9, p=one_hot)``, which is that action.  Levels are static (no spawner),
so the games' RNG streams do not matter; env e plays levels (e + k * stride) % L,
the batched env's level sequence.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (loads the reference with the run-time shims)
from safelife import env_wrappers  # noqa: E402
from safelife.safelife_env import SafeLifeEnv  # noqa: E402
from safelife.safelife_game import SafeLifeGame  # noqa: E402

sys.path.insert(0, mg.refload.REF_ROOT)
from training.ppo import PPO  # noqa: E402

CHANNELS = mg.TRAIN_CHANNELS
VIEW = (9, 9)
N_ENVS, STEPS, TIME_LIMIT = 10, 24, 9
GAMMA, LMDA = 0.97, 0.95
SE_PENALTY, EXIT_DIFFICULTY = 0.25, 0.5


def policy_weights():
    rs = np.random.default_rng(1234)
    return rs.integers(-3, 4, (len(CHANNELS), 9)), rs.integers(-8, 9, len(CHANNELS))


class StubModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        wp, wv = policy_weights()
        self.wp, self.wv = torch.from_numpy(wp), torch.from_numpy(wv)

    def forward(self, obs):
        counts = obs.to(torch.int64).sum(dim=(1, 2))
        action = torch.argmax(counts @ self.wp, dim=1)
        policies = torch.nn.functional.one_hot(action, 9).to(torch.float32)
        values = (counts @ self.wv).to(torch.float32) / 64
        return values, policies


def coprime_stride(n, L):
    import math
    if L <= 1:
        return 0
    s = n % L or 1
    while math.gcd(s, L) != 1:
        s += 1
    return s % L


def main():
    datas = []
    for path in ("mazes",):
        for fname, data in mg.static_games(path, 12):
            game = SafeLifeGame.loaddata(data, auto_cls=False)
            if (game.board & 128).any() or (np.asarray(game.goals) & 128).any():
                continue                                    # keep the fixture deterministic
            datas.append(game.serialize())
    L = len(datas)
    stride = coprime_stride(N_ENVS, L)
    records = [mg.level_record(SafeLifeGame.loaddata(d, auto_cls=False)) for d in datas]

    def level_stream(e):
        k = 0
        while True:
            yield SafeLifeGame.loaddata(datas[(e + k * stride) % L], auto_cls=False)
            k += 1

    envs = []
    for e in range(N_ENVS):
        env = SafeLifeEnv(level_stream(e), time_limit=TIME_LIMIT, view_shape=VIEW,
                          output_channels=CHANNELS, should_calculate_side_effects=False)
        env = env_wrappers.MovementBonusWrapper(env, as_penalty=True)
        env = env_wrappers.ExtraExitBonus(env)
        env = env_wrappers.SimpleSideEffectPenalty(env, penalty_coef=SE_PENALTY)
        env = env_wrappers.MinPerformanceScheduler(env, min_performance_fraction=EXIT_DIFFICULTY)
        envs.append(env)
    algo = PPO.__new__(PPO)
    algo.model = StubModel()
    algo.training_envs = envs
    algo.compute_device = torch.device("cpu")
    algo.gamma, algo.lmda, algo.num_steps = GAMMA, LMDA, 0
    from safelife.random import set_rng
    with set_rng(np.random.default_rng(0)):
        batch = algo.gen_training_batch(STEPS)
    out = {k: np.stack([r[k] for r in records]) for k in records[0]}
    for name in ("obs", "actions", "action_prob", "returns", "advantages", "values"):
        out["batch_" + name] = getattr(batch, name).numpy()
    out["batch_obs"] = out["batch_obs"].astype(np.uint8)
    wp, wv = policy_weights()
    out.update(wp=wp, wv=wv, n_envs=np.int32(N_ENVS), steps=np.int32(STEPS),
               time_limit=np.int32(TIME_LIMIT), gamma=np.float64(GAMMA), lmda=np.float64(LMDA),
               se_penalty=np.float64(SE_PENALTY), exit_difficulty=np.float64(EXIT_DIFFICULTY),
               view_shape=np.array(VIEW, dtype=np.int32), channels=np.array(CHANNELS, dtype=np.int32),
               num_steps=np.int64(algo.num_steps))
    path = os.path.join(HERE, "ppo_batch.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "levels", L, "stride", stride, "steps", out["batch_actions"].shape,
          "episodes ended:", int((out["batch_returns"] != 0).sum()), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
