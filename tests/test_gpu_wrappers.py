"""
Batched training wrappers (safelife_b200/wrappers.py) against rewards recorded
from the reference's wrapper stack (env_wrappers.py via
tests/golden/make_golden.py wrapped).  Rewards are float64 sums of exactly
representable terms in both implementations: compared with rtol 1e-12.
"""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _as_i16(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint16).view(np.int16))


@pytest.mark.parametrize("name", ["v1.0", "mazes", "inaction", "ignore"])
def test_training_stack_matches_reference(name):
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200 import wrappers
    d = gu.load_npz("wrapped_%s.npz" % name)
    levels = gu.levels_of(d)
    L = len(levels)
    base = BatchedSafeLifeEnv(levels, L, auto_reset=False, time_limit=int(d["time_limit"]),
                              view_shape=(5, 5), output_channels=None)
    base.set_rng_state(d["rng_init"])
    env = wrappers.training_stack(base, exit_difficulty=float(d["exit_difficulty"]),
                                  se_penalty=float(d["se_penalty"]), se_baseline=str(d["baseline"]),
                                  ignore_reward_cells=bool(d["ignore_reward_cells"]))
    env.reset()
    exits = 0
    for t in range(d["action"].shape[1]):
        obs, reward, done, info = env.step(d["action"][:, t])
        reward, done = reward.cpu().numpy(), done.cpu().numpy()
        assert reward.dtype == np.float64
        for i in np.flatnonzero(t < d["length"]):
            assert done[i] == bool(d["done"][i, t]), (name, i, t)
            assert np.isclose(reward[i], d["reward"][i, t], rtol=1e-12, atol=1e-12), \
                (name, i, t, reward[i], d["reward"][i, t])
            exits += int(done[i] and not info["times_up"][i].item())
    if name == "mazes":
        assert exits > 0          # ExtraExitBonus was exercised
    if name == "inaction":        # oscillators: the shadow board is not the starting state any more
        start = _as_i16(np.stack([lv["board"] for lv in levels])).cuda()
        assert not (env.env.baseline_board == start).all()


def test_reset_done_restarts_wrapper_state():
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200 import wrappers
    levels = gu.levels_of(gu.load_npz("wrapped_mazes.npz"))
    n = 6
    base = BatchedSafeLifeEnv(levels, n, auto_reset=False, time_limit=5, view_shape=(5, 5))
    env = wrappers.training_stack(base, se_penalty=1.0)
    env.reset()
    for t in range(5):
        obs, reward, done, info = env.step(np.full(n, 6))      # toggle: creates side effects
    assert done.all()
    assert (env.env.last_side_effect > 0).any()
    obs = env.reset_done()
    assert obs.shape[0] == n
    assert base._t["num_steps"].tolist() == [0] * n
    assert base._t["level_id"].tolist() == [(i + base.level_stride) % len(levels) for i in range(n)]
    assert env.env.last_side_effect.tolist() == [0] * n
    assert (env.env.baseline_board == base._t["board"]).all()
    with pytest.raises(ValueError):
        wrappers.ContinuingEnv(BatchedSafeLifeEnv(levels, 2, auto_reset=True))
    with pytest.raises(ValueError):
        wrappers.MovementBonusWrapper(env)       # one of each kind per stack


def test_batched_rollout_follows_act_on_envs():
    import torch
    from oracle import oracle as orc
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.rollout import BatchedRollout
    levels = gu.levels_of(gu.load_traj("mazes"))
    L, n, T = len(levels), 7, 4
    env = BatchedSafeLifeEnv(levels, n, auto_reset=False, time_limit=T, view_shape=(9, 9),
                             output_channels=None)
    runner = BatchedRollout(env)
    obs0, ids0 = runner.obs_for_envs()
    assert ids0.tolist() == [[i, 0, 0] for i in range(n)]
    ora = orc.OracleEnv([levels[i % L] for i in range(n)], [orc.rng_from_seed(0)] * n, time_limit=T)
    rs = np.random.default_rng(0)
    for t in range(T):
        a = rs.integers(0, 9, n)
        obs, ids = runner.obs_for_envs()
        nxt, reward, done = runner.act_on_envs(a)
        r0, d0 = ora.step(a.astype(np.int64))
        assert np.array_equal(reward.cpu().numpy(), r0) and np.array_equal(done.cpu().numpy(), d0)
        for i in range(n):     # next_obs is the observation of the step, before any reset
            want = orc.get_obs(ora.board[i], ora.goals[i], ora.loc[i], ora.exits[i], (9, 9), None)
            assert np.array_equal(nxt[i].cpu().numpy(), want)
    assert done.all()
    obs, ids = runner.obs_for_envs()
    assert ids.tolist() == [[i, 1, 0] for i in range(n)]
    fresh = orc.OracleEnv([levels[(i + env.level_stride) % L] for i in range(n)],
                          [orc.rng_from_seed(0)] * n)
    for i in range(n):
        want = orc.get_obs(fresh.board[i], fresh.goals[i], fresh.loc[i], fresh.exits[i], (9, 9), None)
        assert np.array_equal(obs[i].cpu().numpy(), want)
    assert not torch.equal(obs, nxt)


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("baseline,ignore", [("starting-state", False), ("inaction", False),
                                             ("starting-state", True)])
def test_wrapped_auto_reset_equals_driven_resets(kernel, baseline, ignore):
    """The in-kernel reset of the fused epilogue leaves the wrapper state exactly as a
    reset_done() from outside does, on the warp path and on the generic path."""
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200 import wrappers
    d = gu.load_npz("wrapped_v1.0.npz")
    levels = gu.levels_of(d)
    n, T = 2 * len(levels) + 3, 40
    rs = np.random.default_rng(4)

    def make(auto):
        base = BatchedSafeLifeEnv(levels, n, auto_reset=auto, time_limit=9, view_shape=(7, 7),
                                  kernel=kernel, seed=5)
        return wrappers.training_stack(base, exit_difficulty=0.3, se_penalty=0.7,
                                       se_baseline=baseline, ignore_reward_cells=ignore)
    a, b = make(True), make(False)
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa, ob)
    for t in range(T):
        act = rs.integers(0, 9, n)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert ra.dtype == torch.float64 and torch.equal(ra, rb), t
        assert torch.equal(da, db), t
        ob = b.reset_done()
        assert torch.equal(oa, ob), t
        assert torch.equal(a.unwrapped.board, b.unwrapped.board), t
        assert torch.equal(a.unwrapped._wrap.wstate, b.unwrapped._wrap.wstate), t
        assert torch.equal(a.unwrapped._wrap.baseline, b.unwrapped._wrap.baseline), t
        assert torch.equal(a.unwrapped._t["required"], b.unwrapped._t["required"]), t
    assert int(a.unwrapped._t["num_steps"].max()) < 9                   # episodes did restart


def test_wrapped_pipelined_kernel_matches_generic_path():
    """The wrapper epilogue inside the persistent pipelined kernel (a batch beyond two
    waves) against the generic path's separate launches."""
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200 import wrappers
    d = gu.load_npz("wrapped_v1.0.npz")
    levels = gu.levels_of(d)
    n = 592 * 8 * 2 + 24
    rs = np.random.default_rng(8)

    def make(kernel):
        base = BatchedSafeLifeEnv(levels, n, auto_reset=True, time_limit=5, view_shape=(5, 5),
                                  kernel=kernel, seed=9)
        return wrappers.training_stack(base, exit_difficulty=0.4, se_penalty=0.3)
    a, b = make(3), make(1)                  # kernel 3: pipelined also with the observation
    b.unwrapped.set_rng_state(a.unwrapped.rng_state)
    assert "k_env_step_pipe" in a.unwrapped.step_plan(obs=True)[0]
    a.reset()
    b.reset()
    for t in range(8):
        act = torch.from_numpy(rs.integers(0, 9, n).astype(np.uint8)).cuda()
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act)
        assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(oa, ob), t
        assert torch.equal(a.unwrapped._wrap.wstate, b.unwrapped._wrap.wstate), t
