"""
GPU parity of the batched environment (BatchedSafeLifeEnv -> sl_env_step /
sl_env_reset / sl_get_obs through the C ABI) against the golden reference
trajectories and the oracle.  Bit-exact boards, goals, RNG; exact float32
rewards (integers).
"""
import os

import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

# generic CTA-per-env; warp bit-plane with the goal planes resident in shared
# memory; warp bit-plane with the goal planes left in HBM (forced, so that pools
# with dynamic goals exercise its scratch-slab path)
KERNELS = [1, 2, "2-in-place"]


def _mode(kernel):
    """(kernel id, goals_in_place) of a KERNELS entry."""
    return (2, True) if kernel == "2-in-place" else (kernel, False)


def _make_env(levels, n, time_limit, kernel, rng=None, **kw):
    from safelife_b200.env import BatchedSafeLifeEnv
    H, W = levels[0]["board"].shape
    from safelife_b200 import _lib
    kernel, in_place = _mode(kernel)
    if kernel == 2 and not _lib.lib().sl_warp_kernel_supported(H, W):
        pytest.skip("warp kernel does not handle %dx%d boards" % (H, W))
    env = BatchedSafeLifeEnv(levels, n, time_limit=time_limit, kernel=kernel, auto_reset=False,
                             goals_in_place=in_place, **kw)
    if rng is not None:
        env.set_rng_state(np.asarray(rng, dtype=np.uint64))
    return env


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("group", gu.traj_groups())
def test_env_matches_reference_trajectories(group, kernel):
    d = gu.load_traj(group)
    levels = gu.levels_of(d)
    L = len(levels)
    ch = tuple(int(c) for c in d["channels"]) or None
    vs = tuple(int(v) for v in d["view_shape"])
    env = _make_env(levels, L, int(d["time_limit"]), kernel, rng=d["rng_init"],
                    view_shape=vs, output_channels=ch)
    obs = env.reset().cpu().numpy()
    board = env.board.cpu().numpy()
    for i in range(L):
        assert gu.crc(board[i]) == d["board0_crc"][i], (group, i)
        assert gu.crc(obs[i]) == d["obs0_crc"][i], (group, i)
    for t in range(d["action"].shape[1]):
        obs, reward, done, info = env.step(d["action"][:, t])
        obs, reward, done = obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy()
        board, goals = env.board.cpu().numpy(), env.goals.cpu().numpy()
        loc, rng = env.agent_locs.cpu().numpy()[:, 0], env.rng_state
        success = info["episode"]["success"].cpu().numpy()
        for i in np.flatnonzero(t < d["length"]):
            where = (group, kernel, i, t)
            assert reward[i] == d["reward"][i, t], where
            assert done[i] == bool(d["done"][i, t]), where
            assert success[i] == bool(d["success"][i, t]), where
            assert gu.crc(board[i]) == d["board_crc"][i, t], where
            assert gu.crc(goals[i]) == d["goals_crc"][i, t], where
            assert np.array_equal(rng[i, :2], d["rng"][i, t]), where
            if d["agent_loc"][i][0] >= 0:
                assert np.array_equal(loc[i], d["loc"][i, t]), where
            assert gu.crc(obs[i]) == d["obs_crc"][i, t], where
            if t == d["length"][i] - 1:
                assert np.array_equal(board[i], d["final_board"][i]), where
                assert info["episode"]["reward"][i].item() == d["ep_reward"][i], where
                assert info["episode"]["length"][i].item() == d["ep_length"][i], where


def _random_levels(rs, n, H, W, spawners=True):
    """Synthetic levels with every cell type, an agent, exits and coloured goals."""
    kinds = np.array([0, 0, 0, 0, 0, 0, 9, 9, 1, 16, 32788, 17, 32789, 272, 48, 53, 85, 32884]
                     + ([152, 144] if spawners else []), dtype=np.uint16)
    levels = []
    for _ in range(n):
        b = kinds[rs.integers(0, len(kinds), (H, W))]
        colour = (rs.integers(0, 8, (H, W)) << 9).astype(np.uint16)
        b = np.where((b != 0) & (b != 272), b | colour, b).astype(np.uint16)
        g = ((rs.integers(0, 8, (H, W)) << 9) * (rs.random((H, W)) < 0.3)).astype(np.uint16)
        if rs.random() < 0.3:   # dynamic goals
            g |= (rs.random((H, W)) < 0.2).astype(np.uint16)
        y, x = int(rs.integers(0, H)), int(rs.integers(0, W))
        b[y, x] = 122 | (int(rs.integers(0, 8)) << 9)
        table = rs.integers(-3, 6, (8, 9))
        table[:, 8] = rs.integers(-1, 2, 8) if rs.random() < 0.5 else 0
        levels.append({"board": b, "goals": g, "agent_loc": (y, x), "points_table": table,
                       "min_performance": float(rs.choice([-1, 0.0, 0.01, 0.5])),
                       "spawn_prob": float(rs.choice([0.3, 0.05, 0.9]))})
    return levels


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("shape", [(3, 3), (5, 9), (16, 16), (25, 25), (26, 26), (32, 32), (50, 50),
                                   (33, 40), (20, 50), (47, 21), (64, 64), (33, 33), (70, 30)])
def test_env_random_levels_vs_oracle(shape, kernel):
    H, W = shape
    rs = np.random.default_rng(1000 * H + W)
    n, T = 48, 60
    levels = _random_levels(rs, n, H, W)
    rng = np.stack([orc.rng_from_seed(i) for i in range(n)])
    ora = orc.OracleEnv(levels, list(rng), time_limit=40)
    env = _make_env(levels, n, 40, kernel, rng=rng, view_shape=(9, 11), output_channels=None)
    obs = env.reset().cpu().numpy()
    assert np.array_equal(env.board.cpu().numpy(), ora.board)
    for t in range(T):
        a = rs.integers(0, 9, n)
        r0, d0 = ora.step(a.astype(np.int64))
        obs, r1, d1, info = env.step(a)
        where = (shape, kernel, t)
        assert np.array_equal(env.board.cpu().numpy(), ora.board), where
        assert np.array_equal(env.goals.cpu().numpy(), ora.goals), where
        assert np.array_equal(env.rng_state, ora.rng), where
        assert np.array_equal(env.agent_locs.cpu().numpy()[:, 0], ora.loc), where
        assert np.array_equal(r1.cpu().numpy(), r0), where
        assert np.array_equal(d1.cpu().numpy(), d0), where
        assert np.array_equal(info["times_up"].cpu().numpy(), ora.times_up.astype(bool)), where
        assert np.array_equal(info["episode"]["reward"].cpu().numpy(), ora.ep_reward), where
        assert np.array_equal(info["episode"]["length"].cpu().numpy(), ora.ep_length), where
        assert np.array_equal(env._t["static_goals"].cpu().numpy(), ora.static_goals), where
        obs = obs.cpu().numpy()
        for i in range(0, n, 7):
            want = orc.get_obs(ora.board[i], ora.goals[i], ora.loc[i], ora.exits[i], (9, 11), None)
            assert np.array_equal(obs[i], want), where


@pytest.mark.parametrize("kernel", KERNELS)
def test_auto_reset_cycles_levels(kernel):
    levels = gu.levels_of(gu.load_traj("mazes"))
    L, n = len(levels), 5
    import math
    from safelife_b200.env import BatchedSafeLifeEnv, coprime_stride
    from safelife_b200 import _lib
    kernel, in_place = _mode(kernel)
    env = BatchedSafeLifeEnv(levels, n, time_limit=7, kernel=kernel, view_shape=(5, 5),
                             goals_in_place=in_place)
    env.reset()
    assert env._t["level_id"].tolist() == [0, 1, 2, 3, 4]
    pool_after_reset = None
    for t in range(7):
        obs, r, d, info = env.step(np.zeros(n, dtype=np.int64))
    assert d.all() and info["times_up"].all()
    assert env.level_stride == coprime_stride(n, L) and math.gcd(env.level_stride, L) == 1
    assert env._t["level_id"].tolist() == [(i + env.level_stride) % L for i in range(n)]
    assert env._t["num_steps"].tolist() == [0] * n
    assert info["episode"]["length"].tolist() == [7] * n
    # boards are the fresh levels after update_exit_colors
    ora = orc.OracleEnv([levels[(i + env.level_stride) % L] for i in range(n)],
                        [orc.rng_from_seed(0)] * n, time_limit=7)
    assert np.array_equal(env.board.cpu().numpy(), ora.board)
    obs, r, d, info = env.step(np.ones(n, dtype=np.int64))
    assert info["episode"]["length"].tolist() == [1] * n and not d.any()


def test_no_agent_level_and_argument_checks():
    from safelife_b200.env import BatchedSafeLifeEnv
    b = np.zeros((6, 6), dtype=np.uint16)
    b[1, 1:4] = 9
    level = {"board": b, "goals": np.zeros_like(b), "agent_loc": None,
             "points_table": orc.DEFAULT_POINTS_TABLE, "min_performance": -1, "spawn_prob": 0.3}
    env = BatchedSafeLifeEnv([level], 2, auto_reset=False)
    env.reset()
    obs, r, d, info = env.step([3, 4])
    assert r.tolist() == [0.0, 0.0] and d.tolist() == [True, True]
    want = orc.advance_board(b, 0.3, 1)
    assert np.array_equal(env.board[0].cpu().numpy(), want)
    with pytest.raises(ValueError):
        env.step([1, 2, 3])
    with pytest.raises(ValueError):
        env.step([1, 9])


@pytest.mark.parametrize("kernel", KERNELS)
def test_full_size_properties(kernel):
    """BASELINE-size batch (4096 x 25x25): determinism, shard invariance, oracle spot check."""
    pool_path = os.path.join(gu.GOLDEN, "pools", "append-still-25.npz")
    if not os.path.exists(pool_path):
        pytest.skip("pool fixture missing")
    import torch
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.levels import load_levels
    kernel, in_place = _mode(kernel)
    levels = load_levels(pool_path)
    N, T = 4096, 30
    acts = np.random.default_rng(2).integers(0, 9, (T, N)).astype(np.uint8)

    def run(lo, hi):
        env = BatchedSafeLifeEnv([levels[i % len(levels)] for i in range(lo, hi)], hi - lo,
                                 kernel=kernel, auto_reset=False, seed=1, goals_in_place=in_place)
        from safelife_b200.env import seed_rng_states
        env.set_rng_state(seed_rng_states(1, N)[lo:hi])
        env.reset()
        total = torch.zeros(hi - lo, device="cuda")
        for t in range(T):
            env.step_physics(acts[t, lo:hi])
            total += env._t["reward"]
        return env.board.cpu().numpy(), total.cpu().numpy(), env

    b_all, r_all, env = run(0, N)
    b_again, r_again, _ = run(0, N)
    assert np.array_equal(b_all, b_again) and np.array_equal(r_all, r_again)
    b_lo, r_lo, _ = run(0, N // 2)
    b_hi, r_hi, _ = run(N // 2, N)
    assert np.array_equal(np.concatenate([b_lo, b_hi]), b_all)
    assert np.array_equal(np.concatenate([r_lo, r_hi]), r_all)
    idx = list(range(0, N, 97))
    ora = orc.OracleEnv([levels[i % len(levels)] for i in idx],
                        [orc.rng_from_seed(0)] * len(idx))
    tot = np.zeros(len(idx), dtype=np.float32)
    for t in range(T):
        r, _ = ora.step(acts[t, idx].astype(np.int64))
        tot += r
    assert np.array_equal(ora.board, b_all[idx])
    assert np.array_equal(tot, r_all[idx])


def test_step_host_zero_copy_matches_device_path():
    levels = gu.levels_of(gu.load_traj("procgen25spawn"))
    from safelife_b200.env import BatchedSafeLifeEnv
    n = len(levels)
    rng = np.stack([orc.rng_from_seed(i) for i in range(n)])
    a = BatchedSafeLifeEnv(levels, n, auto_reset=True, time_limit=25)
    b = BatchedSafeLifeEnv(levels, n, auto_reset=True, time_limit=25)
    a.set_rng_state(rng)
    b.set_rng_state(rng)
    a.reset()
    b.reset()
    acts = np.random.default_rng(9).integers(0, 9, (60, n))
    for t in range(60):
        a.step_physics(acts[t], auto_reset=True)
        r, d = b.step_host(acts[t])
        assert np.array_equal(r, a._t["reward"].cpu().numpy())
        assert np.array_equal(d, a._t["done"].cpu().numpy())
        assert np.array_equal(a.board.cpu().numpy(), b.board.cpu().numpy())
    assert d.dtype == np.uint8 and r.dtype == np.float32


@pytest.mark.parametrize("in_place", [False, True])
@pytest.mark.parametrize("n", [592 * 8 + 8, 592 * 8 * 2 + 13, 20011])
def test_pipelined_kernel_large_batches_vs_oracle(n, in_place):
    """Batches large enough for the persistent pipelined kernel (plus a ragged tail)."""
    rs = np.random.default_rng(n)
    base = _random_levels(rs, 64, 25, 25)
    levels = [base[i % 64] for i in range(n)]
    rng = np.stack([orc.rng_from_seed(i % 997) for i in range(n)])
    from safelife_b200.env import BatchedSafeLifeEnv
    env = BatchedSafeLifeEnv(base, n, time_limit=9, auto_reset=True, kernel=2, goals_in_place=in_place)
    env.set_rng_state(rng)
    env.reset()
    idx = np.unique(np.concatenate([np.arange(0, n, 211), [n - 1, n - 2, n - 9, 4735, 4736, 4744]]))
    idx = idx[idx < n]
    ora = orc.OracleEnv([levels[i] for i in idx], list(rng[idx]), time_limit=9)
    live = np.ones(len(idx), dtype=bool)     # the oracle has no auto-reset: follow an env
    for t in range(9):                        # until its first done
        a = rs.integers(0, 9, n)
        r0, d0 = ora.step(a[idx].astype(np.int64))
        env.step_physics(a, auto_reset=True)
        assert np.array_equal(env._t["reward"].cpu().numpy()[idx][live], r0[live]), t
        assert np.array_equal(env._t["done"].cpu().numpy()[idx].astype(bool)[live], d0[live]), t
        cont = live & ~d0
        assert np.array_equal(env.board.cpu().numpy()[idx][cont], ora.board[cont]), t
        assert np.array_equal(env.goals.cpu().numpy()[idx][cont], ora.goals[cont]), t
        assert np.array_equal(env.rng_state[idx][cont], ora.rng[cont]), t
        if t < 8:
            live = cont
    assert live.sum() > len(idx) // 2 and d0[live].all()
    # envs that ran the full episode were auto-reset once: next level, fresh state
    lvl = env._t["level_id"].cpu().numpy()
    sel = idx[live]
    assert np.array_equal(lvl[sel], (sel % 64 + env.level_stride) % 64)
    fresh = orc.OracleEnv([base[l] for l in lvl[sel]], list(rng[sel]), time_limit=9)
    assert np.array_equal(env.board.cpu().numpy()[sel], fresh.board)
    assert np.array_equal(env.goals.cpu().numpy()[sel], fresh.goals)
    assert env._t["num_steps"].cpu().numpy()[sel].max() == 0


# BASELINE.json configs 2-5 as parity cases: procgen pools generated by the
# reference (tests/golden/make_pools.py), a strided subset replayed through the
# oracle with RNG state compared every step.
CONFIGS = [
    ("append-still-25", 4096, 40),
    ("prune-still-25", 16384, 25),
    ("append-spawn-25", 8192, 40),
    ("append-still-26", 2048, 30),
    ("navigation-50", 4096, 25),
]


@pytest.mark.parametrize("pool,n,steps", CONFIGS)
def test_baseline_configs_vs_oracle(pool, n, steps):
    path = os.path.join(gu.GOLDEN, "pools", pool + ".npz")
    if not os.path.exists(path):
        pytest.skip("pool fixture missing")
    from safelife_b200.env import BatchedSafeLifeEnv, seed_rng_states
    from safelife_b200.levels import load_levels
    levels = load_levels(path)
    L = len(levels)
    env = BatchedSafeLifeEnv(levels, n, auto_reset=False, time_limit=1000)
    rng = seed_rng_states(4, n)
    env.set_rng_state(rng)
    env.reset()
    idx = np.arange(0, n, max(1, n // 256))[:256]
    ora = orc.OracleEnv([levels[i % L] for i in idx], list(rng[idx]), time_limit=1000)
    acts = np.random.default_rng(2).integers(0, 9, (steps, n)).astype(np.uint8)
    draws = 0
    for t in range(steps):
        before = ora.rng.copy()
        r0, d0 = ora.step(acts[t, idx].astype(np.int64))
        draws += int((before[:, 1] != ora.rng[:, 1]).sum())
        env.step_physics(acts[t])
        where = (pool, t)
        assert np.array_equal(env._t["reward"].cpu().numpy()[idx], r0), where
        assert np.array_equal(env._t["done"].cpu().numpy()[idx].astype(bool), d0), where
        assert np.array_equal(env.rng_state[idx], ora.rng), where
        assert np.array_equal(env.agent_locs.cpu().numpy()[idx, 0], ora.loc), where
        if t % 5 == 4 or t == steps - 1:
            assert np.array_equal(env.board.cpu().numpy()[idx], ora.board), where
            assert np.array_equal(env.goals.cpu().numpy()[idx], ora.goals), where
    sg = env._t["static_goals"].cpu().numpy()
    assert np.array_equal(sg[idx], ora.static_goals)
    if "spawn" in pool or "navigation" in pool:
        assert draws > 0                      # the stochastic path was exercised
    if "navigation" in pool:
        assert (sg == 2).any()                # dynamic goals: two stencils per step


def test_step_plan_and_launch_floor():
    """sl_env_step_plan names what sl_env_step launches; sl_launch_floor launches an empty kernel."""
    import torch
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(3)
    base = _random_levels(rs, 16, 25, 25)
    small = BatchedSafeLifeEnv(base, 64, kernel=2, goals_in_place=False)
    text, launches = small.step_plan()
    assert "k_env_step_warp<Narrow, goals resident>" in text and launches == 1
    big = BatchedSafeLifeEnv(base, 592 * 8 * 2 + 8 * 40 + 3, kernel=2, goals_in_place=True)
    text, launches = big.step_plan()
    assert "k_env_step_pipe<" in text and "goals in place" in text and "k_env_step_warp" in text
    assert launches == 2                      # pipelined kernel + the ragged tail group
    tiled = BatchedSafeLifeEnv(_random_levels(rs, 4, 70, 41), 8)
    assert "k_env_step_tile<3 x 2 tiles, 1 envs/CTA>" in tiled.step_plan()[0]
    generic = BatchedSafeLifeEnv(_random_levels(rs, 2, 130, 41), 4)
    assert "k_env_step_generic" in generic.step_plan()[0]
    assert _lib.lib().sl_launch_floor(small._batch_ref, small._stream()) == _lib.SL_OK
    torch.cuda.synchronize()


# ---- the observation written by the step launch itself (sl_env_step_obs) ---------
TRAIN_CHANNELS = (0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 25, 26, 27, 12, 13)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("shape,view,channels", [
    ((25, 25), (25, 25), TRAIN_CHANNELS),          # the training observation
    ((25, 25), (15, 15), tuple(range(16)) + (25, 26, 27)),   # the reference's defaults (19 channels)
    ((26, 26), (25, 25), TRAIN_CHANNELS),
    ((9, 7), (11, 40), (3, 1, 0, 9, 25)),          # view larger than the board (torus tiles), 5 channels
    ((50, 50), (25, 25), TRAIN_CHANNELS),          # Wide geometry
    ((40, 33), (7, 9), None),                      # raw uint32 view
    ((25, 25), (5, 5), (4,)),                      # one channel
    ((16, 16), (9, 9), tuple(range(32))),          # all 32 bits
])
def test_fused_observation_matches_oracle(shape, view, channels, kernel):
    """step() returns the observation of the state it leaves behind, for every env,
    with and without auto-reset; checked against the oracle's get_obs."""
    H, W = shape
    rs = np.random.default_rng(7 * H + W + len(view))
    n, T = 40, 24
    levels = _random_levels(rs, 12, H, W)
    rng = np.stack([orc.rng_from_seed(i) for i in range(n)])
    for auto in (False, True):
        from safelife_b200.env import BatchedSafeLifeEnv
        k, in_place = _mode(kernel)
        from safelife_b200 import _lib
        if k == 2 and not _lib.lib().sl_warp_kernel_supported(H, W):
            pytest.skip("warp kernel does not handle %dx%d boards" % (H, W))
        env = BatchedSafeLifeEnv(levels, n, time_limit=9, kernel=k, auto_reset=auto,
                                 goals_in_place=in_place, view_shape=view, output_channels=channels)
        env.set_rng_state(rng)
        env.reset()
        for t in range(T):
            a = rs.integers(0, 9, n)
            obs, r, d, info = env.step(a)
            obs = obs.cpu().numpy()
            board, goals = env.board.cpu().numpy(), env.goals.cpu().numpy()
            loc, lvl = env.agent_locs.cpu().numpy()[:, 0], env._t["level_id"].cpu().numpy()
            for i in range(n):
                y0, x0 = (0, 0) if loc[i][0] < 0 else loc[i]
                want = orc.get_obs(board[i], goals[i], (y0, x0), env.pool.exits[lvl[i]], view, channels)
                assert np.array_equal(obs[i], want), (shape, view, kernel, auto, t, i)
            # and it is what the stand-alone writer produces
            assert np.array_equal(env.get_obs().cpu().numpy(), obs)
            if not auto and t % 9 == 8:
                env.reset_done()


@pytest.mark.parametrize("in_place", [False, True])
def test_fused_observation_pipelined_kernel(in_place):
    """A batch large enough for k_env_step_pipe<..., obs fused> plus a ragged tail."""
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(5)
    n = 592 * 8 * 2 + 8 * 3 + 5
    base = _random_levels(rs, 32, 25, 25)
    # kernel 3: the pipelined kernel also with the observation in the launch (the default
    # plan then prefers the single-pass kernel, which measured faster)
    env = BatchedSafeLifeEnv(base, n, time_limit=6, auto_reset=True, kernel=3, goals_in_place=in_place,
                             view_shape=(25, 25), output_channels=TRAIN_CHANNELS)
    ref = BatchedSafeLifeEnv(base, n, time_limit=6, auto_reset=True, kernel=2, goals_in_place=in_place,
                             view_shape=(25, 25), output_channels=TRAIN_CHANNELS)
    plan, launches = env.step_plan(obs=True)
    assert "k_env_step_pipe<" in plan and "obs fused" in plan and launches == 2
    rng = np.stack([orc.rng_from_seed(i % 499) for i in range(n)])
    env.set_rng_state(rng)
    ref.set_rng_state(rng)
    env.reset()
    ref.reset()
    for t in range(14):
        a = torch.from_numpy(rs.integers(0, 9, n).astype(np.uint8)).cuda()
        obs, r, d, info = env.step(a)
        ref.step_physics(a, auto_reset=True)            # physics only, then the stand-alone writer
        assert torch.equal(env.board, ref.board), t
        assert torch.equal(r, ref._t["reward"]), t
        assert torch.equal(obs, ref.get_obs()), t


def test_step_outputs_survive_the_next_step_and_obs_out():
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(11)
    levels = _random_levels(rs, 8, 25, 25)
    env = BatchedSafeLifeEnv(levels, 64, time_limit=50, auto_reset=True)
    env.reset()
    a = torch.from_numpy(rs.integers(0, 9, (3, 64)).astype(np.uint8)).cuda()
    o0, r0, d0, i0 = env.step(a[0])
    keep = [x.clone() for x in (o0, r0, d0, i0["times_up"], i0["episode"]["success"])]
    o1, r1, d1, i1 = env.step(a[1])
    for held, want in zip((o0, r0, d0, i0["times_up"], i0["episode"]["success"]), keep):
        assert torch.equal(held, want)            # step t's tensors are intact after step t+1
    assert o1.data_ptr() != o0.data_ptr() and r1.data_ptr() != r0.data_ptr()
    buf = torch.zeros_like(o1)
    o2, r2, d2, i2 = env.step(a[2], obs_out=buf)
    assert o2.data_ptr() == buf.data_ptr() and torch.equal(buf, env.get_obs())
    with pytest.raises(ValueError):
        env.step(a[2], obs_out=torch.zeros(3, device="cuda"))


@pytest.mark.parametrize("pool,n", [("append-still-25", 4096), ("prune-still-25", 16384),
                                    ("append-spawn-25", 8192), ("navigation-50", 4096)])
def test_consecutive_episodes_get_different_levels(pool, n):
    """BASELINE batch sizes are multiples of the pool sizes: the level stride must not be 0."""
    import math
    path = os.path.join(gu.GOLDEN, "pools", pool + ".npz")
    if not os.path.exists(path):
        pytest.skip("pool fixture missing")
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.levels import load_levels
    levels = load_levels(path)
    L = len(levels)
    assert n % L == 0
    env = BatchedSafeLifeEnv(levels, n, time_limit=3, auto_reset=True)
    assert env.level_stride != 0 and math.gcd(env.level_stride, L) == 1
    env.reset()
    seen = [env._t["level_id"].cpu().numpy().copy()]
    for t in range(6):                               # two full episodes of 3 steps
        env.step_physics(np.zeros(n, dtype=np.uint8), auto_reset=True)
        if t % 3 == 2:
            seen.append(env._t["level_id"].cpu().numpy().copy())
    assert (seen[1] != seen[0]).all() and (seen[2] != seen[1]).all()
    assert np.array_equal(seen[1], (seen[0] + env.level_stride) % L)


def test_update_pool_hot_swap():
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(21)
    levels = _random_levels(rs, 16, 25, 25, spawners=False)
    for lv in levels:
        lv["goals"] = (lv["goals"] & 0x0E00).astype(np.uint16)      # static goals
        lv["points_table"] = orc.DEFAULT_POINTS_TABLE
    n = 4
    env = BatchedSafeLifeEnv(levels, n, time_limit=5, auto_reset=True, kernel=2)
    env.reset()
    stride = env.level_stride
    nxt = [(i + stride) % 16 for i in range(n)]
    with pytest.raises(ValueError):
        env.update_pool([0], [levels[3]])                            # slot 0 is being played
    fresh = _random_levels(rs, n, 25, 25, spawners=False)
    for lv in fresh:
        lv["points_table"] = orc.DEFAULT_POINTS_TABLE
    fresh[1]["goals"] = (fresh[1]["goals"] | 1).astype(np.uint16)   # one with dynamic goals
    env.update_pool(nxt, fresh)
    assert not env._batch.pool_flags & 2                            # the static-goals promise is gone
    a = np.zeros(n, dtype=np.uint8)
    for t in range(5):
        obs, r, d, info = env.step(a)
    assert d.all()
    ora = orc.OracleEnv(fresh, [orc.rng_from_seed(0)] * n, time_limit=5)
    assert np.array_equal(env.board.cpu().numpy(), ora.board)
    assert np.array_equal(env.goals.cpu().numpy(), ora.goals)
    rng = env.rng_state
    ora2 = orc.OracleEnv(fresh, list(rng), time_limit=5)
    for t in range(4):
        acts = rs.integers(0, 9, n)
        r0, d0 = ora2.step(acts.astype(np.int64))
        obs, r1, d1, info = env.step(acts)
        assert np.array_equal(r1.cpu().numpy(), r0) and np.array_equal(env.board.cpu().numpy(), ora2.board)
        assert np.array_equal(env.goals.cpu().numpy(), ora2.goals)


def test_step_is_cuda_graph_capturable():
    """sl_env_step_obs (physics + auto-reset + observation) captured in a CUDA graph and
    replayed equals the same steps launched one by one."""
    import ctypes
    import torch
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(17)
    levels = _random_levels(rs, 16, 25, 25)
    n, K = 256, 6
    rng = np.stack([orc.rng_from_seed(i) for i in range(n)])
    envs = []
    for _ in range(2):
        e = BatchedSafeLifeEnv(levels, n, time_limit=7, auto_reset=True, view_shape=(25, 25),
                               output_channels=TRAIN_CHANNELS)
        e.set_rng_state(rng)
        e.reset()
        envs.append(e)
    a, b = envs
    acts = torch.from_numpy(rs.integers(0, 9, (K, n)).astype(np.uint8)).cuda()
    obs = torch.zeros((K,) + tuple(a._obs_bufs[0].shape), dtype=torch.uint8, device="cuda")
    specs = [a._make_obs_spec(obs[k]) for k in range(K)]
    a.step(acts[0])                       # warm-up launch outside the capture (sets the
    b.step(acts[0])                       # kernel's shared-memory attribute once)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(K):
            _lib.check(a._lib.sl_env_step_obs(a._batch_ref, acts[k].data_ptr(), ctypes.byref(specs[k]),
                                              0, _lib.SL_STEP_AUTO_RESET, a._stream()))
    for rep in range(3):                  # 18 steps: crosses episode boundaries (time_limit 7)
        g.replay()
        for k in range(K):
            o, r, d, info = b.step(acts[k])
            assert torch.equal(o, obs[k]), (rep, k)
        torch.cuda.synchronize()
        assert torch.equal(a.board, b.board) and torch.equal(a._t["rng"], b._t["rng"]), rep
        assert torch.equal(a._t["ep_reward"], b._t["ep_reward"]), rep


def test_step_host_reports_timing_and_both_wait_modes():
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(23)
    levels = _random_levels(rs, 8, 25, 25)
    env = BatchedSafeLifeEnv(levels, 128, time_limit=20, auto_reset=True, seed=3)
    ref = BatchedSafeLifeEnv(levels, 128, time_limit=20, auto_reset=True, seed=3)
    env.set_rng_state(ref.rng_state)
    env.reset()
    ref.reset()
    timing = np.zeros(2)
    for t in range(30):
        a = rs.integers(0, 9, 128)
        r, d = env.step_host(a, timing=timing)
        out = ref.step_physics(a, auto_reset=True)
        assert np.array_equal(r, out["reward"].cpu().numpy()) and np.array_equal(d, out["done"].cpu().numpy())
        assert timing[0] > 0 and timing[1] >= 0


def test_host_groups_pipeline_matches_whole_batch_steps():
    """enable_host_groups: two half batches stepping on their own streams give exactly the
    results of whole-batch steps with the same actions."""
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(29)
    levels = _random_levels(rs, 8, 25, 25)
    n, T = 128, 25
    a = BatchedSafeLifeEnv(levels, n, time_limit=11, auto_reset=True, seed=3)
    b = BatchedSafeLifeEnv(levels, n, time_limit=11, auto_reset=True, seed=3)
    b.set_rng_state(a.rng_state)
    a.reset()
    b.reset()
    views = a.enable_host_groups(2)
    assert len(views) == 2 and views[1]["reward"].shape == (n // 2,)
    acts = rs.integers(0, 9, (T, n)).astype(np.uint8)
    half = n // 2
    for g in range(2):
        a.host_submit(g, acts[0][g * half:(g + 1) * half])
    for t in range(T):
        out = b.step_physics(acts[t], auto_reset=True)
        want_r, want_d = out["reward"].cpu().numpy(), out["done"].cpu().numpy()
        for g in range(2):
            r, d = a.host_wait(g)
            assert np.array_equal(r, want_r[g * half:(g + 1) * half]), (t, g)
            assert np.array_equal(d, want_d[g * half:(g + 1) * half]), (t, g)
            if t + 1 < T:
                a.host_submit(g, acts[t + 1][g * half:(g + 1) * half])
    a.host_sync_groups()
    assert np.array_equal(a.board.cpu().numpy(), b.board.cpu().numpy())
    with pytest.raises(ValueError):
        a.enable_host_groups(3)


@pytest.mark.parametrize("pool,n", [("append-still-25", 4096), ("prune-still-25", 16384),
                                    ("append-spawn-25", 8192), ("navigation-50", 4096)])
def test_baseline_pools_across_an_episode_boundary(pool, n):
    """BASELINE batch sizes with auto-reset: a strided subset follows the oracle through the
    end of one episode, the in-kernel reset onto the next level, and the episode after it."""
    path = os.path.join(gu.GOLDEN, "pools", pool + ".npz")
    if not os.path.exists(path):
        pytest.skip("pool fixture missing")
    from safelife_b200.env import BatchedSafeLifeEnv, seed_rng_states
    from safelife_b200.levels import load_levels
    levels = load_levels(path)
    L, limit = len(levels), 8
    env = BatchedSafeLifeEnv(levels, n, auto_reset=True, time_limit=limit, view_shape=(25, 25),
                             output_channels=TRAIN_CHANNELS)
    rng = seed_rng_states(6, n)
    env.set_rng_state(rng)
    env.reset()
    idx = np.arange(0, n, max(1, n // 96))[:96]
    acts = np.random.default_rng(3).integers(0, 9, (2 * limit, n)).astype(np.uint8)
    level_of = idx % L
    state = rng[idx]
    for episode in range(2):
        ora = orc.OracleEnv([levels[l] for l in level_of], list(state), time_limit=limit)
        live = np.ones(len(idx), dtype=bool)
        for t in range(limit):
            a = acts[episode * limit + t]
            r0, d0 = ora.step(a[idx].astype(np.int64))
            obs, r1, d1, info = env.step(a)
            where = (pool, episode, t)
            assert np.array_equal(r1.cpu().numpy()[idx][live], r0[live]), where
            assert np.array_equal(d1.cpu().numpy()[idx][live], d0[live]), where
            cont = live & ~d0
            assert np.array_equal(env.board.cpu().numpy()[idx][cont], ora.board[cont]), where
            assert np.array_equal(env.goals.cpu().numpy()[idx][cont], ora.goals[cont]), where
            assert np.array_equal(env.rng_state[idx][cont], ora.rng[cont]), where
            if t < limit - 1:
                live = cont          # an env that left early has already been reset: stop following it
        assert d0[live].all() and live.sum() > len(idx) // 2
        # every followed env was reset by the step kernel onto its next level
        idx, level_of = idx[live], (level_of[live] + env.level_stride) % L
        assert np.array_equal(env._t["level_id"].cpu().numpy()[idx], level_of), pool
        assert env._t["num_steps"].cpu().numpy()[idx].max() == 0
        state = env.rng_state[idx]
        fresh = orc.OracleEnv([levels[l] for l in level_of], list(state), time_limit=limit)
        assert np.array_equal(env.board.cpu().numpy()[idx], fresh.board), pool
        obs = obs.cpu().numpy()
        for k in range(0, len(idx), 9):     # the observation of a finished env is its next episode's first
            want = orc.get_obs(fresh.board[k], fresh.goals[k], fresh.loc[k], fresh.exits[k], (25, 25),
                               TRAIN_CHANNELS)
            assert np.array_equal(obs[idx[k]], want), (pool, k)


@pytest.mark.parametrize("shape,n", [((50, 50), 5), ((70, 30), 3), ((33, 33), 7), ((64, 64), 2)])
def test_tile_kernel_ragged_batches_match_generic_kernel(shape, n):
    """Batches that leave a CTA of the tiled kernel partly empty (and shapes whose env
    groups cannot be bulk-copied), with the observation, against the generic kernel."""
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    H, W = shape
    rs = np.random.default_rng(H + W + n)
    levels = _random_levels(rs, 4, H, W)
    rng = np.stack([orc.rng_from_seed(i) for i in range(n)])
    envs = []
    for kernel in (2, 1):
        e = BatchedSafeLifeEnv(levels, n, time_limit=6, auto_reset=True, kernel=kernel,
                               view_shape=(25, 25), output_channels=TRAIN_CHANNELS)
        e.set_rng_state(rng)
        e.reset()
        envs.append(e)
    a, b = envs
    assert "k_env_step_tile" in a.step_plan(obs=True)[0]
    for t in range(15):
        act = torch.from_numpy(rs.integers(0, 9, n).astype(np.uint8)).cuda()
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), (shape, t)
        assert torch.equal(a.board, b.board) and torch.equal(a.goals, b.goals), (shape, t)
        assert torch.equal(a._t["rng"], b._t["rng"]), (shape, t)


@pytest.mark.parametrize("n", [5, 600, 1500, 2000, 2900, 3500, 4096, 4700, 8192])
@pytest.mark.parametrize("in_place", [True, False])
def test_single_pass_kernel_any_envs_per_cta_matches_generic_kernel(n, in_place):
    """The single-pass kernel spreads a batch evenly over the SMs with 1..8 envs per CTA
    (a CTA's slabs then start anywhere inside a 16-byte unit of the board array: bulk
    loads of the enclosing units, bulk stores of the inner ones, edges by cell), with
    and without the observation in the launch, against the generic kernel."""
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(n)
    levels = _random_levels(rs, 8, 25, 25)
    rng = np.stack([orc.rng_from_seed(i) for i in range(n)])
    envs = []
    for kernel in (2, 1):
        e = BatchedSafeLifeEnv(levels, n, time_limit=5, auto_reset=True, kernel=kernel,
                               view_shape=(9, 9), output_channels=TRAIN_CHANNELS,
                               goals_in_place=in_place)
        e.set_rng_state(rng)
        e.reset()
        envs.append(e)
    a, b = envs
    assert "k_env_step_warp" in a.step_plan()[0]
    for t in range(8):
        act = torch.from_numpy(rs.integers(0, 9, n).astype(np.uint8)).cuda()
        if t % 2:
            oa, ra, da, _ = a.step(act)
            ob, rb, db, _ = b.step(act)
            assert torch.equal(oa, ob), t
        else:
            ra, da = (a.step_physics(act, auto_reset=True)[k] for k in ("reward", "done"))
            rb, db = (b.step_physics(act, auto_reset=True)[k] for k in ("reward", "done"))
        assert torch.equal(ra, rb) and torch.equal(da, db), t
        assert torch.equal(a.board, b.board) and torch.equal(a.goals, b.goals), t
        assert torch.equal(a._t["rng"], b._t["rng"]), t


@pytest.mark.parametrize("n,in_place", [(64, True), (4096, True), (1000, False), (4700, False)])
def test_host_session_matches_launch_per_step(n, in_place):
    """A host session (one resident kernel, actions and results through pinned-memory
    mailboxes) steps exactly like one launch per step -- rewards, flags, boards, goals,
    RNG -- across auto-resets, across an idle timeout that ends the session kernel, and
    when the session is closed by touching the env."""
    import time
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    rs = np.random.default_rng(n)
    levels = _random_levels(rs, 8, 25, 25)
    rng = np.stack([orc.rng_from_seed(i % 311) for i in range(n)])
    envs = []
    for _ in range(2):
        e = BatchedSafeLifeEnv(levels, n, time_limit=7, auto_reset=True, kernel=2, goals_in_place=in_place)
        e.set_rng_state(rng)
        e.reset()
        envs.append(e)
    a, b = envs
    a.begin_host_session(idle_timeout_ms=50)
    for t in range(40):
        act = rs.integers(0, 9, n).astype(np.uint8)
        if t == 20:
            time.sleep(0.2)                       # the session kernel gives up; the next step reopens it
        if t == 30:
            assert torch.equal(a.board, b.board)  # closes the session; the next step runs as a launch
            a.begin_host_session(idle_timeout_ms=50)
        ra, da = a.step_host(act)
        rb, db = b.step_host(act)
        assert np.array_equal(ra, rb) and np.array_equal(da, db), t
        assert np.array_equal(a._host_np["success"], b._host_np["success"]), t
        assert np.array_equal(a._host_np["times_up"], b._host_np["times_up"]), t
    a.end_host_session()
    assert torch.equal(a.board, b.board) and torch.equal(a.goals, b.goals)
    assert torch.equal(a._t["state"], b._t["state"])
