"""
Live differential test of the oracle's environment (OracleEnv, get_obs) against
the UNMODIFIED reference Python + C extension, on procgen levels and seeds that
are not among the committed fixtures.  Runs only where the reference checkout is
present (this container); skipped on the GPU box.  No GPU involved.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refload

pytestmark = pytest.mark.skipif(not refload.have_reference(), reason="reference checkout not present")

TRAIN_CHANNELS = (0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 25, 26, 27, 12, 13)   # env_factory.py:311-327


def _level_of(game):
    """The level as the product / oracle see it: deserialised, before update_exit_colors."""
    from safelife.safelife_game import GameState, SafeLifeGame
    raw = SafeLifeGame(board_size=None)
    GameState.deserialize(raw, game._init_data, as_initial_state=False)
    loc = raw.agent_locs[0] if len(raw.agent_locs) else None
    return {
        "board": raw.board.copy(), "goals": np.array(game._init_data["goals"], dtype=np.uint16),
        "agent_loc": None if loc is None else (int(loc[0]), int(loc[1])),
        "points_table": np.asarray(game.points_table, dtype=np.int64)[0],
        "min_performance": float(game.min_performance), "spawn_prob": float(game.spawn_prob),
    }


@pytest.mark.parametrize("yaml_name,shape,seed0,steps,view", [
    ("random/append-still-easy", (25, 25), 901, 120, (25, 25)),
    ("random/prune-still", (25, 25), 902, 120, (15, 15)),
    ("random/append-spawn", (26, 26), 903, 150, (25, 25)),
    ("random/prune-spawn", (19, 23), 904, 100, (9, 31)),
    ("random/append-dynamic", (26, 26), 905, 100, (25, 25)),
])
def test_oracle_env_follows_reference_env(yaml_name, shape, seed0, steps, view):
    refload.load_reference()
    from safelife.level_iterator import _game_from_data, _load_files
    from safelife.safelife_env import SafeLifeEnv
    (fname, kind, data), = _load_files([yaml_name])
    data = dict(data, board_shape=list(shape))
    for k, seed in enumerate(np.random.SeedSequence(seed0).spawn(2)):
        game = _game_from_data(fname, kind, data, seed=seed)
        game.seed = seed
        env = SafeLifeEnv(iter([game]), time_limit=steps, view_shape=view,
                          output_channels=TRAIN_CHANNELS, should_calculate_side_effects=False)
        obs = env.reset()
        ora = orc.OracleEnv([_level_of(game)], [orc.rng_from_generator(game.rng)], time_limit=steps)

        def ora_obs():
            loc = ora.loc[0] if ora.has_agent[0] else (0, 0)
            return orc.get_obs(ora.board[0], ora.goals[0], loc, ora.exits[0], view, TRAIN_CHANNELS)

        assert np.array_equal(ora.board[0], env.game.board)
        assert np.array_equal(ora_obs(), obs)
        arng = np.random.default_rng(seed0 * 10 + k)
        for t in range(steps):
            a = int(arng.integers(0, 9))
            obs, reward, done, info = env.step(a)
            r0, d0 = ora.step(np.array([a], dtype=np.int64))
            where = (yaml_name, k, t)
            assert r0[0] == np.float32(reward), where
            assert bool(d0[0]) == bool(done), where
            assert bool(ora.success[0]) == bool(info["episode"]["success"]), where
            assert np.array_equal(ora.board[0], env.game.board), where
            assert np.array_equal(ora.goals[0], np.asarray(env.game.goals, dtype=np.uint16)), where
            assert np.array_equal(ora.rng[0], orc.rng_from_generator(game.rng)), where
            if len(env.game.agent_locs):
                assert np.array_equal(ora.loc[0], env.game.agent_locs[0]), where
            assert np.array_equal(ora_obs(), obs), where
            if done:
                break
        assert ora.ep_reward[0] == np.float32(env.episode_reward)
        assert ora.ep_length[0] == env.episode_length


@pytest.mark.parametrize("yaml_name,seed0", [("random/append-still-easy", 911), ("random/append-spawn", 912),
                                             ("random/prune-dynamic", 913)])
def test_oracle_side_effect_score_follows_reference(yaml_name, seed0):
    """side_effects.side_effect_score (with the LP stand-in for pyemd) at the end of a short
    random episode: same keys, same inaction totals, same EMD, same RNG consumption."""
    refload.load_reference()
    from safelife.level_iterator import _game_from_data, _load_files
    from safelife.random import set_rng
    from safelife.safelife_env import SafeLifeEnv
    from safelife.side_effects import side_effect_score
    (fname, kind, data), = _load_files([yaml_name])
    data = dict(data, board_shape=[25, 25])
    seed = np.random.SeedSequence(seed0).spawn(1)[0]
    game = _game_from_data(fname, kind, data, seed=seed)
    game.seed = seed
    env = SafeLifeEnv(iter([game]), time_limit=60, should_calculate_side_effects=False)
    env.reset()
    arng = np.random.default_rng(seed0)
    for t in range(60):
        _, _, done, _ = env.step(int(arng.integers(0, 9)))
        if done:
            break
    g = np.random.default_rng(seed0 + 1)
    rng = orc.rng_from_generator(g)
    with set_rng(g):
        want = side_effect_score(env.game, num_samples=150)
    got, _ = orc.side_effect_score(env.game._init_data["board"], env.game.board, env.game.num_steps,
                                   env.game.spawn_prob, rng, 150)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k][0] - want[k][0]) <= 1e-9 * max(1.0, want[k][0]), k
        assert got[k][1] == want[k][1], k
    assert np.array_equal(rng, orc.rng_from_generator(g))
