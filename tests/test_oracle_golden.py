"""
The oracle (oracle/sl_oracle.c) against the golden vectors generated from the
unmodified reference (tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc


@pytest.fixture(scope="module")
def ops():
    return gu.load_npz("ops.npz")


def _split(flat, sizes):
    out, pos = [], 0
    for s in sizes:
        out.append(flat[pos:pos + s])
        pos += s
    assert pos == len(flat)
    return out


def test_pcg64_matches_numpy():
    for seed in (0, 1, 12345):
        g = np.random.default_rng(seed)
        r = orc.rng_from_generator(g)
        for _ in range(50):
            assert g.random() == orc.next_double(r)
        assert orc.rng_state_int(r)[0] == g.bit_generator.state["state"]["state"]


def test_advance_board_golden(ops):
    meta = ops["adv_shape_n"]
    ins = _split(ops["adv_in"], [h * w for h, w, _ in meta])
    outs = _split(ops["adv_out"], [h * w for h, w, _ in meta])
    for i, (h, w, n) in enumerate(meta):
        rng = ops["adv_rng0"][i].copy()
        res = orc.advance_board(ins[i].reshape(h, w), float(ops["adv_prob"][i]), int(n), rng)
        assert np.array_equal(res.ravel(), outs[i]), i
        assert np.array_equal(rng, ops["adv_rng1"][i]), i


def test_alive_counts_golden(ops):
    sizes = [h * w for h, w in ops["ac_shape"]]
    bs, gs = _split(ops["ac_board"], sizes), _split(ops["ac_goals"], sizes)
    for i in range(len(sizes)):
        assert np.array_equal(orc.alive_counts(bs[i], gs[i]), ops["ac_out"][i])


def test_execute_actions_golden(ops):
    meta = ops["ea_meta"]
    ins = _split(ops["ea_in"], [h * w for h, w, _, _ in meta])
    outs = _split(ops["ea_out"], [h * w for h, w, _, _ in meta])
    li = _split(ops["ea_loc_in"], [2 * na for _, _, na, _ in meta])
    lo = _split(ops["ea_loc_out"], [2 * na for _, _, na, _ in meta])
    acts = _split(ops["ea_act"], [k for _, _, _, k in meta])
    for i, (h, w, na, _) in enumerate(meta):
        b = ins[i].reshape(h, w).copy()
        loc = li[i].reshape(na, 2).copy()
        orc.execute_actions(b, loc, acts[i])
        assert np.array_equal(b.ravel(), outs[i]), i
        assert np.array_equal(loc.ravel(), lo[i]), i


def test_life_occupancy_golden(ops):
    meta = ops["occ_meta"]
    ins = _split(ops["occ_in"], [h * w for h, w, _ in meta])
    outs = _split(ops["occ_out"], [h * w * 8 for h, w, _ in meta])
    for i, (h, w, n) in enumerate(meta):
        rng = ops["occ_rng0"][i].copy()
        res = orc.life_occupancy(ins[i].reshape(h, w), 0.3, int(n), rng)
        assert np.array_equal(res.ravel(), outs[i]), i
        assert np.array_equal(rng, ops["occ_rng1"][i]), i


def test_wrapped_label_golden(ops):
    meta = ops["wl_meta"]
    ins = _split(ops["wl_in"], [h * w for h, w, _ in meta])
    outs = _split(ops["wl_out"], [h * w for h, w, _ in meta])
    for i, (h, w, n) in enumerate(meta):
        lab, num = orc.wrapped_label(ins[i].reshape(h, w))
        assert num == n, i
        assert np.array_equal(lab.ravel(), outs[i]), i


@pytest.mark.parametrize("fixture", ["occupancy.npz", "occupancy_nav50.npz"])
def test_side_effect_occupancy_golden(fixture):
    d = gu.load_npz(fixture)
    ns = int(d["num_samples"])
    for i in range(len(d["init_board"])):
        rng = d["rng0"][i].copy()
        p = float(d["spawn_prob"][i])
        b1 = orc.advance_board(d["init_board"][i], p, int(d["num_steps"][i]), rng)
        assert np.array_equal(b1, d["inaction_board"][i])
        c0 = orc.life_occupancy(b1, p, ns, rng)
        c1 = orc.life_occupancy(d["final_board"][i], p, ns, rng)
        assert np.array_equal(np.stack([c0, c1]), d["counts"][i])
        assert np.array_equal(rng, d["rng1"][i])


@pytest.mark.parametrize("group", gu.traj_groups())
def test_env_trajectories_golden(group):
    """OracleEnv reproduces reference SafeLifeEnv episodes step by step."""
    d = gu.load_traj(group)
    levels = gu.levels_of(d)
    env = orc.OracleEnv(levels, list(d["rng_init"]), time_limit=int(d["time_limit"]))
    ch = tuple(int(c) for c in d["channels"]) or None
    vs = tuple(int(v) for v in d["view_shape"])

    def obs_of(i):
        loc = env.loc[i] if env.has_agent[i] else (0, 0)
        return orc.get_obs(env.board[i], env.goals[i], loc, env.exits[i], vs, ch)

    for i in range(env.n):
        assert gu.crc(env.board[i]) == d["board0_crc"][i]
        assert gu.crc(obs_of(i)) == d["obs0_crc"][i]
    for t in range(d["action"].shape[1]):
        reward, done = env.step(d["action"][:, t].astype(np.int64))
        for i in np.flatnonzero(t < d["length"]):
            where = (group, i, t)
            assert reward[i] == d["reward"][i, t], where
            assert done[i] == bool(d["done"][i, t]), where
            assert env.success[i] == d["success"][i, t], where
            assert gu.crc(env.board[i]) == d["board_crc"][i, t], where
            assert gu.crc(env.goals[i]) == d["goals_crc"][i, t], where
            assert np.array_equal(env.rng[i, :2], d["rng"][i, t]), where
            if env.has_agent[i]:
                assert np.array_equal(env.loc[i], d["loc"][i, t]), where
            assert gu.crc(obs_of(i)) == d["obs_crc"][i, t], where
            if t == d["length"][i] - 1:
                assert np.array_equal(env.board[i], d["final_board"][i]), where
                assert np.array_equal(env.goals[i], d["final_goals"][i]), where
                assert env.static_goals[i] == d["final_static_goals"][i], where
                assert env.ep_reward[i] == d["ep_reward"][i], where
                assert env.ep_length[i] == d["ep_length"][i], where
