"""
Generate golden fixtures by running the UNMODIFIED reference (Python from
/root/reference + its C extension compiled into oracle/_ref) in this
container.  The outputs are committed under tests/golden/ and are what the
oracle is pinned against; the reference itself never travels to the GPU box.

    python tests/golden/make_golden.py            # all fixture groups

Fixture groups written (np.savez_compressed):
  traj_<group>.npz   one episode per level through reference SafeLifeEnv
                     (safelife_env.py:148-218) with seeded random actions:
                     initial level data, PCG64 state, per-step reward / done /
                     success / agent loc / crc32(board) / crc32(goals) /
                     crc32(obs) / PCG64 state, and the final board + goals.
                     Groups: benchmark / puzzle / maze level files and procgen levels
                     of every BASELINE pool (25x25 append-still / append-spawn /
                     prune-still, 50x50 navigation, 26x26 append-still).
  ops.npz            direct calls of the reference C functions on random
                     bit-pattern boards (advance_board, alive_counts,
                     execute_actions, life_occupancy, wrapped_label).
  occupancy.npz      the device part of side_effect_score
                     (side_effects.py:101-111) at episode end.
  side_effects.npz   the reference's whole side_effect_score (side_effects.py:60-154,
                     num_samples=200) at the end of short random episodes, under
                     set_rng; pyemd is absent here, so the EMD inside it is the
                     exact LP of oracle/refload.py (parity unpinned at pyemd).
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from oracle import oracle as orc  # noqa: E402

sl = refload.load_reference()
from safelife import speedups  # noqa: E402
from safelife.level_iterator import SafeLifeLevelIterator, _load_files, _game_from_data  # noqa: E402
from safelife.random import set_rng  # noqa: E402
from safelife.safelife_env import SafeLifeEnv  # noqa: E402
from safelife.safelife_game import SafeLifeGame, CellTypes  # noqa: E402

TRAIN_CHANNELS = (0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 25, 26, 27, 12, 13)  # env_factory.py:311-327


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def level_record(game):
    """
    The level as GameState.deserialize leaves it (safelife_game.py:211-234:
    board copy, agent location, `orientation` key applied) BEFORE
    GameWithGoals.deserialize runs update_exit_colors (:622-637) -- this is the
    state on which exit_locs and initial_counts are taken.
    """
    from safelife.safelife_game import GameState
    raw = SafeLifeGame(board_size=None)
    GameState.deserialize(raw, game._init_data, as_initial_state=False)
    table = np.asarray(game.points_table, dtype=np.int64)
    assert table.shape[0] <= 1 or len(game.agent_locs) == table.shape[0]
    loc = raw.agent_locs[0] if len(raw.agent_locs) else np.array([-1, -1])
    return {
        "board": raw.board.copy(),
        "goals": np.array(game._init_data["goals"], dtype=np.uint16),
        "agent_loc": np.asarray(loc, dtype=np.int32),
        "points_table": table[0] if len(table) else orc.DEFAULT_POINTS_TABLE,
        "min_performance": float(game.min_performance),
        "spawn_prob": float(game.spawn_prob),
    }


def run_episode(game, seed, time_limit, view_shape, channels, action_seed):
    """One reference episode; returns the level record and per-step logs."""
    game.seed = seed
    env = SafeLifeEnv(iter([game]), time_limit=time_limit, view_shape=view_shape,
                      output_channels=channels, should_calculate_side_effects=False)
    obs = env.reset()
    # State as the product / oracle must see it at reset: the level as
    # deserialised, before update_exit_colors (safelife_game.py:622-637).
    rec = level_record(game)
    rec["rng_init"] = orc.rng_from_generator(game.rng)
    rec["obs0_crc"] = crc(obs)
    rec["board0_crc"] = crc(env.game.board)
    arng = np.random.default_rng(action_seed)
    log = {k: [] for k in ("action", "reward", "done", "success", "loc", "board_crc",
                           "goals_crc", "obs_crc", "rng")}
    for t in range(time_limit):
        a = int(arng.integers(0, 9))
        obs, reward, done, info = env.step(a)
        loc = env.game.agent_locs[0] if len(env.game.agent_locs) else (-1, -1)
        log["action"].append(a)
        log["reward"].append(np.float32(reward))
        log["done"].append(bool(done))
        log["success"].append(bool(info["episode"]["success"]))
        log["loc"].append(np.asarray(loc, dtype=np.int32))
        log["board_crc"].append(crc(env.game.board))
        log["goals_crc"].append(crc(np.asarray(env.game.goals, dtype=np.uint16)))
        log["obs_crc"].append(crc(obs))
        log["rng"].append(orc.rng_from_generator(game.rng)[:2])
        if done:
            break
    rec["final_board"] = env.game.board.copy()
    rec["final_goals"] = np.asarray(env.game.goals, dtype=np.uint16).copy()
    rec["final_static_goals"] = np.uint8(1 if env.game._static_goals else 2)
    rec["ep_reward"] = np.float32(env.episode_reward)
    rec["ep_length"] = np.int32(env.episode_length)
    return rec, log, env.game


def run_wrapped_episode(game, seed, time_limit, action_seed, se_penalty, exit_difficulty,
                        baseline="starting-state", ignore_reward_cells=False):
    """One episode through the training wrapper stack of training/env_factory.py:274-283."""
    from safelife import env_wrappers
    game.seed = seed
    env = SafeLifeEnv(iter([game]), time_limit=time_limit, view_shape=(5, 5),
                      output_channels=None, should_calculate_side_effects=False)
    env = env_wrappers.MovementBonusWrapper(env, as_penalty=True)
    env = env_wrappers.ExtraExitBonus(env)
    env = env_wrappers.SimpleSideEffectPenalty(env, baseline=baseline, penalty_coef=se_penalty,
                                               ignore_reward_cells=ignore_reward_cells)
    env = env_wrappers.MinPerformanceScheduler(env, min_performance_fraction=exit_difficulty)
    rec = level_record(game)        # before reset: the scheduler rescales min_performance
    env.reset()
    rec["rng_init"] = orc.rng_from_generator(game.rng)
    arng = np.random.default_rng(action_seed)
    log = {"action": [], "reward": [], "done": []}
    for t in range(time_limit):
        a = int(arng.integers(0, 9))
        obs, reward, done, info = env.step(a)
        log["action"].append(a)
        log["reward"].append(np.float64(reward))
        log["done"].append(bool(done))
        if done:
            break
    return rec, log


def group_wrapped(name, paths, limit, time_limit, seed0, se_penalty, exit_difficulty,
                  min_performance=None, baseline="starting-state", ignore_reward_cells=False):
    recs, logs = [], []
    ss = np.random.SeedSequence(seed0)
    for path in paths:
        for fname, data in static_games(path, limit):
            game = SafeLifeGame.loaddata(data, auto_cls=False)
            if min_performance is not None:
                game.min_performance = min_performance
                game._init_data = game.serialize()
            rec, log = run_wrapped_episode(game, ss.spawn(1)[0], time_limit,
                                           len(recs) + 1000 * seed0, se_penalty, exit_difficulty,
                                           baseline, ignore_reward_cells)
            recs.append(rec)
            logs.append(log)
    L, T = len(recs), max(len(lg["action"]) for lg in logs)
    out = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    out["length"] = np.array([len(lg["action"]) for lg in logs], dtype=np.int32)
    for key, dtype in (("action", np.uint8), ("reward", np.float64), ("done", np.uint8)):
        arr = np.zeros((L, T), dtype=dtype)
        for i, lg in enumerate(logs):
            arr[i, :len(lg[key])] = np.asarray(lg[key], dtype=dtype)
        out[key] = arr
    out["time_limit"] = np.int32(time_limit)
    out["se_penalty"] = np.float64(se_penalty)
    out["exit_difficulty"] = np.float64(exit_difficulty)
    out["baseline"] = np.array(baseline)
    out["ignore_reward_cells"] = np.uint8(ignore_reward_cells)
    path = os.path.join(HERE, "wrapped_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("wrote", path, "levels", L, "max steps", T, os.path.getsize(path) // 1024, "KiB")


def save_group(name, recs, logs, time_limit, view_shape, channels):
    L = len(recs)
    T = max(len(lg["action"]) for lg in logs)
    out = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    out["length"] = np.array([len(lg["action"]) for lg in logs], dtype=np.int32)

    def pad(key, dtype, tail=()):
        arr = np.zeros((L, T) + tail, dtype=dtype)
        for i, lg in enumerate(logs):
            n = len(lg[key])
            arr[i, :n] = np.asarray(lg[key], dtype=dtype).reshape((n,) + tail)
        return arr

    out["action"] = pad("action", np.uint8)
    out["reward"] = pad("reward", np.float32)
    out["done"] = pad("done", np.uint8)
    out["success"] = pad("success", np.uint8)
    out["loc"] = pad("loc", np.int32, (2,))
    out["board_crc"] = pad("board_crc", np.uint32)
    out["goals_crc"] = pad("goals_crc", np.uint32)
    out["obs_crc"] = pad("obs_crc", np.uint32)
    out["rng"] = pad("rng", np.uint64, (2,))
    out["time_limit"] = np.int32(time_limit)
    out["view_shape"] = np.array(view_shape, dtype=np.int32)
    out["channels"] = np.array(channels if channels else [], dtype=np.int32)
    path = os.path.join(HERE, "traj_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("wrote", path, "levels", L, "max steps", T, os.path.getsize(path) // 1024, "KiB")


def static_games(path, limit=None, only_safelife=True):
    games = []
    for fname, kind, data in _load_files([path]):
        keys = data.dtype.fields if hasattr(data, "dtype") else data
        cls = str(data["class"]) if "class" in keys else "SafeLifeGame"
        if only_safelife and not cls.endswith("SafeLifeGame"):
            continue   # GameOfLife-tagged files run a different physics
        games.append((fname, data))
        if limit and len(games) >= limit:
            break
    return games


def group_static(name, paths, limit, time_limit, view_shape, channels, seed0,
                 min_performance=None):
    recs, logs = [], []
    ss = np.random.SeedSequence(seed0)
    for path in paths:
        for fname, data in static_games(path, limit):
            game = SafeLifeGame.loaddata(data, auto_cls=False)
            if min_performance is not None:
                # as MinPerformanceScheduler would (env_wrappers.py:131-145), but
                # stored in the level so that reset sees it
                game.min_performance = min_performance
                game._init_data = game.serialize()
            seed = ss.spawn(1)[0]
            rec, log, _ = run_episode(game, seed, time_limit, view_shape, channels,
                                      action_seed=len(recs) + 1000 * seed0)
            recs.append(rec)
            logs.append(log)
    save_group(name, recs, logs, time_limit, view_shape, channels)


def group_procgen(name, yaml_name, shape, count, time_limit, view_shape, channels, seed0):
    recs, logs = [], []
    (fname, kind, data), = _load_files([yaml_name])
    data = dict(data, board_shape=list(shape))
    for seed in np.random.SeedSequence(seed0).spawn(count):
        game = _game_from_data(fname, kind, data, seed=seed)
        rec, log, _ = run_episode(game, seed, time_limit, view_shape, channels,
                                  action_seed=len(recs) + 1000 * seed0)
        recs.append(rec)
        logs.append(log)
    save_group(name, recs, logs, time_limit, view_shape, channels)


def make_ops():
    rs = np.random.default_rng(20240)
    out = {}
    # advance_board on random bit patterns, with draws
    boards, results, rngs0, rngs1, probs, nsteps = [], [], [], [], [], []
    for it in range(120):
        H, W = (int(v) for v in rs.integers(3, 34, 2))
        if it < 8:
            H, W = [(3, 3), (3, 7), (7, 3), (25, 25), (26, 26), (32, 32), (33, 31), (50, 50)][it]
        dens = [1.0, 0.3, 0.1][it % 3]
        b = (rs.integers(0, 65536, (H, W)) * (rs.random((H, W)) < dens)).astype(np.uint16)
        if it % 4 == 3:
            b &= rs.integers(0, 65536, (H, W)).astype(np.uint16)
        p = float(np.float32(rs.random()))
        n = int(rs.integers(1, 4))
        g = np.random.default_rng(it)
        r0 = orc.rng_from_generator(g)
        with set_rng(g):
            res = speedups.advance_board(b, p, n)
        boards.append(b.ravel()); results.append(res.ravel())
        rngs0.append(r0); rngs1.append(orc.rng_from_generator(g))
        probs.append(p); nsteps.append((H, W, n))
    out["adv_in"] = np.concatenate(boards)
    out["adv_out"] = np.concatenate(results)
    out["adv_rng0"] = np.stack(rngs0); out["adv_rng1"] = np.stack(rngs1)
    out["adv_prob"] = np.array(probs, dtype=np.float32)
    out["adv_shape_n"] = np.array(nsteps, dtype=np.int32)
    # alive_counts
    ac_b, ac_g, ac_o, ac_shape = [], [], [], []
    for it in range(60):
        H, W = (int(v) for v in rs.integers(3, 34, 2))
        b = rs.integers(0, 65536, (H, W)).astype(np.uint16)
        g = rs.integers(0, 65536, (H, W)).astype(np.uint16)
        ac_b.append(b.ravel()); ac_g.append(g.ravel())
        ac_o.append(speedups.alive_counts(b, g)); ac_shape.append((H, W))
    out["ac_board"] = np.concatenate(ac_b); out["ac_goals"] = np.concatenate(ac_g)
    out["ac_out"] = np.stack(ac_o); out["ac_shape"] = np.array(ac_shape, dtype=np.int32)
    # execute_actions, incl. multi-agent and broadcast actions
    kinds = np.array([0, 0, 0, 0, 9, 1, 16, 32788, 17, 32789, 152, 144, 272, 784, 48, 53, 85,
                      32884, 122, 122 | 256, 126, 122 | 0x8000, 4, 0x8000, 0x8004, 8, 256],
                     dtype=np.uint16)
    ea = {k: [] for k in ("in", "out", "loc_in", "loc_out", "act", "meta")}
    for it in range(1500):
        H, W = (int(v) for v in rs.integers(3, 8, 2))
        b = kinds[rs.integers(0, len(kinds), (H, W))]
        if it % 4 == 0:
            b = b | (rs.integers(0, 8, (H, W)) << 9).astype(np.uint16)
        na = int(rs.integers(1, 4))
        locs = np.stack([rs.integers(0, H, na), rs.integers(0, W, na)], axis=1).astype(np.int64)
        for k in range(na):
            if rs.random() < 0.85:
                b[locs[k, 0], locs[k, 1]] = (
                    122 | (256 if rs.random() < 0.5 else 0) | (int(rs.integers(0, 8)) << 9)
                    | (int(rs.integers(0, 4)) << 12) | (4 if rs.random() < 0.1 else 0)
                    | (0x8000 if rs.random() < 0.1 else 0))
        acts = rs.integers(0, 9, na if rs.random() < 0.8 else 1).astype(np.int64)
        b1, l1 = b.copy(), locs.copy()
        speedups.execute_actions(b1, l1, acts)
        ea["in"].append(b.ravel()); ea["out"].append(b1.ravel())
        ea["loc_in"].append(locs.ravel()); ea["loc_out"].append(l1.ravel())
        ea["act"].append(acts)
        ea["meta"].append((H, W, na, len(acts)))
    for k in ("in", "out", "loc_in", "loc_out", "act"):
        out["ea_" + k] = np.concatenate(ea[k])
    out["ea_meta"] = np.array(ea["meta"], dtype=np.int32)
    # life_occupancy
    oc = {k: [] for k in ("in", "out", "meta", "rng0", "rng1")}
    for it in range(12):
        H, W = (int(v) for v in rs.integers(3, 30, 2))
        b = (rs.integers(0, 65536, (H, W)) & rs.integers(0, 65536, (H, W))).astype(np.uint16)
        n = int(rs.integers(1, 40))
        g = np.random.default_rng(100 + it)
        oc["rng0"].append(orc.rng_from_generator(g))
        with set_rng(g):
            res = speedups.life_occupancy(b, 0.3, n)
        oc["rng1"].append(orc.rng_from_generator(g))
        oc["in"].append(b.ravel()); oc["out"].append(res.ravel()); oc["meta"].append((H, W, n))
    out["occ_in"] = np.concatenate(oc["in"]); out["occ_out"] = np.concatenate(oc["out"])
    out["occ_meta"] = np.array(oc["meta"], dtype=np.int32)
    out["occ_rng0"] = np.stack(oc["rng0"]); out["occ_rng1"] = np.stack(oc["rng1"])
    # wrapped_label
    wl = {k: [] for k in ("in", "out", "meta")}
    for it in range(150):
        H, W = (int(v) for v in rs.integers(3, 26, 2))
        d = ((rs.random((H, W)) < rs.random()) * rs.integers(-1, 5, (H, W))).astype(np.int32)
        lab, n = speedups.wrapped_label(d)
        wl["in"].append(d.ravel()); wl["out"].append(lab.ravel()); wl["meta"].append((H, W, n))
    out["wl_in"] = np.concatenate(wl["in"]); out["wl_out"] = np.concatenate(wl["out"])
    out["wl_meta"] = np.array(wl["meta"], dtype=np.int32)
    path = os.path.join(HERE, "ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_occupancy():
    """side_effects.py:101-111 at the end of a short random episode."""
    recs = []
    ss = np.random.SeedSequence(77)
    picks = [("benchmarks/v1.0/append-spawn.npz", 2), ("benchmarks/v1.0/prune-still.npz", 1),
             ("benchmarks/v1.0/navigation.npz", 2), ("benchmarks/v1.0/prune-spawn.npz", 1)]
    for path, limit in picks:
        for fname, data in static_games(path, limit):
            game = SafeLifeGame.loaddata(data, auto_cls=False)
            seed = ss.spawn(1)[0]
            rec, log, game = run_episode(game, seed, 60, (15, 15), None, action_seed=len(recs))
            g = np.random.default_rng(4242 + len(recs))
            r0 = orc.rng_from_generator(g)
            b0 = game._init_data["board"]
            with set_rng(g):
                b1 = speedups.advance_board(b0, game.spawn_prob, game.num_steps)
                c0 = speedups.life_occupancy(b1, game.spawn_prob, 200)
                c1 = speedups.life_occupancy(game.board, game.spawn_prob, 200)
            recs.append({
                "init_board": np.asarray(b0, dtype=np.uint16), "final_board": game.board.copy(),
                "num_steps": np.int32(game.num_steps), "spawn_prob": np.float32(game.spawn_prob),
                "rng0": r0, "rng1": orc.rng_from_generator(g),
                "inaction_board": b1, "counts": np.stack([c0, c1]),
            })
    out = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    out["num_samples"] = np.int32(200)
    path = os.path.join(HERE, "occupancy.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_occupancy_nav50():
    """The same on procgen 50x50 navigation levels (BASELINE config 5: the wide kernels)."""
    recs = []
    (fname, kind, data), = _load_files(["random/navigation"])
    data = dict(data, board_shape=[50, 50])
    for k, seed in enumerate(np.random.SeedSequence(55).spawn(2)):
        game = _game_from_data(fname, kind, data, seed=seed)
        rec, log, game = run_episode(game, seed, 40, (15, 15), None, action_seed=500 + k)
        g = np.random.default_rng(5050 + k)
        r0 = orc.rng_from_generator(g)
        b0 = game._init_data["board"]
        with set_rng(g):
            b1 = speedups.advance_board(b0, game.spawn_prob, game.num_steps)
            c0 = speedups.life_occupancy(b1, game.spawn_prob, 300)
            c1 = speedups.life_occupancy(game.board, game.spawn_prob, 300)
        recs.append({
            "init_board": np.asarray(b0, dtype=np.uint16), "final_board": game.board.copy(),
            "num_steps": np.int32(game.num_steps), "spawn_prob": np.float32(game.spawn_prob),
            "rng0": r0, "rng1": orc.rng_from_generator(g),
            "inaction_board": b1, "counts": np.stack([c0, c1]),
        })
    out = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    out["num_samples"] = np.int32(300)
    path = os.path.join(HERE, "occupancy_nav50.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_side_effects():
    """side_effects.side_effect_score end to end on the reference."""
    from safelife.side_effects import side_effect_score
    recs = []
    ss = np.random.SeedSequence(78)
    picks = [("benchmarks/v1.0/append-spawn.npz", 2), ("benchmarks/v1.0/prune-still.npz", 2),
             ("benchmarks/v1.0/navigation.npz", 2), ("benchmarks/v1.0/prune-spawn.npz", 1),
             ("benchmarks/v1.0/append-still.npz", 1)]
    for path, limit in picks:
        for fname, data in static_games(path, limit):
            game = SafeLifeGame.loaddata(data, auto_cls=False)
            seed = ss.spawn(1)[0]
            rec, log, game = run_episode(game, seed, 80, (15, 15), None, action_seed=100 + len(recs))
            g = np.random.default_rng(777 + len(recs))
            r0 = orc.rng_from_generator(g)
            with set_rng(g):
                scores = side_effect_score(game, num_samples=200)
            keys = sorted(scores)
            vals = np.array([[scores[k][0], scores[k][1]] for k in keys], dtype=np.float64)
            pad = 16 - len(keys)
            assert pad >= 0
            recs.append({
                "init_board": np.asarray(game._init_data["board"], dtype=np.uint16),
                "final_board": game.board.copy(), "num_steps": np.int32(game.num_steps),
                "spawn_prob": np.float32(game.spawn_prob), "rng0": r0,
                "rng1": orc.rng_from_generator(g), "n_keys": np.int32(len(keys)),
                "keys": np.array(keys + [0] * pad, dtype=np.int32),
                "scores": np.concatenate([vals.reshape(-1, 2), np.zeros((pad, 2))]),
            })
    out = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    out["num_samples"] = np.int32(200)
    path = os.path.join(HERE, "side_effects.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    print("keys per record", out["n_keys"], "max emd", out["scores"][..., 0].max())


def main():
    which = set(sys.argv[1:])

    def want(name):
        return not which or name in which

    if want("ops"):
        make_ops()
    if want("occupancy"):
        make_occupancy()
    if want("side_effects"):
        make_side_effects()
    if want("occupancy_nav50"):
        make_occupancy_nav50()
    v10 = ["benchmarks/v1.0/%s.npz" % n for n in (
        "append-still", "append-dynamic", "append-spawn", "prune-dynamic", "prune-spawn",
        "prune-still", "prune-still-hard", "navigation")]
    if want("v1.0"):
        group_static("v1.0", v10, 4, 150, (25, 25), TRAIN_CHANNELS, seed0=10)
    if want("v0.1"):
        group_static("v0.1", ["benchmarks/v0.1/*.npz"], None, 120, (15, 15),
                     SafeLifeEnv.output_channels, seed0=11)
    if want("puzzles"):
        group_static("puzzles", ["puzzles/*.npz"], None, 200, (15, 15), None, seed0=12)
    if want("mazes"):
        # exits open from the start (min_performance -1) so that random walks exit
        group_static("mazes", ["mazes/*.npz"], None, 1000, (9, 31), TRAIN_CHANNELS, seed0=13,
                     min_performance=-1)
    if want("navtrain"):
        group_static("navtrain", ["training/navigation.npz"], 8, 150, (25, 25),
                     TRAIN_CHANNELS, seed0=14)
    if want("wrapped"):
        group_wrapped("v1.0", v10[:6], 3, 120, seed0=21, se_penalty=0.7, exit_difficulty=0.5)
        # the "inaction" baseline (a second stencil on a shadow board per step,
        # env_wrappers.py:179-180) on levels without spawners: the reference draws the
        # shadow board's spawns from the process-global generator, which has no batched
        # counterpart, so only deterministic levels can be pinned -- the -dynamic tasks
        # (oscillators: the shadow board really moves) and one still task
        quiet = ["benchmarks/v1.0/%s.npz" % n for n in ("append-dynamic", "prune-dynamic", "prune-still")]
        group_wrapped("inaction", quiet, 3, 120, seed0=23, se_penalty=0.5, exit_difficulty=0.7,
                      baseline="inaction")
        # ignore_reward_cells=True (env_wrappers.py:197-205), with spawners on the real board
        group_wrapped("ignore", [v10[0], v10[2], v10[4]], 3, 120, seed0=24, se_penalty=0.3,
                      exit_difficulty=1.0, ignore_reward_cells=True)
        group_wrapped("mazes", ["mazes/*.npz"], None, 600, seed0=22, se_penalty=0.25,
                      exit_difficulty=1.0, min_performance=0.0)
    if want("procgen25"):
        # BASELINE config 1: 25x25 append-still-easy, SeedSequence(0).spawn(1)[0], 1000 steps
        group_procgen("procgen25", "random/append-still-easy", (25, 25), 1, 1000, (25, 25),
                      TRAIN_CHANNELS, seed0=0)
    if want("procgen25spawn"):
        group_procgen("procgen25spawn", "random/append-spawn", (25, 25), 6, 200, (25, 25),
                      TRAIN_CHANNELS, seed0=4)
    # the other BASELINE pools, straight through the reference env: config 3 (prune-still
    # 25x25), config 5 (navigation 50x50: dynamic goals, spawners, the wide kernels) and
    # the reference-native 26x26 append-still
    if want("procgen25prune"):
        group_procgen("procgen25prune", "random/prune-still", (25, 25), 6, 250, (25, 25),
                      TRAIN_CHANNELS, seed0=3)
    if want("procgen50nav"):
        group_procgen("procgen50nav", "random/navigation", (50, 50), 3, 150, (25, 25),
                      TRAIN_CHANNELS, seed0=5)
    if want("procgen26"):
        group_procgen("procgen26", "random/append-still", (26, 26), 4, 300, (15, 15),
                      SafeLifeEnv.output_channels, seed0=1)


if __name__ == "__main__":
    main()
