"""
side_effects.side_effect_score: oracle and CUDA path against
the reference's output (tests/golden/side_effects.npz, occupancy.npz) and
against each other on random boards.
"""
import ctypes

import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc

EMD_TOL = 1e-9   # double-precision solvers of the same LP; absolute, scores are O(1..100)


def _emd_host(a, b, dist, penalty):
    from safelife_b200 import _lib
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    out = ctypes.c_double()
    _lib.check(_lib.lib().sl_emd_host(a.ctypes.data, b.ctypes.data, dist.ctypes.data, len(a),
                                      float(penalty), ctypes.byref(out)))
    return out.value


@pytest.mark.parametrize("k", [1, 2, 3, 7, 16, 40])
def test_emd_host_vs_lp(k):
    """The library's exact transport solver against the oracle's LP (no GPU involved)."""
    rs = np.random.default_rng(k)
    for trial in range(12):
        pts = rs.integers(0, 26, (k, 2))
        d = pts[:, None, :] - pts[None, :, :]
        if trial % 2 == 0:
            d = np.abs(d)        # symmetric torus metric; odd trials keep the reference's
        d = np.abs(np.minimum(d, 26 - d)).sum(-1)   # signed wrap (side_effects.py:45-50): asymmetric
        dist = np.tanh(d / 5.0)
        a = rs.random(k) * (rs.random(k) < 0.7)
        b = rs.random(k) * (rs.random(k) < 0.7)
        if trial % 3 == 0:
            b *= a.sum() / max(b.sum(), 1e-9)          # balanced
        if trial % 4 == 1:
            a = np.round(a * 4) / 4                      # ties / degenerate supplies
            b = np.round(b * 4) / 4
        for penalty in (1.0, -1.0, 0.0):
            want = orc.emd_lp(a, b, dist, penalty)
            got = _emd_host(a, b, dist, penalty)
            assert abs(got - want) <= EMD_TOL * max(1.0, abs(want)), (k, trial, penalty, got, want)


def test_exact_solver_within_fastemd_quantisation_of_pyemd():
    """
    pyemd (not installable here) rounds masses and costs to 1e-6 of their maxima
    before solving (oracle.emd_pyemd_quantised restates that conversion).  The
    product's exact solver therefore differs from it by at most the rounding of k
    masses and of the costs on the moved mass: (k + 1) * 1e-6 * max_sum * max_cost.
    Measured here on problems shaped like side_effect_score's (tanh-of-torus-
    distance costs, densities in [0, 1]): far below 1e-5 relative.
    """
    rs = np.random.default_rng(17)
    worst = 0.0
    for trial in range(12):
        k = int(rs.integers(2, 30))
        ys, xs = rs.integers(0, 26, k), rs.integers(0, 26, k)
        dx = np.abs(xs[:, None] - xs[None, :])
        dy = np.abs(ys[:, None] - ys[None, :])
        dist = np.tanh((np.minimum(dx, 26 - dx) + np.minimum(dy, 26 - dy)) / 5.0)
        a = rs.random(k) * (rs.random(k) < 0.7)
        b = rs.random(k) * (rs.random(k) < 0.7)
        if a.sum() == 0 or b.sum() == 0 or dist.max() == 0:
            continue
        exact = _emd_host(a, b, dist, 1.0)
        quantised = orc.emd_pyemd_quantised(a, b, dist, 1.0)
        bound = (k + 1) * 1e-6 * max(a.sum(), b.sum()) * dist.max()
        assert abs(exact - quantised) <= bound, (trial, k, exact, quantised, bound)
        worst = max(worst, abs(exact - quantised) / max(exact, 1e-12))
    assert worst < 1e-5, worst


def test_emd_host_batch_threads_and_larger_problems():
    """sl_emd_host_batch: same values as one-by-one calls and as the LP, any thread count,
    including problems where most mass is common to both histograms (pre-flowed)."""
    from safelife_b200 import side_effects as se
    rs = np.random.default_rng(99)
    problems = [None]
    for k in (5, 60, 150):
        a = np.zeros((26, 26))
        cells = rs.choice(26 * 26, k, replace=False)
        a.flat[cells] = rs.random(k)
        b = a.copy()
        moved = rs.choice(cells, k // 2, replace=False)
        b.flat[moved] *= rs.random(k // 2)                       # common mass + a deficit
        b.flat[rs.choice(26 * 26, k // 3, replace=False)] += 0.2 * rs.random(k // 3)
        problems.append(se._emd_problem(a, b))
    problems.append(None)
    want = [0.0 if p is None else orc.emd_lp(p[0], p[1], p[2], 1.0) for p in problems]
    for threads in (1, 3, 0):
        got = se.solve_emd_problems(problems, 1.0, threads)
        assert got.shape == (len(problems),)
        for g, w in zip(got, want):
            assert abs(g - w) <= EMD_TOL * max(1.0, abs(w)), (threads, g, w)
    for p, w in zip(problems, want):
        if p is not None:
            assert abs(_emd_host(p[0], p[1], p[2], 1.0) - w) <= EMD_TOL * max(1.0, abs(w))


def test_emd_host_argument_errors():
    from safelife_b200 import _lib
    out = ctypes.c_double()
    a = np.array([1.0, -1.0])
    assert _lib.lib().sl_emd_host(a.ctypes.data, a.ctypes.data, np.zeros(4).ctypes.data, 2, 1.0,
                                  ctypes.byref(out)) == _lib.SL_ERR_ARG
    assert _lib.lib().sl_emd_host(None, None, None, 0, 1.0, ctypes.byref(out)) == _lib.SL_OK
    assert out.value == 0.0


def _golden_scores(d, i):
    n = int(d["n_keys"][i])
    return {int(k): list(v) for k, v in zip(d["keys"][i][:n], d["scores"][i][:n])}


def test_oracle_side_effect_score_golden():
    d = gu.load_npz("side_effects.npz")
    ns = int(d["num_samples"])
    for i in range(len(d["init_board"])):
        rng = d["rng0"][i].copy()
        got, _ = orc.side_effect_score(d["init_board"][i], d["final_board"][i], int(d["num_steps"][i]),
                                       float(d["spawn_prob"][i]), rng, ns)
        want = _golden_scores(d, i)
        assert set(got) == set(want), i
        for k in want:
            assert abs(got[k][0] - want[k][0]) <= EMD_TOL * max(1.0, want[k][0]), (i, k)
            assert got[k][1] == want[k][1], (i, k)
        assert np.array_equal(rng, d["rng1"][i])


def test_host_scores_from_reference_counts():
    """
    The product's host half (key set, distributions, EMD through sl_emd_host) on the
    reference's own occupancy histograms: no GPU involved.
    """
    from safelife_b200 import side_effects as se
    d = gu.load_npz("occupancy.npz")
    ns = int(d["num_samples"])
    for i in range(len(d["init_board"])):
        rng = d["rng0"][i].copy()
        want, counts = orc.side_effect_score(d["init_board"][i], d["final_board"][i],
                                             int(d["num_steps"][i]), float(d["spawn_prob"][i]), rng, ns)
        assert np.array_equal(counts, d["counts"][i])
        got = se.scores_from_counts(d["counts"][i], d["init_board"][i], d["final_board"][i], ns)
        assert set(got) == set(want), i
        for k in want:
            assert abs(got[k][0] - want[k][0]) <= EMD_TOL * max(1.0, want[k][0]), (i, k)
            assert got[k][1] == want[k][1], (i, k)
        only = se.scores_from_counts(d["counts"][i], d["init_board"][i], d["final_board"][i], ns,
                                     exclude=[min(want)]) if want else {}
        assert set(only) == set(want) - {min(want)} if want else True


def test_earth_mover_distance_signature_and_edge_cases():
    """side_effects.earth_mover_distance (side_effects.py:13-57): defaults, no change, pure extra mass."""
    from safelife_b200 import side_effects as se
    a = np.zeros((6, 7))
    assert se.earth_mover_distance(a, a) == 0.0
    b = a.copy()
    b[2, 3] = 1.0
    assert se.earth_mover_distance(a, b) == pytest.approx(1.0)              # extra_mass_penalty * 1
    assert se.earth_mover_distance(a, b, extra_mass_penalty=0.25) == pytest.approx(0.25)
    c = a.copy()
    c[2, 5] = 1.0                                                             # two columns to the right
    assert se.earth_mover_distance(c, b) == pytest.approx(np.tanh(2 / 5.0))
    assert se.earth_mover_distance(c, b, tanh_scale=0) == pytest.approx(2.0)
    assert se.earth_mover_distance(c, b, metric="euclidean", tanh_scale=0) == pytest.approx(2.0)
    # the signed wrap of the reference: moving left across the seam is short, right is not
    e, f = a.copy(), a.copy()
    e[0, 0], f[0, 6] = 1.0, 1.0
    assert se.earth_mover_distance(f, e, tanh_scale=0) == pytest.approx(1.0)
    assert se.earth_mover_distance(e, f, tanh_scale=0) == pytest.approx(6.0)
    assert orc.earth_mover_distance(e, f, tanh_scale=1e9) == pytest.approx(np.tanh(6 / 1e9))


def test_cell_names():
    from safelife_b200 import side_effects as se
    assert se.cell_name(9 | (2 << 9)) == "life-green"
    assert se.cell_name(152 | (3 << 9)) == "spawner-yellow"
    assert se.name_to_cell("life-green") == 9 | (2 << 9)
    assert se.name_to_cell("spawner-yellow") == 152 | (3 << 9)
    assert se.cell_name(32788) == "crate-gray" and se.cell_name(122 | (1 << 12)) == "agent-gray"
    for c in (0, 9, 1, 16, 32788, 32789, 17, 32884, 85, 53, 152, 144, 272, 48):
        for col in range(8):
            assert se.name_to_cell(se.cell_name(c | (col << 9))) == c | (col << 9)


# ------------------------------------------------------------------- GPU
def _dev(a, dtype=None):
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:
        a = a.view(np.int16)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    return torch.from_numpy(a).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["occupancy.npz", "occupancy_nav50.npz"])
def test_side_effect_occupancy_golden_fused(fixture):
    """sl_side_effect_occupancy on the reference's occupancy fixtures (26x26 benchmark levels:
    the narrow kernel; 50x50 navigation: the wide one), all episodes in one launch."""
    from safelife_b200 import side_effects as se
    from safelife_b200.levels import spawn_threshold
    d = gu.load_npz(fixture)
    thr = np.array([spawn_threshold(float(p)) for p in d["spawn_prob"]], dtype=np.int64)
    rng = _dev(d["rng0"].copy())
    counts = se.occupancy_counts(_dev(d["init_board"]), _dev(d["final_board"]), _dev(d["num_steps"]),
                                 _dev(thr), rng, int(d["num_samples"]))
    assert np.array_equal(counts.cpu().numpy(), d["counts"])
    assert np.array_equal(rng.cpu().numpy().view(np.uint64), d["rng1"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(25, 25), (32, 32), (9, 14), (40, 35), (50, 50), (64, 64), (66, 30)])
def test_side_effect_occupancy_random_vs_oracle(shape):
    """Warp (H, W <= 32) and generic kernels against the oracle's three separate calls."""
    from safelife_b200 import side_effects as se
    from safelife_b200.levels import spawn_threshold
    H, W = shape
    rs = np.random.default_rng(H * 131 + W)
    n, ns = 10, 70
    init = (rs.integers(0, 65536, (n, H, W)) * (rs.random((n, H, W)) < 0.35)).astype(np.uint16)
    init[: n // 2] &= np.uint16(0x0E89)                  # plain life + spawner-ish bits, all colours
    final = init.copy()
    final[rs.random((n, H, W)) < 0.05] = 9 | (2 << 9)
    steps = rs.integers(0, 40, n).astype(np.int32)
    steps[0] = 1
    probs = rs.choice([0.0, 0.1, 0.3, 1.0], n).astype(np.float32)
    rng0 = np.stack([orc.rng_from_seed(1000 + i) for i in range(n)])
    thr = np.array([spawn_threshold(float(p)) for p in probs], dtype=np.int64)
    rng = _dev(rng0.copy())
    got = se.occupancy_counts(_dev(init), _dev(final), _dev(steps), _dev(thr), rng, ns).cpu().numpy()
    for i in range(n):
        r = rng0[i].copy()
        b1 = init[i] if steps[i] == 0 else orc.advance_board(init[i], float(probs[i]), int(steps[i]), r)
        c0 = orc.life_occupancy(b1, float(probs[i]), ns, r)
        c1 = orc.life_occupancy(final[i], float(probs[i]), ns, r)
        assert np.array_equal(got[i, 0], c0), (shape, i)
        assert np.array_equal(got[i, 1], c1), (shape, i)
        assert np.array_equal(rng[i].cpu().numpy().view(np.uint64), r), (shape, i)


@pytest.mark.gpu
def test_side_effect_score_golden():
    """Whole score (device occupancy + host EMD) against the reference's output."""
    import torch
    from safelife_b200 import side_effects as se
    from safelife_b200.levels import spawn_threshold
    d = gu.load_npz("side_effects.npz")
    ns = int(d["num_samples"])
    thr = np.array([spawn_threshold(float(p)) for p in d["spawn_prob"]], dtype=np.int64)
    rng = _dev(d["rng0"].copy())
    counts = se.occupancy_counts(_dev(d["init_board"]), _dev(d["final_board"]), _dev(d["num_steps"]),
                                 _dev(thr), rng, ns).cpu().numpy()
    assert np.array_equal(rng.cpu().numpy().view(np.uint64), d["rng1"])
    for i in range(len(thr)):
        got = se.scores_from_counts(counts[i], d["init_board"][i], d["final_board"][i], ns)
        want = _golden_scores(d, i)
        assert set(got) == set(want), i
        for k in want:
            assert abs(got[k][0] - want[k][0]) <= EMD_TOL * max(1.0, want[k][0]), (i, k)
            assert got[k][1] == want[k][1], (i, k)
        named = se.scores_from_counts(counts[i], d["init_board"][i], d["final_board"][i], ns,
                                      strkeys=True, include=["life-green", "spawner-yellow"])
        assert set(named) <= {"life-green", "spawner-yellow"}
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_single_game_side_effect_score_matches_reference_call_order():
    """The drop-in single-game function draws from the global bit generator like the reference."""
    from safelife_b200 import side_effects as se, speedups as sp
    d = gu.load_npz("side_effects.npz")
    ns = int(d["num_samples"])
    for i in (0, 4):
        st = d["rng0"][i]
        bg = np.random.PCG64()
        state = bg.state
        state["state"]["state"] = (int(st[0]) << 64) | int(st[1])
        state["state"]["inc"] = (int(st[2]) << 64) | int(st[3])
        bg.state = state
        sp.set_bit_generator(bg)

        class Game(object):
            pass
        g = Game()
        g._init_data = {"board": d["init_board"][i]}
        g.board, g.num_steps, g.spawn_prob = d["final_board"][i], int(d["num_steps"][i]), float(d["spawn_prob"][i])
        got = se.side_effect_score(g, num_samples=ns)
        want = _golden_scores(d, i)
        assert set(got) == set(want)
        for k in want:
            assert abs(got[k][0] - want[k][0]) <= EMD_TOL * max(1.0, want[k][0])
        assert np.array_equal(orc.rng_from_generator(bg), d["rng1"][i])


@pytest.mark.gpu
def test_env_reports_side_effects_at_episode_end():
    """BatchedSafeLifeEnv(should_calculate_side_effects=True): safelife_env.py:176-187."""
    from safelife_b200.env import BatchedSafeLifeEnv
    d = gu.load_traj("procgen25spawn")
    levels = gu.levels_of(d)
    n = len(levels)
    weights = {"life-green": 1.0, "spawner-yellow": 2.0}
    env = BatchedSafeLifeEnv(levels, n, device="cuda:0", time_limit=12, auto_reset=False,
                             should_calculate_side_effects=True, side_effect_weights=weights,
                             side_effect_samples=60)
    env.set_rng_state(d["rng_init"])
    ora = orc.OracleEnv(levels, list(d["rng_init"]), time_limit=12)
    env.reset()
    seen = 0
    finished = np.zeros(n, dtype=bool)
    for t in range(12):
        a = d["action"][:, t]
        obs, reward, done, info = env.step(a)
        r0, d0 = ora.step(a.astype(np.int64))
        assert np.array_equal(done.cpu().numpy(), d0)
        se_info = info["episode"].get("side_effects", {})
        newly = d0 & ~finished          # scored once, when the episode ends
        finished |= d0
        assert set(se_info) == set(np.flatnonzero(newly).tolist())
        for e, scores in se_info.items():
            rng = ora.rng[e].copy()
            want, _ = orc.side_effect_score(levels[e]["board"], ora.board[e], int(ora.num_steps[e]),
                                            float(levels[e]["spawn_prob"]), rng, 60)
            total = np.zeros(2)
            for key, w in weights.items():
                total += w * np.array(scores.get(key, 0))
            assert np.allclose(scores["total"], total)
            from safelife_b200 import side_effects as se
            for k, v in want.items():
                assert abs(scores[se.cell_name(k)][0] - v[0]) <= EMD_TOL * max(1.0, v[0])
                assert scores[se.cell_name(k)][1] == v[1]
            seen += 1
    assert seen == n          # every env timed out at step 12 at the latest
