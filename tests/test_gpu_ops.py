"""
GPU parity of the board operators, called through the C ABI
(safelife_b200.speedups -> sl_*_host; device-pointer entry points via torch),
against the golden vectors and the oracle.  Bit-exact.
"""
import ctypes

import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    from safelife_b200 import speedups
    return speedups


@pytest.fixture(scope="module")
def ops():
    return gu.load_npz("ops.npz")


def _split(flat, sizes):
    out, pos = [], 0
    for s in sizes:
        out.append(flat[pos:pos + s])
        pos += s
    return out


def _bitgen(words):
    g = np.random.PCG64(0)
    st = g.state
    st["state"]["state"] = (int(words[0]) << 64) | int(words[1])
    st["state"]["inc"] = (int(words[2]) << 64) | int(words[3])
    g.state = st
    return g


def test_advance_board_golden(sp, ops):
    meta = ops["adv_shape_n"]
    ins = _split(ops["adv_in"], [h * w for h, w, _ in meta])
    outs = _split(ops["adv_out"], [h * w for h, w, _ in meta])
    for i, (h, w, n) in enumerate(meta):
        bg = _bitgen(ops["adv_rng0"][i])
        sp.set_bit_generator(bg)
        res = sp.advance_board(ins[i].reshape(h, w), float(ops["adv_prob"][i]), int(n))
        assert np.array_equal(res.ravel(), outs[i]), (i, h, w, n)
        assert np.array_equal(orc.rng_from_generator(bg), ops["adv_rng1"][i]), i


def test_alive_counts_golden(sp, ops):
    sizes = [h * w for h, w in ops["ac_shape"]]
    bs, gs = _split(ops["ac_board"], sizes), _split(ops["ac_goals"], sizes)
    for i, (h, w) in enumerate(ops["ac_shape"]):
        assert np.array_equal(sp.alive_counts(bs[i].reshape(h, w), gs[i].reshape(h, w)),
                              ops["ac_out"][i])


def test_execute_actions_golden(sp, ops):
    meta = ops["ea_meta"]
    ins = _split(ops["ea_in"], [h * w for h, w, _, _ in meta])
    outs = _split(ops["ea_out"], [h * w for h, w, _, _ in meta])
    li = _split(ops["ea_loc_in"], [2 * na for _, _, na, _ in meta])
    lo = _split(ops["ea_loc_out"], [2 * na for _, _, na, _ in meta])
    acts = _split(ops["ea_act"], [k for _, _, _, k in meta])
    for i, (h, w, na, _) in enumerate(meta[:400]):
        b = ins[i].reshape(h, w).copy()
        loc = li[i].reshape(na, 2).copy()
        assert sp.execute_actions(b, loc, acts[i]) is None
        assert np.array_equal(b.ravel(), outs[i]), i
        assert np.array_equal(loc.ravel(), lo[i]), i


def test_execute_actions_negative_stride_view_is_written_back(sp):
    # the reference relies on WRITEBACKIFCOPY for the old-format agent_loc view
    # (safelife_game.py:219-221); a non-contiguous locations array must still be updated
    b = np.zeros((5, 5), dtype=np.uint16)
    b[2, 2] = 122
    locs = np.array([2, 2], dtype=np.int64)[None, ::-1]
    sp.execute_actions(b, locs, 2)   # move right
    assert locs.tolist() == [[2, 3]] and b[2, 3] & 2 and b[2, 2] == 0


def test_life_occupancy_golden(sp, ops):
    meta = ops["occ_meta"]
    ins = _split(ops["occ_in"], [h * w for h, w, _ in meta])
    outs = _split(ops["occ_out"], [h * w * 8 for h, w, _ in meta])
    for i, (h, w, n) in enumerate(meta):
        bg = _bitgen(ops["occ_rng0"][i])
        sp.set_bit_generator(bg)
        res = sp.life_occupancy(ins[i].reshape(h, w), 0.3, int(n))
        assert res.dtype == np.int32 and res.shape == (h, w, 8)
        assert np.array_equal(res.ravel(), outs[i]), i
        assert np.array_equal(orc.rng_from_generator(bg), ops["occ_rng1"][i]), i


def test_wrapped_label_golden(sp, ops):
    meta = ops["wl_meta"]
    ins = _split(ops["wl_in"], [h * w for h, w, _ in meta])
    outs = _split(ops["wl_out"], [h * w for h, w, _ in meta])
    for i, (h, w, n) in enumerate(meta):
        lab, num = sp.wrapped_label(ins[i].reshape(h, w))
        assert num == n and np.array_equal(lab.ravel(), outs[i]), i


@pytest.mark.parametrize("fixture", ["occupancy.npz", "occupancy_nav50.npz"])
def test_side_effect_occupancy_golden(sp, fixture):
    d = gu.load_npz(fixture)
    ns = int(d["num_samples"])
    for i in range(len(d["init_board"])):
        bg = _bitgen(d["rng0"][i])
        sp.set_bit_generator(bg)
        p = float(d["spawn_prob"][i])
        b1 = sp.advance_board(d["init_board"][i], p, int(d["num_steps"][i]))
        assert np.array_equal(b1, d["inaction_board"][i])
        c0 = sp.life_occupancy(b1, p, ns)
        c1 = sp.life_occupancy(d["final_board"][i], p, ns)
        assert np.array_equal(np.stack([c0, c1]), d["counts"][i])
        assert np.array_equal(orc.rng_from_generator(bg), d["rng1"][i])


@pytest.mark.parametrize("shape", [(3, 3), (3, 9), (9, 3), (25, 25), (26, 26), (31, 32), (32, 32),
                                   (33, 20), (50, 50), (64, 64), (70, 41)])
def test_advance_board_batch_random_vs_oracle(sp, shape):
    H, W = shape
    rs = np.random.default_rng(H * 100 + W)
    n = 24
    boards = (rs.integers(0, 65536, (n, H, W)) * (rs.random((n, H, W)) < 0.5)).astype(np.uint16)
    boards[: n // 2] &= rs.integers(0, 65536, (n // 2, H, W)).astype(np.uint16)
    rng = np.stack([orc.rng_from_seed(i) for i in range(n)])
    for steps in (1, 3):
        want_rng = rng.copy()
        want = np.stack([orc.advance_board(boards[i], 0.35, steps, want_rng[i]) for i in range(n)])
        got_rng = rng.copy()
        got = sp.advance_board_batch(boards, 0.35, steps, got_rng)
        assert np.array_equal(got, want)
        assert np.array_equal(got_rng, want_rng)


def test_spawn_prob_extremes(sp):
    rs = np.random.default_rng(5)
    boards = (rs.integers(0, 65536, (8, 12, 13)) & rs.integers(0, 65536, (8, 12, 13))).astype(np.uint16)
    for p in (0.0, 1.0, 1e-7, 0.999999):
        rng = np.stack([orc.rng_from_seed(i) for i in range(8)])
        want_rng = rng.copy()
        want = np.stack([orc.advance_board(boards[i], p, 2, want_rng[i]) for i in range(8)])
        got = sp.advance_board_batch(boards, p, 2, rng)
        assert np.array_equal(got, want) and np.array_equal(rng, want_rng)


def test_device_pointer_entry_points():
    import torch
    from safelife_b200 import _lib
    lib = _lib.lib()
    rs = np.random.default_rng(3)
    n, H, W = 16, 25, 25
    boards = (rs.integers(0, 65536, (n, H, W)) & rs.integers(0, 65536, (n, H, W))).astype(np.uint16)
    goals = rs.integers(0, 65536, (n, H, W)).astype(np.uint16)
    rng = np.stack([orc.rng_from_seed(100 + i) for i in range(n)])
    want_rng = rng.copy()
    want = np.stack([orc.advance_board(boards[i], 0.3, 2, want_rng[i]) for i in range(n)])
    d_b = torch.from_numpy(boards.view(np.int16)).cuda()
    d_o = torch.empty_like(d_b)
    d_r = torch.from_numpy(rng.view(np.int64)).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        _lib.check(lib.sl_advance_board(d_b.data_ptr(), d_o.data_ptr(), n, H, W, 0.3, 2,
                                        d_r.data_ptr(), ctypes.c_void_p(s.cuda_stream)))
    s.synchronize()
    assert np.array_equal(d_o.cpu().numpy().view(np.uint16), want)
    assert np.array_equal(d_r.cpu().numpy().view(np.uint64), want_rng)
    # alive_counts on device
    d_g = torch.from_numpy(goals.view(np.int16)).cuda()
    d_c = torch.zeros((n, 72), dtype=torch.int64, device="cuda")
    _lib.check(lib.sl_alive_counts(d_b.data_ptr(), d_g.data_ptr(), d_c.data_ptr(), n, H * W, None))
    torch.cuda.synchronize()
    want_c = np.stack([orc.alive_counts(boards[i], goals[i]) for i in range(n)])
    assert np.array_equal(d_c.cpu().numpy().reshape(n, 8, 9), want_c)


def test_error_conventions(sp):
    with pytest.raises(ValueError):
        sp.advance_board(np.zeros((2, 5), dtype=np.uint16))
    with pytest.raises(ValueError):
        sp.alive_counts(np.zeros((4, 4), np.uint16), np.zeros((4, 5), np.uint16))
    with pytest.raises(ValueError):
        sp.execute_actions(np.zeros((2, 2), np.uint16), np.zeros((1, 2), np.int64), 1)
    with pytest.raises(ValueError):
        sp.execute_actions(np.zeros((5, 5), np.uint16), np.zeros((3, 2), np.int64), [1, 2])
    with pytest.raises(ValueError):
        sp.advance_board(np.zeros((5, 5), dtype=np.uint16), 0.3, 0)


@pytest.mark.parametrize("shape,steps", [((25, 25), 1100), ((7, 9), 2050), ((32, 32), 300),
                                         ((40, 33), 60), ((50, 50), 300), ((64, 64), 270),
                                         ((20, 50), 40), ((70, 41), 30)])
def test_life_occupancy_long_runs_vs_oracle(sp, shape, steps):
    """
    More steps than an occupancy counter holds (10-bit sliced counters of the narrow
    kernel, byte counters of the wide one: flush paths) and a generic-size board.
    """
    H, W = shape
    rs = np.random.default_rng(steps)
    n = 6
    boards = (rs.integers(0, 65536, (n, H, W)) & rs.integers(0, 65536, (n, H, W))).astype(np.uint16)
    boards[0] = 0
    boards[0, 1, 1:4] = 9 | (3 << 9)          # a blinker that lives forever
    rng = np.stack([orc.rng_from_seed(50 + i) for i in range(n)])
    want_rng = rng.copy()
    want = np.stack([orc.life_occupancy(boards[i], 0.3, steps, want_rng[i]) for i in range(n)])
    got = sp.life_occupancy_batch(boards, 0.3, steps, rng)
    assert np.array_equal(got, want)
    assert np.array_equal(rng, want_rng)
    assert got[0, 1, 2, 3] == steps           # the blinker's centre is alive after every step


def test_advance_board_many_steps_vs_oracle(sp):
    rs = np.random.default_rng(77)
    boards = (rs.integers(0, 65536, (10, 26, 26)) & rs.integers(0, 65536, (10, 26, 26))).astype(np.uint16)
    rng = np.stack([orc.rng_from_seed(i) for i in range(10)])
    want_rng = rng.copy()
    want = np.stack([orc.advance_board(boards[i], 0.3, 150, want_rng[i]) for i in range(10)])
    got = sp.advance_board_batch(boards, 0.3, 150, rng)
    assert np.array_equal(got, want) and np.array_equal(rng, want_rng)
