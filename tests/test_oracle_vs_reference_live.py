"""
Live differential tests: the oracle (oracle/sl_oracle.c) against the reference's
own C extension compiled from its sources into oracle/_ref (oracle/Makefile),
on seeded random inputs well beyond the committed fixtures -- every bit
pattern, odd shapes, multi-agent actions, long runs, the RNG stream.  Skipped
where the compiled reference is absent.  No GPU involved.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refload

pytestmark = pytest.mark.skipif(not refload.have_ref_so(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ref():
    return refload.load_ref_speedups()


def _boards(rs, H, W, style):
    b = rs.integers(0, 65536, (H, W)).astype(np.uint16)
    if style == 0:
        return b                                             # every bit pattern
    if style == 1:
        return (b * (rs.random((H, W)) < 0.3)).astype(np.uint16)   # sparse
    if style == 2:
        return b & np.uint16(0x0E99)                         # life-like: alive / destructible / frozen / spawning / colours
    return (b & rs.integers(0, 65536, (H, W)).astype(np.uint16)) * (rs.random((H, W)) < 0.6)


def _bitgen(seed):
    g = np.random.default_rng(seed)
    return g, orc.rng_from_generator(g)


@pytest.mark.parametrize("seed", range(6))
def test_advance_board_and_rng_stream(ref, seed):
    rs = np.random.default_rng(1000 + seed)
    for it in range(60):
        H, W = (int(v) for v in rs.integers(3, 41, 2))
        board = _boards(rs, H, W, it % 4).astype(np.uint16)
        p = float(np.float32(rs.random())) if it % 5 else [0.0, 1.0][it % 2]
        n = int(rs.integers(1, 5))
        g, words = _bitgen(seed * 100 + it)
        ref.set_bit_generator(g.bit_generator)
        want = ref.advance_board(board, p, n)
        got = orc.advance_board(board, p, n, words)
        assert np.array_equal(got, want), (seed, it, H, W, p, n)
        assert np.array_equal(words, orc.rng_from_generator(g)), (seed, it)      # same number of draws


@pytest.mark.parametrize("seed", range(3))
def test_life_occupancy(ref, seed):
    rs = np.random.default_rng(2000 + seed)
    for it in range(12):
        H, W = (int(v) for v in rs.integers(3, 34, 2))
        board = _boards(rs, H, W, 1 + it % 3).astype(np.uint16)
        p = float(np.float32(rs.random()))
        n = int(rs.integers(1, 80))
        g, words = _bitgen(seed * 100 + it)
        ref.set_bit_generator(g.bit_generator)
        want = ref.life_occupancy(board, p, n)
        got = orc.life_occupancy(board, p, n, words)
        assert np.array_equal(got, want), (seed, it)
        assert np.array_equal(words, orc.rng_from_generator(g)), (seed, it)


@pytest.mark.parametrize("seed", range(4))
def test_alive_counts(ref, seed):
    rs = np.random.default_rng(3000 + seed)
    for it in range(100):
        H, W = (int(v) for v in rs.integers(1, 30, 2))
        board, goals = _boards(rs, H, W, it % 4), _boards(rs, H, W, (it + 1) % 4)
        want = ref.alive_counts(board.astype(np.uint16), goals.astype(np.uint16))
        assert np.array_equal(orc.alive_counts(board, goals), want), (seed, it)


@pytest.mark.parametrize("seed", range(4))
def test_execute_actions_multi_agent(ref, seed):
    rs = np.random.default_rng(4000 + seed)
    agent = np.uint16(2 | 8 | 16 | 32 | 64)          # CellTypes.player = 122
    for it in range(400):
        H, W = (int(v) for v in rs.integers(3, 12, 2))
        board = _boards(rs, H, W, it % 4).astype(np.uint16)
        board[rs.random((H, W)) < 0.4] = 0
        n_agents = int(rs.integers(1, 4))
        locs = np.stack([rs.integers(0, H, n_agents), rs.integers(0, W, n_agents)], axis=1).astype(np.int64)
        for k, (y, x) in enumerate(locs):
            if rs.random() < 0.85:                    # sometimes the agent is missing (skip path)
                extra = rs.choice([0, 4, 0x8000, 256, 4 | 0x8000])    # pushable / pullable / on an open exit
                board[y, x] = agent | np.uint16(extra) | np.uint16(int(rs.integers(0, 8)) << 9)
        actions = rs.integers(0, 9, n_agents).astype(np.int64)
        if it % 7 == 0:
            actions = actions[:1]                     # one action broadcast to all agents
        b1, l1 = board.copy(), locs.copy()
        b2, l2 = board.copy(), locs.copy()
        ref.execute_actions(b1, l1, actions)
        orc.execute_actions(b2, l2, actions)
        assert np.array_equal(b1, b2), (seed, it)
        assert np.array_equal(l1, l2), (seed, it)


@pytest.mark.parametrize("seed", range(3))
def test_wrapped_label_numbering(ref, seed):
    rs = np.random.default_rng(5000 + seed)
    for it in range(80):
        H, W = (int(v) for v in rs.integers(1, 24, 2))
        data = (rs.random((H, W)) < [0.1, 0.35, 0.6, 0.9][it % 4]).astype(np.int32) * rs.integers(1, 9, (H, W)).astype(np.int32)
        want, n_want = ref.wrapped_label(data)
        got, n_got = orc.wrapped_label(data)
        assert n_got == n_want, (seed, it)
        assert np.array_equal(got, want), (seed, it)   # label ids follow the reference's visiting order
