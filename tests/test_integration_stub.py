"""
INTEGRATION.md section 1, executed: a CPython extension in the shape of the
reference's speedups_src/module.c (tests/integration/speedups_stub.c) is compiled
with gcc against libsafelife_b200.so and called like `safelife.speedups`.
"""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def _stub():
    sys.path.insert(0, os.path.join(HERE, "integration"))
    from integration import build_stub
    build_stub.build()
    import speedups_stub
    return speedups_stub


def test_stub_compiles_links_and_reports_errors_like_the_reference():
    stub = _stub()
    from safelife_b200 import _lib
    assert stub.abi_version == _lib.SL_ABI_VERSION
    with pytest.raises(ValueError, match="3x3"):
        stub.advance_board(np.zeros((2, 2), dtype=np.uint16))
    with pytest.raises(ValueError):
        stub.advance_board(np.zeros(9, dtype=np.uint16))
    with pytest.raises(ValueError, match="same size"):       # module.c:112-114
        stub.alive_counts(np.zeros((4, 4), dtype=np.uint16), np.zeros((4, 5), dtype=np.uint16))


@pytest.mark.gpu
def test_stub_matches_the_oracle_on_the_gpu():
    stub = _stub()
    rs = np.random.default_rng(5)
    for shape in ((25, 25), (26, 26), (50, 50), (9, 70)):
        board = rs.choice(np.array([0, 0, 0, 9, 9 | (2 << 9), 1, 16, 152, 144 | (1 << 9)], dtype=np.uint16), shape)
        goals = (rs.integers(0, 8, shape) << 9).astype(np.uint16)
        rng = orc.rng_from_seed(7)
        want_rng = rng.copy()
        want = orc.advance_board(board, 0.3, 5, want_rng)
        words = rng.copy()
        got = np.frombuffer(stub.advance_board(board, spawn_prob=0.3, n_step=5, rng=words),
                            dtype=np.uint16).reshape(shape)
        assert np.array_equal(got, want) and np.array_equal(words, want_rng), shape
        counts = np.frombuffer(stub.alive_counts(board, goals), dtype=np.int64).reshape(8, 9)
        assert np.array_equal(counts, orc.alive_counts(board, goals)), shape
