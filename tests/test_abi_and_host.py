"""CPU-side checks: the C-ABI library builds, loads and exports every declared
symbol; host logic (level constants, spawn threshold, kwargs) is right."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc
from safelife_b200 import _lib, levels as lv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "safelife_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(handle, name), name
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)
    assert handle.sl_abi_version() == 3


def test_struct_layout_matches_header():
    # 8 int32 + 16 pointers; SlEnvState is 96 bytes
    assert ctypes.sizeof(_lib.SlEnvBatch) == 32 + 16 * 8
    assert _lib.SlEnvBatch.board.offset == 32
    dt = _lib.ENV_STATE_DTYPE
    assert dt.itemsize == 96 and dt.fields["loc"][1] == 32 and dt.fields["spawn_thr"][1] == 64
    assert dt.fields["count_points"][1] == 80 and dt.fields["ep_reward"][1] == 56
    # SlHostSession: five pointers, then idle_timeout_ms, seq, n_envs, reserved
    assert ctypes.sizeof(_lib.SlHostSession) == 5 * 8 + 4 * 4
    assert _lib.SlHostSession.host_out.offset == 32 and _lib.SlHostSession.seq.offset == 44


def test_argument_errors_without_gpu():
    lib = _lib.lib()
    b = np.zeros((2, 2), dtype=np.uint16)
    rc = lib.sl_advance_board_host(b.ctypes.data, b.ctypes.data, 1, 2, 2, 0.3, 1, None)
    assert rc == _lib.SL_ERR_ARG
    assert b"3x3" in lib.sl_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    b = np.zeros((3, 3), dtype=np.uint16)
    assert lib.sl_advance_board_host(b.ctypes.data, b.ctypes.data, 1, 3, 3, 0.3, 0, None) == _lib.SL_ERR_ARG


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from safelife_b200 import speedups
    with pytest.raises(RuntimeError):
        speedups.advance_board(np.zeros((5, 5), dtype=np.uint16))
    from safelife_b200.env import BatchedSafeLifeEnv
    level = {"board": np.zeros((5, 5), np.uint16), "goals": np.zeros((5, 5), np.uint16),
             "agent_loc": None, "points_table": lv.DEFAULT_POINTS_TABLE,
             "min_performance": -1, "spawn_prob": 0.3}
    with pytest.raises(RuntimeError):
        BatchedSafeLifeEnv([level], 1)
    with pytest.raises(ValueError):
        BatchedSafeLifeEnv([level], 1, no_such_option=1)


def test_product_tree_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "safelife_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), os.path.join(dirpath, f)


def test_spawn_threshold_is_the_float_double_compare():
    rs = np.random.default_rng(0)
    for p in list(rs.random(200)) + [0.0, 1.0, 0.3, 1e-9, -0.5, 2.0, 0.5]:
        thr = lv.spawn_threshold(p)
        pf = float(np.float32(p))
        u = rs.integers(0, 1 << 53, 2000, dtype=np.uint64)
        # values right around the threshold
        around = np.clip(np.array([thr - 2, thr - 1, thr, thr + 1], dtype=np.int64), 0, (1 << 53) - 1)
        u = np.concatenate([u, around.astype(np.uint64)])
        draws = u.astype(np.float64) * (1.0 / 9007199254740992.0)
        assert np.array_equal(draws < pf, u < np.uint64(thr)), p


@pytest.mark.parametrize("group", gu.traj_groups())
def test_level_pool_constants_match_oracle(group):
    d = gu.load_traj(group)
    levels = gu.levels_of(d)
    pool = lv.LevelPool(levels)
    for i, level in enumerate(levels):
        exits, init_points, required = orc.level_constants(
            level["board"], level["goals"], level["points_table"], level["min_performance"])
        assert np.array_equal(pool.exits[i][pool.exits[i] >= 0], exits)
        assert pool.init_points[i] == init_points
        assert pool.required[i] == required
        assert np.array_equal(lv.alive_counts_np(level["board"], level["goals"]),
                              orc.alive_counts(level["board"], level["goals"]))


def test_pool_roundtrip(tmp_path):
    levels = gu.levels_of(gu.load_traj("mazes"))
    path = str(tmp_path / "pool.npz")
    lv.save_pool(path, levels)
    back = lv.load_levels(path)
    assert len(back) == len(levels)
    for a, b in zip(levels, back):
        assert np.array_equal(a["board"], b["board"]) and np.array_equal(a["goals"], b["goals"])
        assert a["agent_loc"] == b["agent_loc"]
        assert np.array_equal(np.asarray(a["points_table"]), b["points_table"])
        assert a["min_performance"] == b["min_performance"]


def test_normalize_level_old_and_new_keys():
    board = np.zeros((6, 7), dtype=np.uint16)
    board[2, 5] = 122 | (3 << 12)
    old = {"board": board, "goals": np.zeros_like(board), "agent_loc": np.array([5, 2]),
           "orientation": np.array(1)}
    n = lv.normalize_level(old)
    assert n["agent_loc"] == (2, 5)
    assert n["board"][2, 5] == 122 | (1 << 12)
    new = {"board": board, "goals": np.zeros_like(board), "agent_locs": np.array([[2, 5]]),
           "min_performance": np.float64(0.5), "points_table": lv.DEFAULT_POINTS_TABLE[None]}
    n = lv.normalize_level(new)
    assert n["agent_loc"] == (2, 5) and n["min_performance"] == 0.5
    with pytest.raises(ValueError):
        lv.normalize_level({"board": board, "goals": board, "agent_locs": np.zeros((2, 2), int)})


def test_host_session_mailbox_protocol_host_side():
    """
    sl_host_session_step is pure host code: it tags the actions into the mailbox and
    waits for the done word.  A Python thread plays the device side of the protocol
    (csrc/session.cuh: every byte of the action buffer must carry the low nibble of
    the step's sequence number before the step may run; the done word carries the
    sequence number; a dead session is reported as SL_SESSION_ENDED with nothing
    stepped).
    """
    import ctypes
    import threading
    import time
    lib = _lib.lib()
    n = 37                                   # not a multiple of 16: the padding is tagged too
    n_pad = (n + 15) & ~15
    # (numpy buffers are 16-byte aligned for these sizes; the library stores 16 bytes at a time)
    raw = np.zeros(n_pad + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    host_actions = raw[off:off + n_pad]
    host_ctrl = np.zeros(4, dtype=np.uint32)
    host_out = np.zeros(n, dtype=np.uint64)
    s = _lib.SlHostSession(host_actions=host_actions.ctypes.data, dev_actions=1, dev_ctrl=1,
                           host_ctrl=host_ctrl.ctypes.data, host_out=host_out.ctypes.data,
                           idle_timeout_ms=50, seq=0, n_envs=n)
    host_ctrl[1] = 1                         # alive
    seen, stop = [], threading.Event()

    def device():
        seq = 1
        while not stop.is_set():
            tag = (seq & 15) << 4
            if np.all((host_actions & 0xF0) == tag):
                seen.append((host_actions[:n] & 0x0F).copy())
                assert np.all(host_actions[n:] == tag)
                host_out[:] = seq                  # "result records"
                host_ctrl[0] = seq                 # done word
                seq += 1
            else:
                time.sleep(0)
    th = threading.Thread(target=device, daemon=True)
    th.start()
    rs = np.random.default_rng(0)
    timing = np.zeros(2)
    for step in range(1, 40):                # crosses the 4-bit tag's wrap-around twice
        a = rs.integers(0, 9, n).astype(np.uint8)
        rc = lib.sl_host_session_step(ctypes.byref(s), a.ctypes.data, timing.ctypes.data)
        assert rc == _lib.SL_OK and s.seq == step
        assert np.array_equal(seen[-1], a) and int(host_out[0]) == step
        assert timing[1] >= 0.0
    stop.set()
    th.join()
    # a session that ended (idle time-out on the device) steps nothing and says so
    host_ctrl[1] = 0
    a = rs.integers(0, 9, n).astype(np.uint8)
    assert lib.sl_host_session_step(ctypes.byref(s), a.ctypes.data, None) == _lib.SL_SESSION_ENDED
    assert s.seq == 39 and len(seen) == 39


def test_slab_offset_copy_arithmetic():
    """
    The arithmetic of the single-pass kernel's slab-offset copies (csrc/warp_kernels.cuh,
    k_env_step_warp; csrc/session.cuh), restated: for any number of envs per CTA the bulk
    load covers the CTA's slabs with whole 16-byte units inside the array, and the bulk
    store plus the two cell-wise edges write exactly the CTA's own bytes, once.
    """
    rs = np.random.default_rng(0)
    for _ in range(2000):
        H, W = int(rs.integers(3, 33)), int(rs.integers(3, 33))
        cells, epc, n_envs = H * W, int(rs.integers(1, 9)), int(rs.integers(1, 5000))
        total = n_envs * cells * 2
        written = np.zeros(total, dtype=np.uint8)
        for cta in range((n_envs + epc - 1) // epc):
            e0 = cta * epc
            nenv = min(epc, n_envs - e0)
            group_bytes, byte0 = nenv * cells * 2, e0 * cells * 2
            mis = byte0 & 15
            load_bytes = (mis + group_bytes + 15) & ~15
            bulk = group_bytes >= 64 and byte0 - mis + load_bytes <= (total & ~15)
            if not bulk:                         # scalar copies of exactly the CTA's bytes
                written[byte0:byte0 + group_bytes] += 1
                continue
            assert (byte0 - mis) % 16 == 0 and load_bytes % 16 == 0
            assert byte0 - mis + load_bytes >= byte0 + group_bytes        # the load covers the slabs
            head = (16 - mis) & 15
            body = (group_bytes - head) & ~15
            tail = group_bytes - head - body
            assert (byte0 + head) % 16 == 0 and body % 16 == 0 and 0 <= tail < 16 and head % 2 == 0
            assert head // 2 <= 8 and tail // 2 <= 8                      # eight lanes per edge
            written[byte0:byte0 + head] += 1                              # lanes 0-7, cell by cell
            written[byte0 + head:byte0 + head + body] += 1                # bulk store
            written[byte0 + head + body:byte0 + group_bytes] += 1         # lanes 8-15
        assert (written == 1).all()
