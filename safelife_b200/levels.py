"""
Level data for the batched engine: loading SafeLife level files and deriving
the per-level constants the step kernels need.

Host-side numpy only.  Mirrors what the reference does at load / reset time:

* ``GameState.deserialize`` / ``GameWithGoals.deserialize``
  (safelife/safelife_game.py:211-234, 622-637): board copy, ``agent_loc``
  (old key, stored as (x, y)) or ``agent_locs`` (new key, (row, col)),
  optional ``orientation`` written into the agent cell, ``spawn_prob``,
  ``min_performance``, ``points_table``.
* ``update_exit_locs`` (:533-535), ``setup_initial_counts`` (:665-675),
  ``initial_available_points`` / ``required_points`` (:696-714).

Only single-agent levels (at most one agent) are supported by the env engine.
"""
import numpy as np

ALIVE, AGENT, PUSHABLE, DESTRUCTIBLE, FROZEN = 1, 2, 4, 8, 16
SPAWNING, EXIT, COLORS, ORIENT, PULLABLE = 128, 256, 0x0E00, 0x3000, 0x8000

DEFAULT_POINTS_TABLE = np.array([   # GameWithGoals.default_points_table, safelife_game.py:595-605
    # k   r   g   y   b   m   c   w  empty
    [+0, -1, +0, +0, +0, +0, +0, +0, 0],
    [-3, +3, -3, +0, -3, +0, -3, -3, 0],
    [+0, -3, +5, +0, +0, +0, +3, +0, 0],
    [-3, +0, +0, +3, +0, +0, +0, +0, 0],
    [+3, -3, +3, +0, +5, +3, +3, +3, 0],
    [-3, +3, -3, +0, -3, +5, -3, -3, 0],
    [+3, -3, +3, +0, +3, +0, +5, +3, 0],
    [+0, -1, +0, +0, +0, +0, +0, +0, 0],
], dtype=np.int64)


def spawn_threshold(spawn_prob):
    """
    Integer form of the reference's ``random_float() < spawn_prob``
    (advance_board.c:115): spawn_prob is parsed as a C float (module.c:23), the
    draw is (u64 >> 11) * 2^-53, so the draw succeeds iff
    (u64 >> 11) < ceil(float32(p) * 2^53).
    """
    t = float(np.float32(spawn_prob)) * 9007199254740992.0
    if not t > 0.0:
        return 0
    if t >= 9007199254740992.0:
        return 1 << 53
    return int(np.ceil(t))


def alive_counts_np(board, goals):
    """numpy form of alive_counts (advance_board.c:192-207) for load-time constants."""
    b = np.asarray(board, dtype=np.uint16).ravel().astype(np.int64)
    g = np.asarray(goals, dtype=np.uint16).ravel().astype(np.int64)
    counted = ((b & (DESTRUCTIBLE | PUSHABLE | PULLABLE)) != 0) | ((b & FROZEN) == 0)
    col = np.where(b & ALIVE, (b & COLORS) >> 9, 8)
    bins = ((g & COLORS) >> 9) * 9 + col
    return np.bincount(bins[counted], minlength=72).reshape(8, 9).astype(np.int64)


def normalize_level(data):
    """
    One level (npz mapping, structured-array record or dict) -> dict with
    board, goals, agent_loc ((row, col) or None), points_table int64[8,9],
    min_performance, spawn_prob -- the state right after GameState.deserialize,
    before update_exit_colors.
    """
    keys = data.dtype.fields if hasattr(data, "dtype") and data.dtype.fields else data
    keys = list(keys.keys()) if hasattr(keys, "keys") else list(keys)
    board = np.array(data["board"], dtype=np.uint16)
    goals = np.array(data["goals"], dtype=np.uint16) if "goals" in keys else np.zeros_like(board)
    if "agent_loc" in keys and data["agent_loc"] is not None and "agent_locs" not in keys:
        locs = np.asarray(data["agent_loc"]).ravel()[::-1].reshape(1, 2)   # stored (x, y), :219-221
    elif "agent_locs" in keys:
        locs = np.asarray(data["agent_locs"]).reshape(-1, 2)
    else:
        locs = np.zeros((0, 2), dtype=np.int64)
    if len(locs) > 1:
        raise ValueError("multi-agent levels are not supported by the batched env")
    agent_loc = None if len(locs) == 0 else (int(locs[0, 0]), int(locs[0, 1]))
    if "orientation" in keys and agent_loc is not None:   # :228-229, :322-332
        o = (int(np.asarray(data["orientation"]).ravel()[0]) & 3) << 12
        board[agent_loc] = (int(board[agent_loc]) & ~ORIENT & 0xFFFF) | o
    if "points_table" in keys:
        table = np.asarray(data["points_table"])
        if np.issubdtype(table.dtype, np.floating) and not np.all(table == np.round(table)):
            raise ValueError("non-integer points_table is not supported")
        table = table.astype(np.int64).reshape(-1, 8, 9)[0]
    else:
        table = DEFAULT_POINTS_TABLE
    return {
        "board": board, "goals": goals, "agent_loc": agent_loc, "points_table": table,
        "min_performance": float(data["min_performance"]) if "min_performance" in keys else -1.0,
        "spawn_prob": float(data["spawn_prob"]) if "spawn_prob" in keys else 0.3,
    }


def load_levels(path):
    """
    Load a SafeLife ``.npz``: a single level, a reference archive with a
    ``levels`` structured array (level_iterator.py:86-99), or a pool written by
    ``save_pool``.
    """
    with np.load(path, allow_pickle=False) as d:
        if "levels" in d.files:
            return [normalize_level(rec) for rec in d["levels"]]
        if "pool_board" in d.files:
            out = []
            for i in range(len(d["pool_board"])):
                loc = d["pool_agent_loc"][i]
                out.append({
                    "board": d["pool_board"][i], "goals": d["pool_goals"][i],
                    "agent_loc": None if loc[0] < 0 else (int(loc[0]), int(loc[1])),
                    "points_table": d["pool_points_table"][i].astype(np.int64),
                    "min_performance": float(d["pool_min_performance"][i]),
                    "spawn_prob": float(d["pool_spawn_prob"][i]),
                })
            return out
        return [normalize_level({k: d[k] for k in d.files})]


def save_pool(path, levels):
    """Write a list of normalised levels as one compressed archive."""
    np.savez_compressed(
        path,
        pool_board=np.stack([lv["board"] for lv in levels]).astype(np.uint16),
        pool_goals=np.stack([lv["goals"] for lv in levels]).astype(np.uint16),
        pool_agent_loc=np.array([(-1, -1) if lv["agent_loc"] is None else lv["agent_loc"]
                                 for lv in levels], dtype=np.int32),
        pool_points_table=np.stack([np.asarray(lv["points_table"]).reshape(8, 9)
                                    for lv in levels]).astype(np.int8),
        pool_min_performance=np.array([lv["min_performance"] for lv in levels], dtype=np.float64),
        pool_spawn_prob=np.array([lv["spawn_prob"] for lv in levels], dtype=np.float64),
    )


class LevelPool(object):
    """
    Stacked level arrays plus per-level constants, as numpy arrays with the
    dtypes of struct SlEnvBatch's pool_* members (include/safelife_b200.h).
    """

    def __init__(self, levels, min_performance_fraction=1.0):
        if not levels:
            raise ValueError("empty level pool")
        shape = levels[0]["board"].shape
        for lv in levels:
            if lv["board"].shape != shape or lv["goals"].shape != shape:
                raise ValueError("all levels of a pool must share one board shape")
        if shape[0] < 3 or shape[1] < 3:
            raise ValueError("Board must be at least 3x3.")
        self.levels = list(levels)
        self.min_performance_fraction = float(min_performance_fraction)
        self.H, self.W = shape
        n = len(levels)
        self.board = np.stack([np.asarray(lv["board"], dtype=np.uint16) for lv in levels])
        self.goals = np.stack([np.asarray(lv["goals"], dtype=np.uint16) for lv in levels])
        self.loc = np.array([(-1, -1) if lv["agent_loc"] is None else lv["agent_loc"]
                             for lv in levels], dtype=np.int32).reshape(n, 2)
        self.table = np.stack([np.asarray(lv["points_table"], dtype=np.int64).reshape(72)
                               for lv in levels]).astype(np.int32)
        exits = [np.flatnonzero((b.ravel() & (EXIT | AGENT)) == EXIT) for b in self.board]
        self.max_exits = max(1, max(len(e) for e in exits))
        self.exits = np.full((n, self.max_exits), -1, dtype=np.int32)
        for i, e in enumerate(exits):
            self.exits[i, :len(e)] = e
        self.init_points = np.zeros(n, dtype=np.int32)
        self.required = np.zeros(n, dtype=np.int32)
        self.spawn_thr = np.zeros(n, dtype=np.uint64)
        for i, lv in enumerate(levels):
            table = self.table[i].reshape(8, 9).astype(np.int64)
            counts = alive_counts_np(self.board[i], self.goals[i])
            init_points = int((table * counts).sum())
            colors = np.zeros(9, dtype=bool)
            gen = self.board[i][(self.board[i] & (AGENT | ALIVE | SPAWNING)) > 0] & COLORS
            colors[np.unique(gen) >> 9] = True
            colors[8] = True
            max_points = (table * colors).max(axis=1)
            available = int((max_points * counts.sum(axis=1)).sum()) - init_points
            # MinPerformanceScheduler scales min_performance at reset (env_wrappers.py:131-145)
            mp = float(lv["min_performance"]) * float(min_performance_fraction)
            self.init_points[i] = init_points
            self.required[i] = max(0, int(np.int64(np.ceil(mp * available))))
            self.spawn_thr[i] = spawn_threshold(lv["spawn_prob"])

    def __len__(self):
        return len(self.levels)
