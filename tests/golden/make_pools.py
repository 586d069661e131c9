"""
Generate the level pools the benchmark and the full-size parity tests run on,
with the UNMODIFIED reference procgen (safelife/proc_gen.py:576-687 through
level_iterator._game_from_data) in this container.  Procgen is out of scope
for the engine; the pools are committed as fixtures
because the reference does not travel to the GPU box.

    python tests/golden/make_pools.py [name ...]

BASELINE.json configs: 25x25 overrides of random/append-still (1024 levels,
SeedSequence(1)), random/prune-still (1024, SeedSequence(3)),
random/append-spawn (4096, SeedSequence(4)), 50x50 random/navigation (256,
SeedSequence(5)), plus the reference-native 26x26 append-still.
"""
import multiprocessing
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

POOLS = {
    # name: (yaml, shape, count, seed)
    "append-still-25": ("random/append-still", (25, 25), 1024, 1),
    "append-still-26": ("random/append-still", (26, 26), 1024, 1),
    "prune-still-25": ("random/prune-still", (25, 25), 1024, 3),
    "append-spawn-25": ("random/append-spawn", (25, 25), 4096, 4),
    "navigation-50": ("random/navigation", (50, 50), 256, 5),
}


def _gen(args):
    yaml_name, shape, seed = args
    from oracle import refload
    refload.load_reference()
    from safelife.level_iterator import _load_files, _game_from_data
    from safelife.safelife_game import GameState, SafeLifeGame
    (fname, kind, data), = _load_files([yaml_name])
    data = dict(data, board_shape=list(shape))
    game = _game_from_data(fname, kind, data, seed=seed)
    raw = SafeLifeGame(board_size=None)
    GameState.deserialize(raw, game._init_data, as_initial_state=False)
    loc = raw.agent_locs[0] if len(raw.agent_locs) else (-1, -1)
    return {
        "board": raw.board.copy(), "goals": np.array(game._init_data["goals"], dtype=np.uint16),
        "agent_loc": None if loc[0] < 0 else (int(loc[0]), int(loc[1])),
        "points_table": np.asarray(game.points_table, dtype=np.int64)[0],
        "min_performance": float(game.min_performance), "spawn_prob": float(game.spawn_prob),
    }


def main():
    from safelife_b200.levels import save_pool
    names = sys.argv[1:] or list(POOLS)
    with multiprocessing.Pool(os.cpu_count()) as pool:
        for name in names:
            yaml_name, shape, count, seed = POOLS[name]
            t0 = time.time()
            seeds = np.random.SeedSequence(seed).spawn(count)
            levels = pool.map(_gen, [(yaml_name, shape, s) for s in seeds], chunksize=4)
            path = os.path.join(HERE, "pools", name + ".npz")
            save_pool(path, levels)
            print("wrote %s: %d levels in %.0f s, %d KiB" % (
                path, count, time.time() - t0, os.path.getsize(path) // 1024), flush=True)


if __name__ == "__main__":
    main()
