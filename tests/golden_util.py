"""Helpers shared by the CPU and GPU parity tests: golden fixture access."""
import glob
import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def traj_groups():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


def load_traj(group):
    with np.load(os.path.join(GOLDEN, "traj_%s.npz" % group)) as d:
        return {k: d[k] for k in d.files}


def levels_of(d):
    """Fixture arrays -> list of level dicts (the format OracleEnv / the product loader take)."""
    out = []
    for i in range(len(d["board"])):
        loc = d["agent_loc"][i]
        out.append({
            "board": d["board"][i], "goals": d["goals"][i],
            "agent_loc": None if loc[0] < 0 else (int(loc[0]), int(loc[1])),
            "points_table": d["points_table"][i],
            "min_performance": float(d["min_performance"][i]),
            "spawn_prob": float(d["spawn_prob"][i]),
        })
    return out


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as d:
        return {k: d[k] for k in d.files}
