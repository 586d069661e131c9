"""
Where a host-session step spends its time: wall clock per step on the host beside the
device-side timeline of the last step (globaltimer stamps the session kernel leaves in
dev_ctrl, see csrc/session.cuh).

    python tools/trace_session.py [envs] [steps]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    import torch
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.levels import load_levels
    levels = load_levels(os.path.join(ROOT, "tests", "golden", "pools", "append-still-25.npz"))
    env = BatchedSafeLifeEnv(levels, n, time_limit=1000, seed=3)
    env.reset()
    acts = np.random.default_rng(0).integers(0, 9, (steps, n)).astype(np.uint8)
    names = ["courier: poll starts", "courier: all tags seen", "courier: published", "CTA 0: sees the step",
             "CTA 0: phase A done", "CTA 0: records written", "last arrival",
             "last arrival: fenced, done word sent"]
    env.begin_host_session()
    timing = np.zeros(2)
    rows, walls, inside = [], [], []
    for t in range(steps):
        t0 = time.perf_counter()
        env.step_host(acts[t], timing=timing)
        walls.append(1e6 * (time.perf_counter() - t0))
        inside.append((timing[0], timing[1]))
        if t >= 10 and t % 8 == 0:
            st = env.session_timeline()[:8]
            rows.append(st - st[0])
            env.begin_host_session()
            env.step_host(acts[t])
    env.end_host_session()
    walls = np.array(walls[10:])
    print("%d envs: host wall clock per step_host call: median %.2f us, p10 %.2f, p90 %.2f" % (
        n, np.median(walls), np.percentile(walls, 10), np.percentile(walls, 90)))
    inside = np.array(inside[10:])
    print("inside sl_host_session_step (median us): post %.2f, wait %.2f" % (
        np.median(inside[:, 0]), np.median(inside[:, 1])))
    rows = np.array(rows)
    print("device timeline of a step (ns after the courier's successful poll began; median over %d steps)" % len(rows))
    for k, name in enumerate(names):
        print("  %-28s %8.0f" % (name, np.median(rows[:, k])))


if __name__ == "__main__":
    main()
