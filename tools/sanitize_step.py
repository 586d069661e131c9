"""Workload for compute-sanitizer (racecheck / synccheck / memcheck): a few steps
of every step-kernel variant that mixes mbarriers, async-proxy fences, named
barriers, the scratch-slab lock and shared-memory flags.
    compute-sanitizer --tool racecheck python tools/sanitize_step.py
(prints nothing that is a bench value)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from safelife_b200 import speedups, wrappers
    from safelife_b200.env import BatchedSafeLifeEnv
    from test_gpu_env import _random_levels, TRAIN_CHANNELS
    only = sys.argv[1:] or ["single", "epc", "pipe", "pipe_obs", "in_place_dynamic", "tile", "multistep",
                            "session"]
    rs = np.random.default_rng(1)
    lv25 = _random_levels(rs, 16, 25, 25)            # spawners, dynamic goals on ~30 % of the levels
    lv50 = _random_levels(rs, 8, 50, 50)

    def run(name, env, steps, top=None):
        top = top or env
        top.reset()
        n = env.num_envs
        for t in range(steps):
            top.step(torch.randint(0, 9, (n,), device="cuda", dtype=torch.uint8))
        torch.cuda.synchronize()
        print("ok", name, env.step_plan(obs=True)[0])

    if "single" in only:      # single-pass, per-warp tail, observation + wrapper epilogue, resets
        env = BatchedSafeLifeEnv(lv25, 64, time_limit=3, auto_reset=True, kernel=2, view_shape=(25, 25),
                                 output_channels=TRAIN_CHANNELS)
        run("single-pass + obs + wrappers", env, 7, wrappers.training_stack(env, se_penalty=0.5))
        env = BatchedSafeLifeEnv(lv25, 67, time_limit=3, auto_reset=True, kernel=2)
        for t in range(5):
            env.step_physics(torch.randint(0, 9, (67,), device="cuda", dtype=torch.uint8), auto_reset=True)
        torch.cuda.synchronize()
        print("ok single-pass, scalar tail on warp 0, ragged tail group")
    if "epc" in only:         # 1..8 envs per CTA: slab-offset bulk loads, edge stores next to bulk stores
        for n in (5, 1500, 2900, 4096):
            env = BatchedSafeLifeEnv(lv25, n, time_limit=3, auto_reset=True, kernel=2, view_shape=(9, 9),
                                     output_channels=TRAIN_CHANNELS)
            env.reset()
            for t in range(3):
                a = torch.randint(0, 9, (n,), device="cuda", dtype=torch.uint8)
                env.step(a)
                env.step_physics(a, auto_reset=True)
            torch.cuda.synchronize()
            print("ok", env.step_plan()[0])
    if "session" in only:     # the resident session kernel (mailboxes in pinned memory)
        env = BatchedSafeLifeEnv(lv25, 300, time_limit=3, auto_reset=True, kernel=2)
        env.reset()
        env.begin_host_session(idle_timeout_ms=20000)
        for t in range(5):
            env.step_host(rs.integers(0, 9, 300).astype(np.uint8))
        env.end_host_session()
        torch.cuda.synchronize()
        print("ok host session")
    if "pipe" in only:        # persistent pipelined kernel, scalar warp
        n = 148 * 4 * 2 * 8 + 8 * 9 + 5
        env = BatchedSafeLifeEnv(lv25, n, time_limit=2, auto_reset=True, kernel=2, goals_in_place=False)
        env.reset()
        for t in range(3):
            env.step_physics(torch.randint(0, 9, (n,), device="cuda", dtype=torch.uint8), auto_reset=True)
        torch.cuda.synchronize()
        print("ok", env.step_plan()[0])
    if "pipe_obs" in only:
        n = 148 * 4 * 2 * 8 + 8 * 9
        env = BatchedSafeLifeEnv(lv25, n, time_limit=2, auto_reset=True, kernel=3, view_shape=(25, 25),
                                 output_channels=TRAIN_CHANNELS)
        run("pipelined + obs", env, 3)
    if "in_place_dynamic" in only:   # goals left in HBM although some are dynamic: scratch slab + lock
        env = BatchedSafeLifeEnv(lv25, 128, time_limit=4, auto_reset=True, kernel=2, goals_in_place=True)
        env.reset()
        for t in range(6):
            env.step_physics(torch.randint(0, 9, (128,), device="cuda", dtype=torch.uint8), auto_reset=True)
        torch.cuda.synchronize()
        print("ok", env.step_plan()[0])
        n = 148 * 4 * 2 * 8 + 16
        env = BatchedSafeLifeEnv(lv25, n, time_limit=2, auto_reset=True, kernel=2, goals_in_place=True)
        env.reset()
        for t in range(3):
            env.step_physics(torch.randint(0, 9, (n,), device="cuda", dtype=torch.uint8), auto_reset=True)
        torch.cuda.synchronize()
        print("ok", env.step_plan()[0])
    if "tile" in only:        # several warps per env: named barriers, shared counters
        env = BatchedSafeLifeEnv(lv50, 33, time_limit=3, auto_reset=True, kernel=2, view_shape=(25, 25),
                                 output_channels=TRAIN_CHANNELS)
        run("tiles + obs + wrappers", env, 7, wrappers.training_stack(env, se_penalty=0.5))
    if "multistep" in only:   # register-resident multi-step kernels incl. the tiled one
        for shape in ((25, 25), (50, 50), (70, 33)):
            b = np.stack([lv["board"] for lv in _random_levels(rs, 3, *shape)])
            speedups.set_bit_generator(np.random.default_rng(3).bit_generator)
            for board in b:
                speedups.advance_board(board, 0.3, 12)
                speedups.life_occupancy(board, 0.3, 12)
        torch.cuda.synchronize()
        print("ok multistep")


if __name__ == "__main__":
    main()
