"""Driver for ncu captures of the step kernels: a few steps of one configuration.
    python tools/prof_step.py [--pool NAME] [--envs N] [--steps K] [--obs] [--view 25]
(run under `ncu -k regex:k_env_step ...`; nothing printed here is a bench value)
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool", default="append-still-25")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--obs", action="store_true")
    ap.add_argument("--kernel", type=int, default=0)
    args = ap.parse_args()
    import torch
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.levels import load_levels
    levels = load_levels(os.path.join(ROOT, "tests", "golden", "pools", args.pool + ".npz"))
    ch = (0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 25, 26, 27, 12, 13)
    env = BatchedSafeLifeEnv(levels, args.envs, kernel=args.kernel, time_limit=1000, seed=3,
                             view_shape=(25, 25), output_channels=ch)
    env.reset()
    acts = torch.randint(0, 9, (args.steps, args.envs), device="cuda", dtype=torch.uint8)
    print(env.step_plan(obs=args.obs))
    for t in range(args.steps):
        spec = ctypes.byref(env._obs_specs[t & 1]) if args.obs else None
        _lib.check(env._lib.sl_env_step_obs(env._batch_ref, acts[t].data_ptr(), spec, args.kernel,
                                            _lib.SL_STEP_AUTO_RESET, env._stream()))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
