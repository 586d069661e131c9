"""Where the microseconds of a host-driven step go (diagnostic, not a bench value):
wall clock of launch + completion for an empty kernel and for the step kernel, with
the outputs in pinned host memory (zero-copy) or in device memory, waiting on the
pinned completion word or on cudaStreamSynchronize.
    python tools/host_latency.py [--envs 4096]
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=300)
    args = ap.parse_args()
    import torch
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.levels import load_levels
    levels = load_levels(os.path.join(ROOT, "tests", "golden", "pools", "append-still-25.npz"))
    env = BatchedSafeLifeEnv(levels, args.envs, time_limit=1000, seed=3)
    env.reset()
    env.enable_host_io()
    lib = env._lib
    n = args.envs
    acts = np.random.default_rng(0).integers(0, 9, n).astype(np.uint8)
    dev_actions = torch.from_numpy(acts).cuda()
    stream = env._stream()
    timing = np.zeros(2)

    def run(name, fn):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e6
        print("%-58s median %6.2f us  p10 %6.2f  p90 %6.2f" % (name, np.median(ts), np.percentile(ts, 10),
                                                            np.percentile(ts, 90)))

    def empty_sync():
        lib.sl_launch_floor(env._batch_ref, stream)
        torch.cuda.current_stream().synchronize()
    run("empty kernel + torch stream.synchronize()", empty_sync)

    def step_dev_sync():
        lib.sl_env_step(env._batch_ref, dev_actions.data_ptr(), 0, 1, stream)
        torch.cuda.current_stream().synchronize()
    run("step (device outputs) + torch stream.synchronize()", step_dev_sync)

    def step_dev_flag():
        lib.sl_env_step_host(env._batch_ref, dev_actions.data_ptr(), 0, 1, stream, timing.ctypes.data)
    run("step (device actions/outputs) + completion word", step_dev_flag)
    print("    inside: launch %.2f us, wait %.2f us" % tuple(timing))

    def step_host_flag():
        lib.sl_env_step_host(env._batch_host_ref, env._host_actions_ptr, 0, 1, stream, timing.ctypes.data)
    np.copyto(env._host_np["actions"], acts)
    run("step (pinned actions + pinned outputs) + completion word", step_host_flag)
    print("    inside: launch %.2f us, wait %.2f us" % tuple(timing))

    def step_host_dev_actions():
        lib.sl_env_step_host(env._batch_host_ref, dev_actions.data_ptr(), 0, 1, stream, timing.ctypes.data)
    run("step (device actions + pinned outputs) + completion word", step_host_dev_actions)
    print("    inside: launch %.2f us, wait %.2f us" % tuple(timing))

    os.environ["SAFELIFE_B200_HOST_WAIT"] = "0"
    run("public step_host()", lambda: env.step_host(acts))


if __name__ == "__main__":
    main()
