"""
Per-CTA timeline of the single-pass step kernel (k_env_step_warp) on the GPU box.

Builds a diagnostic copy of the library with -DSL_TRACE (globaltimer marks per
CTA: start, loads landed, phase A done, CTA barrier passed, phase B done,
stores read), runs a few L2-flushed steps of the bench workload and prints the
distribution of every mark relative to the first CTA start of the launch.

    python tools/trace_step.py [envs] [pool]
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "safelife_b200", "libsafelife_b200_trace.so")


def build():
    from safelife_b200 import _lib
    cmd = ["nvcc"] + _lib.NVCC_FLAGS + ["-DSL_TRACE", "-o", TRACE_LIB,
                                        os.path.join(ROOT, "safelife_b200", "csrc", "abi.cu")]
    subprocess.check_call(cmd)


def main():
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(pos[0]) if len(pos) > 0 else 4096
    pool = pos[1] if len(pos) > 1 else "append-still-25"
    if not os.path.exists(TRACE_LIB) or "--build" in sys.argv:
        build()
    if "--build-only" in sys.argv:
        return
    os.environ["SAFELIFE_B200_LIB"] = TRACE_LIB
    import numpy as np
    import torch
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.levels import load_levels
    levels = load_levels(os.path.join(ROOT, "tests", "golden", "pools", pool + ".npz"))
    dev = torch.device("cuda", 0)
    env = BatchedSafeLifeEnv([levels[i % len(levels)] for i in range(n)], n, device=dev, time_limit=1000)
    env.reset()
    lib = _lib.lib()
    lib.sl_debug_trace.argtypes = [ctypes.c_void_p]
    grid = n                   # upper bound: the single-pass kernel takes 1..8 envs per CTA
    buf = torch.zeros((grid, 8), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    actions = torch.randint(0, 9, (24, n), device=dev, dtype=torch.uint8)
    stream = env._stream()
    for t in range(8):
        _lib.check(lib.sl_env_step(env._batch_ref, actions[t].data_ptr(), 2, 1, stream))
    _lib.check(lib.sl_debug_trace(buf.data_ptr()))
    # marks of the single-pass kernel without per-warp tails: 0 start, 1 loads landed,
    # 2 phase A done (warp 0), 4 all warps through phase A (before the CTA barrier), 5 stores read
    names, cols = ["start", "loaded", "phaseA", "phaseA_all", "stored"], [0, 1, 2, 4, 5]
    acc = []
    for t in range(8, 24):
        flush.fill_(t)
        buf.zero_()
        torch.cuda.synchronize()
        _lib.check(lib.sl_env_step(env._batch_ref, actions[t].data_ptr(), 2, 1, stream))
        torch.cuda.synchronize()
        a = buf.cpu().numpy()[:, cols].astype(np.float64)
        a = a[a[:, 0] > 0]
        acc.append(a - a[:, 0].min())
    a = np.concatenate(acc)
    print("%d envs, %d CTAs x %d launches (%s); ns since the first CTA start of the launch" % (
        n, len(acc[0]), len(acc), env.step_plan()[0]))
    print("%-10s %8s %8s %8s %8s" % ("mark", "min", "median", "p90", "max"))
    for k, name in enumerate(names):
        v = a[:, k]
        print("%-10s %8.0f %8.0f %8.0f %8.0f" % (name, v.min(), np.median(v), np.percentile(v, 90), v.max()))
    print("per-CTA durations (median ns): load %.0f  phaseA %.0f  wait for the CTA's warps %.0f  "
          "tails + stores %.0f" % tuple(np.median(a[:, k + 1] - a[:, k]) for k in range(4)))


if __name__ == "__main__":
    main()
