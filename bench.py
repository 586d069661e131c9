#!/usr/bin/env python
"""
bench.py -- env-steps/sec of the batched SafeLife step path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--envs E] [--pool NAME] [--kernel 0|1|2]

One "step" = one pass of the hot path (SafeLifeEnv.step minus observation:
execute_actions + advance_board (+ goals when dynamic) + points / exit colours /
reward / done, then the auto-reset of finished envs) over the whole batch.
Workload (BASELINE.json configs[1]): 4096 envs per GPU of 25x25 append-still
procedural boards (pool generated with the reference procgen, committed under
tests/golden/pools), uniform random actions, time_limit 1000.  N GPUs run N
independent shards of 4096 envs (weak scaling, no collective on the step path).

Prints ONE JSON line on rank 0.  `value` is device-timed (one CUDA-event pair around
the K steps) with the state resident in HBM and the steps rotating over batches that
together exceed the L2, so that every launch reads from HBM; `e2e` goes through the public
BatchedSafeLifeEnv API with pinned host actions in and host reward/done out.
`--impl reference` times the reference's own C extension (oracle/_ref, compiled
from the reference sources) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
POOL_DIR = os.path.join(ROOT, "tests", "golden", "pools")
METRIC = "env-steps/sec (batched boards) at 1/2/4/8 B200 vs C-ext on host cores"


_REAL_STDOUT = None


def quiet_stdout():
    """
    Everything but the JSON line goes to stderr: libraries (NCCL prints its version
    banner on stdout) must not add lines to the one-line contract.
    """
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--pool", default="append-still-25")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps")
    ap.add_argument("--flush-mode", default="fill", choices=["fill", "fill+read"],
                    help="fill: write 256 MiB; fill+read: then read another 256 MiB, so that the "
                         "L2 holds clean lines (no write-back of the flush inside the timed step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3-5")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--batches", type=int, default=0,
                    help="batches of the rotation (0: enough for three times the L2); a small number "
                         "keeps the set-up short for launch-list captures and is NOT a valid timing")
    ap.add_argument("--quick", action="store_true",
                    help="A/B runs: only `value` and `step_with_obs` (rotating batches), printed as a short line")
    return ap.parse_args()


def kernel_source_hash():
    """sha1 of the CUDA sources: the ncu traffic figures are only valid for the code they measured."""
    import hashlib
    h = hashlib.sha1()
    src = os.path.join(ROOT, "safelife_b200", "csrc")
    for name in sorted(os.listdir(src)):
        with open(os.path.join(src, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()


def measured_traffic(plan, n, H, W):
    """
    DRAM bytes (read + write) per launch of the kernel `plan` names, from the ncu
    capture recorded in profiles/r2_traffic.json by tools/ncu_traffic.py -- or None
    if there is no capture of this kernel at this batch, or if the CUDA sources have
    changed since the capture (the file carries their hash).
    """
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            data = json.load(f)
        if data.get("source_sha1") != kernel_source_hash():
            return None
        entry = data["entries"].get("%s|%dx%dx%d" % (plan.split(" grid ")[0], n, H, W))
        return entry["dram_bytes_read"] + entry["dram_bytes_write"] if entry else None
    except (OSError, KeyError, ValueError):
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------
# CPU arm: the reference C extension ("C-ext trio")
# --------------------------------------------------------------------------
def _cpu_worker(args):
    """Steps a subset of envs with the reference speedups; returns timed seconds."""
    levels, seeds, actions, warmup, steps, barrier, kind = args
    if kind == "reference":
        from oracle import refload
        sp = refload.load_ref_speedups()
    else:
        sp = None
    from oracle import oracle as orc
    n = len(levels)
    boards = [np.array(lv["board"], dtype=np.uint16) for lv in levels]
    goals = [np.array(lv["goals"], dtype=np.uint16) for lv in levels]
    locs = [np.array([lv["agent_loc"]], dtype=np.int64) for lv in levels]
    tables = [np.asarray(lv["points_table"], dtype=np.int64).reshape(8, 9) for lv in levels]
    gens = [np.random.default_rng(s) for s in seeds]
    rngs = [orc.rng_from_generator(g) for g in gens]
    static = [None] * n
    total = 0

    def one_step(t):
        nonlocal total
        for i in range(n):
            a = int(actions[t, i])
            p = levels[i]["spawn_prob"]
            if kind == "reference":
                # SafeLifeGame.execute_actions / advance_board / alive_counts
                # (safelife_game.py:380-389, 746-761, 657-663) on the reference C code
                sp.execute_actions(boards[i], locs[i], a)
                sp.set_bit_generator(gens[i].bit_generator)
                boards[i] = sp.advance_board(boards[i], p)
                if not static[i]:
                    new_goals = sp.advance_board(goals[i], p)
                    if static[i] is None:
                        static[i] = (not (new_goals & 128).any()) and (new_goals == goals[i]).all()
                    goals[i] = new_goals
                counts = sp.alive_counts(boards[i], goals[i])
            else:
                orc.execute_actions(boards[i], locs[i], a)
                boards[i] = orc.advance_board(boards[i], p, 1, rngs[i])
                if not static[i]:
                    new_goals = orc.advance_board(goals[i], p, 1, rngs[i])
                    if static[i] is None:
                        static[i] = (not (new_goals & 128).any()) and (new_goals == goals[i]).all()
                    goals[i] = new_goals
                counts = orc.alive_counts(boards[i], goals[i])
            total += int((tables[i] * counts).sum())

    for t in range(warmup):
        one_step(t)
    barrier.wait()
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        one_step(t)
    return time.perf_counter() - t0, total


def cpu_arm(levels, n_envs, steps, warmup, seconds_budget):
    """
    Aggregate env-steps/s of the reference C-extension trio + 72-term dot
    product, one process per host core.  The sample (envs per step) is bounded
    so that warmup + steps fit the time budget.
    """
    import multiprocessing as mp
    from oracle import refload
    kind = "reference" if refload.have_ref_so() else "port"
    cores = os.cpu_count() or 1
    # calibrate on one core
    m = mp.Manager()
    probe = 64
    acts = np.random.default_rng(7).integers(0, 9, (warmup + steps + 2, max(n_envs, probe))).astype(np.uint8)
    t, _ = _cpu_worker(([levels[i % len(levels)] for i in range(probe)], list(range(probe)),
                        acts, 0, 2, m.Barrier(1), kind))
    per_env_step = t / (2 * probe)
    cap = int(seconds_budget * cores / (per_env_step * (warmup + steps)))
    sample = max(cores, min(n_envs, cap))
    sample -= sample % cores
    per = sample // cores
    barrier = m.Barrier(cores)
    jobs = []
    for w in range(cores):
        idx = list(range(w * per, (w + 1) * per))
        jobs.append(([levels[i % len(levels)] for i in idx], idx, acts[:, idx], warmup, steps,
                     barrier, kind))
    with mp.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    elapsed = max(r[0] for r in res)
    value = sample * steps / elapsed
    extra = _cpu_pure_c(levels) if kind == "reference" else {}
    try:
        extra.update(_cpu_full_env(levels))
    except Exception as exc:   # noqa: BLE001  (a baseline extra must not take the bench down)
        extra["reference_env_step"] = None
        extra["reference_env_step_note"] = repr(exc)
    return {
        "value": value, "unit": "env-steps/s", "cores": cores, "kind": kind,
        "sample": "%d envs x %d steps (of %d envs/step), %s speedups execute_actions+advance_board"
                  "+alive_counts+dot from a Python loop, one process per core" % (
                      sample, steps, n_envs, "reference C" if kind == "reference" else "oracle"),
        **extra,
    }, elapsed


def _cpu_full_env_worker(args):
    """
    BASELINE.md section 3.2: the UNMODIFIED reference SafeLifeEnv.step (physics, reward,
    get_obs with the training view) on this process's share of the levels, reset on done.
    """
    levels, seed, warmup, seconds, barrier = args
    from oracle import refload
    refload.load_reference()
    from safelife.safelife_env import SafeLifeEnv
    from safelife.safelife_game import SafeLifeGame

    def games():
        k = 0
        while True:
            lv = levels[k % len(levels)]
            k += 1
            loc = lv["agent_loc"]
            yield SafeLifeGame.loaddata({
                "board": np.array(lv["board"], dtype=np.uint16), "goals": np.array(lv["goals"], dtype=np.uint16),
                "agent_locs": np.zeros((0, 2), dtype=np.int64) if loc is None else np.array([loc], dtype=np.int64),
                "points_table": np.asarray(lv["points_table"], dtype=np.int64).reshape(1, 8, 9),
                "min_performance": lv["min_performance"], "spawn_prob": lv["spawn_prob"]}, auto_cls=False)
    env = SafeLifeEnv(games(), time_limit=1000, view_shape=(25, 25), output_channels=TRAIN_CHANNELS,
                      should_calculate_side_effects=False)
    rs = np.random.default_rng(seed)
    env.reset()
    for _ in range(warmup):
        _, _, done, _ = env.step(int(rs.integers(0, 9)))
        if done:
            env.reset()
    barrier.wait()
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            _, _, done, _ = env.step(int(rs.integers(0, 9)))
            if done:
                env.reset()
        steps += 50
    return steps, time.perf_counter() - t0


def _cpu_full_env(levels, seconds=4.0):
    """Aggregate env-steps/s of the reference's own SafeLifeEnv.step, one process per host core."""
    import multiprocessing as mp
    from oracle import refload
    if not refload.have_reference():
        return {"reference_env_step": None,
                "reference_env_step_note": "reference Python not shipped (baseline/_ref missing)"}
    cores = os.cpu_count() or 1
    m = mp.Manager()
    barrier = m.Barrier(cores)
    jobs = [([levels[(w * 8 + k) % len(levels)] for k in range(8)], 100 + w, 20, seconds, barrier)
            for w in range(cores)]
    with mp.Pool(cores) as pool:
        res = pool.map(_cpu_full_env_worker, jobs)
    rate = sum(r[0] / r[1] for r in res)
    return {"reference_env_step": {
        "value": rate, "unit": "env-steps/s", "cores": cores, "per_core": rate / cores,
        "sample": "unmodified reference SafeLifeEnv.step (physics + reward + get_obs, 25x25 view, 15 "
                  "channels, reset on done), one env per process, one process per core, %.0f s" % seconds}}


def _cpu_pure_c(levels):
    """
    BASELINE.md section 3, items 3 and 4, on ONE host core: the reference's C loops
    with no Python per step -- advance_board(board, p, n_step=1000) and
    life_occupancy(board, p, 1000), in board-steps/s.
    """
    from oracle import refload
    sp = refload.load_ref_speedups()
    gen = np.random.default_rng(11)
    sp.set_bit_generator(gen.bit_generator)
    boards = [np.array(lv["board"], dtype=np.uint16) for lv in levels[:6]]
    probs = [float(lv["spawn_prob"]) for lv in levels[:6]]
    t0 = time.perf_counter()
    for b, p in zip(boards, probs):
        sp.advance_board(b, p, 1000)
    t_adv = time.perf_counter() - t0
    t0 = time.perf_counter()
    for b, p in zip(boards[:3], probs[:3]):
        sp.life_occupancy(b, p, 1000)
    t_occ = time.perf_counter() - t0
    return {"pure_c_advance_board_steps_per_s_one_core": len(boards) * 1000 / t_adv,
            "pure_c_life_occupancy_steps_per_s_one_core": 3 * 1000 / t_occ}


def run_reference(args, levels):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total = args.envs * args.gpus
    base, elapsed = cpu_arm(levels, n_total, args.steps, args.warmup, 60.0)
    H, W = levels[0]["board"].shape
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        # the same workload string as our arm's line; the CPU sample of it is in cpu_baseline.sample
        "config": {"workload": "%d envs/GPU x %dx%d %s procgen pool (%d levels), uniform random "
                               "actions, time_limit 1000, auto-reset" % (args.envs, H, W, args.pool, len(levels)),
                   "envs_per_gpu": args.envs, "board": [H, W], "pool": args.pool,
                   "timing": "wall clock around the timed steps of a bounded sample of the workload "
                             "on the host cores (cpu_baseline.sample)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons (NVML) polled while the timed regions run."""

    def __init__(self, index, period_s=0.004):
        self.sm, self.reasons, self.max_mhz, self.err = [], 0, None, None
        self.marks = []
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:   # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, args=(period_s,), daemon=True)
            self.thread.start()
        except Exception as exc:   # noqa: BLE001
            self.err = repr(exc)
            self.thread = None

    def _run(self, period_s):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append((time.perf_counter(),
                                float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception as exc:   # noqa: BLE001
                self.err = repr(exc)
                break
            time.sleep(period_s)

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml: %s" % self.err]}
        self._stop.set()
        self.thread.join(timeout=2)
        nv = self.nv
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (0, float("inf"))
        inside = [v for t, v in self.sm if lo <= t <= hi] or [v for _, v in self.sm]
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        return {"sm_mhz": float(np.median(inside)) if inside else None,
                "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, bit in names.items() if self.reasons & bit),
                "samples": len(inside), "source": "nvml, sampled during the timed regions"}


class _Flush(object):
    """L2 flush between timed steps (never inside a timed interval)."""

    def __init__(self, torch, dev, mode):
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.mode = mode
        self.clean = torch.zeros(256 << 18, dtype=torch.int32, device=dev) if mode == "fill+read" else None

    def fill_(self, v):
        self.buf.fill_(v)
        if self.clean is not None:
            self.sink = self.clean.max()      # a 256 MiB read: the L2 ends up with clean lines


TRAIN_CHANNELS = (0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 25, 26, 27, 12, 13)   # env_factory.py:311-327


def step_bytes(H, W, n, dyn, baseline=False):
    """Algorithmic bytes of one launch: static goals 3*2HW + 64, dynamic 4*2HW + 64."""
    extra = 2 * H * W if baseline else 0
    return (n - dyn) * (6 * H * W + 64 + extra) + dyn * (8 * H * W + 64 + extra)


def timed_steps(torch, launch, K, flush, first=0):
    """Mean ms of `launch(t)` over K steps, one CUDA-event pair per step, L2 flushed before each."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    torch.cuda.synchronize()
    for t in range(K):
        if flush is not None:
            flush.fill_(t & 0xFF)
        ev[t][0].record()
        launch(first + t)
        ev[t][1].record()
    torch.cuda.synchronize()
    return sum(e[0].elapsed_time(e[1]) for e in ev) / K


class Rotation(object):
    """
    Inputs larger than the L2: R independent batches of the same workload whose board
    planes together are at least three times the L2, stepped round-robin.  When a batch comes
    round again everything it touches has been evicted, so every launch reads its state
    from HBM (and the previous launches' dirty lines are written back inside the timed
    region) -- with no flush kernel between the launches and ONE event pair around the K
    timed steps.  Each step is one launch over one batch of n envs, as in the workload.
    """

    forced = 0          # --batches

    def __init__(self, torch, dev, first_env, make_env, board_bytes):
        l2 = int(torch.cuda.get_device_properties(dev).L2_cache_size)
        self.l2_bytes = l2
        self.R = int(max(2, min(96, -(-3 * l2 // board_bytes))))
        if Rotation.forced:
            self.R = Rotation.forced
        self.board_bytes = board_bytes
        self.envs = [first_env] + [make_env(r) for r in range(1, self.R)]

    def note(self):
        if self.R * self.board_bytes < 2 * self.l2_bytes:
            return ("NOT larger than L2 (--batches %d: %.0f MB of board planes against %.0f MB of L2): "
                    "a launch-list / debugging run, not a timing" % (
                        self.R, self.R * self.board_bytes / 1e6, self.l2_bytes / 1e6))
        return ("inputs larger than L2: steps rotate over %d independent batches of the workload "
                "(board planes %.1f MB each, %.0f MB together against %.0f MB of L2), so every "
                "launch reads its batch from HBM; no flush inside or between the timed steps" % (
                    self.R, self.board_bytes / 1e6, self.R * self.board_bytes / 1e6, self.l2_bytes / 1e6))

    def timed(self, torch, launch, K, first=0):
        """ms per step of `launch(env, t)` over K back-to-back steps, one CUDA-event pair."""
        for r in range(self.R):                 # every batch has run this launch once (untimed)
            launch(self.envs[r], first + r)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for t in range(K):
            launch(self.envs[t % self.R], first + t)
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / K


def sweep(args, levels, dev, peak):
    """
    The same device-timed measurement (rotating batches larger than the L2, one event
    pair around 60 steps) at larger batches: where the kernel leaves the latency regime;
    with and without the observation in the same launch.
    """
    import ctypes
    import torch
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv
    out = []
    H, W = levels[0]["board"].shape
    pool = None
    for n in (16384, 65536):
        def make(r):
            e = BatchedSafeLifeEnv(pool if pool is not None else levels, n, device=dev,
                                   kernel=args.kernel, time_limit=1000, view_shape=(25, 25),
                                   output_channels=TRAIN_CHANNELS, seed=300 + r)
            e.reset()
            return e
        env = make(0)
        pool = env.pool
        rot = Rotation(torch, dev, env, make, n * 2 * H * W)
        K = 60
        actions = torch.randint(0, 9, (K + rot.R + 5, n), device=dev, dtype=torch.uint8)

        def physics(e, t):
            _lib.check(e._lib.sl_env_step(e._batch_ref, actions[t].data_ptr(), args.kernel,
                                          _lib.SL_STEP_AUTO_RESET, e._stream()))

        def with_obs(e, t):
            _lib.check(e._lib.sl_env_step_obs(e._batch_ref, actions[t].data_ptr(),
                                              ctypes.byref(e._obs_specs[0]), args.kernel,
                                              _lib.SL_STEP_AUTO_RESET, e._stream()))
        ms = rot.timed(torch, physics, K)
        ms_obs = rot.timed(torch, with_obs, K)
        nbytes = step_bytes(H, W, n, 0)
        obs_bytes = nbytes + n * (2 * H * W + 25 * 25 * len(TRAIN_CHANNELS))
        plan = env.step_plan()[0]
        out.append({"envs": n, "batches": rot.R, "ms_per_step": ms, "env_steps_per_s": n / (ms * 1e-3),
                    "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                    "kernel": plan, "traffic": measured_traffic(plan, n, H, W),
                    "step_with_obs": {"ms": ms_obs, "env_steps_per_s": n / (ms_obs * 1e-3),
                                      "achieved_gbs": obs_bytes / (ms_obs * 1e-3) / 1e9,
                                      "frac": obs_bytes / (ms_obs * 1e-3) / 1e9 / peak,
                                      "kernel": env.step_plan(obs=True)[0]}})
        del env, rot
        torch.cuda.empty_cache()
    return out


def kernels_agree(torch, levels, dev, n=64, steps=16):
    """
    In-run parity spot check without the CPU oracle (which only the tests may use):
    the warp bit-plane kernels against the independent generic CTA-per-env kernel
    on the first n envs of the workload -- boards, goals, RNG and rewards bit for bit.
    """
    from safelife_b200.env import BatchedSafeLifeEnv, seed_rng_states
    rng = seed_rng_states(11, n)
    envs = []
    for kernel in (0, 1):
        e = BatchedSafeLifeEnv(levels, n, device=dev, kernel=kernel, time_limit=1000, auto_reset=True)
        e.set_rng_state(rng)
        e.reset()
        envs.append(e)
    acts = torch.randint(0, 9, (steps, n), device=dev, dtype=torch.uint8)
    ok = True
    for t in range(steps):
        outs = [e.step_physics(acts[t], auto_reset=True) for e in envs]
        ok &= bool(torch.equal(outs[0]["reward"], outs[1]["reward"]))
    a, b = envs
    ok &= bool(torch.equal(a.board, b.board)) and bool(torch.equal(a.goals, b.goals))
    ok &= bool(torch.equal(a._t["rng"], b._t["rng"]))
    return ok


def other_configs(args, dev, world, rank, peak, barrier, allmax):
    """
    BASELINE.json configs 3, 4 and 5 as per-GPU shares: 16384 prune-still,
    8192 append-spawn (65536 / 8) and 4096 navigation 50x50 (32768 / 8) envs per GPU.
    At --gpus 8 the totals are exactly configs 4 and 5.  Config 5 steps with the
    SimpleSideEffectPenalty cell count fused into the launch (sl_env_step_wrapped,
    starting-state baseline) and adds the episode-end simulation of side_effect_score
    (sl_side_effect_occupancy on one finished episode per pool level) in board-steps/s.
    """
    import torch
    from safelife_b200 import _lib, wrappers
    from safelife_b200.env import BatchedSafeLifeEnv, seed_rng_states
    from safelife_b200.levels import load_levels
    out = []
    flush = None if args.no_flush else _Flush(torch, dev, args.flush_mode)
    K = max(20, min(args.steps, 100))
    for cfg, pool, n, penalty in ((3, "prune-still-25", 16384, False), (4, "append-spawn-25", 8192, False),
                                  (5, "navigation-50", 4096, True)):
        path = os.path.join(POOL_DIR, pool + ".npz")
        if not os.path.exists(path):
            continue
        levels = load_levels(path)
        H, W = levels[0]["board"].shape
        states = seed_rng_states(4, n * world)[rank * n:(rank + 1) * n]
        made = {}

        def make(r):
            e = BatchedSafeLifeEnv(made["pool"] if made else levels, n, device=dev, kernel=args.kernel,
                                   time_limit=1000, env_offset=rank * n + r * 7, total_envs=n * world)
            made["pool"] = e.pool
            st = states.copy()
            st[:, 1] ^= np.uint64(r)                     # batch r: other streams, other first levels
            e.set_rng_state(st)
            top = wrappers.SimpleSideEffectPenalty(e, penalty_coef=1.0) if penalty else e
            top.reset()
            e._bench_top = top
            return e
        env = make(0)
        rot = Rotation(torch, dev, env, make, n * 2 * H * W)
        actions = torch.randint(0, 9, (K + rot.R + 5, n), device=dev, dtype=torch.uint8)

        def launch(e, t):
            if penalty:
                _lib.check(e._lib.sl_env_step_wrapped(
                    e._batch_ref, actions[t].data_ptr(), None, e._wrap.fill_spec(0), args.kernel,
                    _lib.SL_STEP_AUTO_RESET, e._stream()))
            else:
                _lib.check(e._lib.sl_env_step(e._batch_ref, actions[t].data_ptr(), args.kernel,
                                              _lib.SL_STEP_AUTO_RESET, e._stream()))
        for t in range(5):
            launch(env, t)
        barrier()
        ms = allmax(rot.timed(torch, launch, K))
        flushed_ms = None
        if flush is not None:
            flushed_ms = allmax(timed_steps(torch, lambda t: launch(env, t), min(K, 40), flush))
        dyn = int((env._t["static_goals"] == 2).sum())
        nbytes = step_bytes(H, W, n, dyn, baseline=penalty)
        entry = {
            "config": cfg, "workload": "%d envs/GPU x %dx%d %s (%d levels)%s" % (
                n, H, W, pool, len(levels), ", SimpleSideEffectPenalty fused" if penalty else ""),
            "envs_total": n * world, "ms_per_step": ms, "value": n * world / (ms * 1e-3),
            "unit": "env-steps/s", "kernel": env.step_plan()[0], "dynamic_goal_envs": dyn,
            "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": nbytes},
            "l2": rot.note(), "single_step_flushed_ms": flushed_ms,
            "parity_spot_check": {"warp_kernels_equal_generic_kernel": kernels_agree(torch, levels, dev),
                                  "note": "64 envs x 16 steps, bit for bit; oracle parity: tests/"},
        }
        if penalty:
            # episode end (safelife_env.py:176-187): advance_board(init, num_steps) +
            # 2 x life_occupancy(..., 1000) for one finished episode per pool level
            L, samples, lead = len(levels), 1000, 500
            init = env._pool_t["pool_board"].contiguous()
            final = env._t["board"][:L].contiguous()
            counts = torch.zeros((L, 2, H, W, 8), dtype=torch.int32, device=dev)
            thr = env._pool_t["pool_spawn_thr"].contiguous()
            steps_t = torch.full((L,), lead, dtype=torch.int32, device=dev)
            rng = env._t["rng"][:L].clone().contiguous()

            def occupancy(_t):
                _lib.check(env._lib.sl_side_effect_occupancy(
                    init.data_ptr(), final.data_ptr(), counts.data_ptr(), L, H, W, thr.data_ptr(),
                    steps_t.data_ptr(), samples, rng.data_ptr(), env._stream()))
            occupancy(0)
            occ_ms = allmax(timed_steps(torch, occupancy, 3, None))
            entry["episode_end"] = {
                "what": "side_effect_score simulation (sl_side_effect_occupancy): %d finished episodes "
                        "per GPU, %d lead steps + 2 x %d counted steps each" % (L, lead, samples),
                "ms": occ_ms, "board_steps_per_s": world * L * (lead + 2 * samples) / (occ_ms * 1e-3)}
        out.append(entry)
        del env, rot
        torch.cuda.empty_cache()
    return out


def e2e_line(n, world, K, launch_s, launch_parts, session):
    """
    The end-to-end number: BatchedSafeLifeEnv.step_host with HOST buffers, wall clock.
    With a host session open (begin_host_session) the same call posts the actions in
    pinned memory to a resident kernel; without one it launches a kernel per step.
    The faster of the two modes that ran is the headline, the other is kept beside it.
    """
    base = {"unit": "env-steps/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": 8 * n}
    if session and "seconds" in session and session["seconds"] < launch_s:
        s = session["seconds"]
        return dict(base, value=n * world * K / s, us_per_step=1e6 * s / K,
                    host_us={"post": session["post_us"], "wait": session["wait_us"],
                             "note": "inside sl_host_session_step: tag + copy the actions into the "
                                     "pinned mailbox, then spin on the pinned done word"},
                    timing="wall clock (time.perf_counter) around K back-to-back step_host calls incl. "
                           "the host's read of the result every step; max over ranks.  No L2 flush: in "
                           "a session the envs live in shared memory, every step's inputs (actions) "
                           "come from pinned host memory over PCIe and its results go back the same way",
                    api="BatchedSafeLifeEnv.begin_host_session() + step_host(actions): a resident "
                        "kernel (sl_host_session_*) reads the uint8 actions from pinned host memory "
                        "and writes float32 reward + uint8 done/success/times_up into pinned host "
                        "memory as one packed 8-byte record per env; no kernel launch and no stream "
                        "synchronisation per step",
                    mode="host session")
    return dict(base, value=n * world * K / launch_s, us_per_step=1e6 * launch_s / K,
                host_us={"launch": launch_parts[0] / K, "wait": launch_parts[1] / K},
                timing="wall clock around every step_host call; L2 flushed before each call, outside "
                       "the clock; max over ranks",
                api="BatchedSafeLifeEnv.step_host(auto_reset=True), one launch per step (see "
                    "e2e_launch_per_step)", mode="launch per step",
                session=session)


def rollout_bench(torch, env, steps_per_env=20, reps=3):
    """
    PPO data collection end to end on the device (training/ppo.py:75-143 through
    safelife_b200.rollout.BatchedRollout): every step is the fused gym step (physics +
    auto-reset + observation written straight into the rollout buffer) followed by a
    policy forward pass and an action sample; then trajectory stitching, returns and
    advantages.  The policy is a STUB of the reference's shape (training/models.py:9-42:
    conv 5x5/2 x32, 3x3/2 x64, 3x3 x64, dense 512, 9 logits + value; random weights, bf16,
    plain torch modules) -- it is the consumer, not the product.
    """
    from safelife_b200.rollout import BatchedRollout
    nn = torch.nn
    dev = env.device
    C = len(TRAIN_CHANNELS)

    class Stub(nn.Module):
        def __init__(self):
            super().__init__()
            self.cnn = nn.Sequential(nn.Conv2d(C, 32, 5, 2), nn.ReLU(), nn.Conv2d(32, 64, 3, 2), nn.ReLU(),
                                     nn.Conv2d(64, 64, 3, 1), nn.ReLU())
            self.dense = nn.Sequential(nn.Linear(64 * 3 * 3, 512), nn.ReLU())
            self.logits, self.value = nn.Linear(512, 9), nn.Linear(512, 1)

        def forward(self, obs):
            x = obs.permute(0, 3, 2, 1).to(torch.bfloat16)
            x = self.dense(self.cnn(x).flatten(1))
            return self.value(x)[:, 0].float(), torch.softmax(self.logits(x).float(), dim=-1)
    torch.manual_seed(0)
    net = Stub().to(dev).to(torch.bfloat16).eval()
    runner = BatchedRollout(env)
    with torch.no_grad():
        runner.gen_training_batch(net, steps_per_env=2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            batch = runner.gen_training_batch(net, steps_per_env=steps_per_env)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # the policy alone on the same observations, for the split
        obs = batch.obs[:env.num_envs]
        net(obs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps_per_env):
            net(obs)
        torch.cuda.synchronize()
        policy_ms = 1e3 * (time.perf_counter() - t0) / steps_per_env
    n = env.num_envs
    return {"env_steps_per_s": n * steps_per_env * reps / dt, "ms_per_step": 1e3 * dt / (steps_per_env * reps),
            "policy_forward_ms": policy_ms, "steps_per_env": steps_per_env,
            "note": "supplementary: BatchedRollout.gen_training_batch (fused step + observation, stub CNN "
                    "policy in bf16, multinomial sample, trajectory stitching + GAE), wall clock, one "
                    "sync per batch; rank 0.  policy_forward_ms is the stub alone."}


def graph_replay(torch, env, lib, actions, kernel, reps=10):
    """K steps captured once in a CUDA graph (device-resident actions) and replayed."""
    from safelife_b200 import _lib
    Kg = min(50, actions.shape[0])
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        st = env._stream()
        for k in range(Kg):
            _lib.check(lib.sl_env_step(env._batch_ref, actions[k].data_ptr(), kernel,
                                       _lib.SL_STEP_AUTO_RESET, st))
    g.replay()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        g.replay()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / (reps * Kg), Kg


def run_ours(args, levels):
    Rotation.forced = max(0, args.batches)
    import ctypes
    import torch
    import torch.distributed as dist
    from safelife_b200 import _lib
    from safelife_b200.env import BatchedSafeLifeEnv, seed_rng_states

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (
            args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x):
        """max over ranks of a host float"""
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.envs
    lo = rank * n                      # contiguous env range of this rank
    H, W = levels[0]["board"].shape
    # observations as the reference's training envs take them (env_factory.py:301-327):
    # 25x25 view, 15 channels; `value` is the step path alone, `step_with_obs` adds them
    base_states = seed_rng_states(4, n * world)[lo:lo + n]          # global-index-derived streams
    made = {}

    def make_env(r):
        e = BatchedSafeLifeEnv(made["pool"] if made else levels, n, device=dev, kernel=args.kernel,
                               time_limit=1000, view_shape=(25, 25), output_channels=TRAIN_CHANNELS,
                               env_offset=lo + 7 * r, total_envs=n * world)
        made["pool"] = e.pool
        st = base_states.copy()
        st[:, 1] ^= np.uint64(r)                # batch r of the rotation: other streams, other levels
        e.set_rng_state(st)
        e.reset()
        return e
    env = make_env(0)
    K, Wm = args.steps, args.warmup
    g = torch.Generator(device=dev)
    g.manual_seed(2 + rank)
    flush = None if args.no_flush else _Flush(torch, dev, args.flush_mode)
    rot = Rotation(torch, dev, env, make_env, n * 2 * H * W)
    actions = torch.randint(0, 9, (K + Wm + rot.R, n), generator=g, device=dev, dtype=torch.uint8)
    lib, bref = env._lib, env._batch_ref
    stream = env._stream()
    # generic kernel: the auto-reset is a second (masked) launch; warp kernels: fused
    fused = bool(lib.sl_warp_kernel_supported(H, W)) and args.kernel != 1
    plan, launches_per_step = env.step_plan()
    if not fused:
        launches_per_step += 1

    def step_of(e, t):
        _lib.check(lib.sl_env_step(e._batch_ref, actions[t].data_ptr(), args.kernel,
                                   _lib.SL_STEP_AUTO_RESET, stream))

    def step_kernels(t):
        step_of(env, t)

    for t in range(Wm):
        step_of(rot.envs[t % rot.R], t)
    barrier()
    # ---- value: K steps back to back, device-timed with one event pair, every step's
    # batch read from HBM (rotation over batches larger than the L2), max over ranks ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.mark()
    step_ms = allmax(rot.timed(torch, step_of, K, Wm))                     # step incl. auto-reset
    total_ms = step_ms * K
    # the round-1 measurement, kept for continuity: ONE batch, the L2 flushed before every
    # step, one event pair per step (each interval then also holds the launch latency of a
    # lone kernel: see launch_floor_ms)
    flushed_ms = None
    if flush is not None and not args.quick:
        flushed_ms = allmax(timed_steps(torch, step_kernels, min(K, 300), flush, Wm))
    dyn = int((env._t["static_goals"] == 2).sum())
    bytes_per_launch = step_bytes(H, W, n, dyn)

    if args.quick:
        def quick_obs(e, t):
            _lib.check(lib.sl_env_step_obs(e._batch_ref, actions[t].data_ptr(),
                                           ctypes.byref(e._obs_specs[0]), args.kernel,
                                           _lib.SL_STEP_AUTO_RESET, stream))
        obs_ms = allmax(rot.timed(torch, quick_obs, min(K, 300), Wm))
        if rank == 0:
            peak, _ = peaks()
            ob = bytes_per_launch + n * (2 * H * W + 25 * 25 * len(TRAIN_CHANNELS))
            emit({"quick": True, "envs": n, "pool": args.pool, "ms_per_step": step_ms,
                  "frac": bytes_per_launch / (step_ms * 1e-3) / 1e9 / peak, "kernel": plan,
                  "obs_ms": obs_ms, "obs_frac": ob / (obs_ms * 1e-3) / 1e9 / peak,
                  "obs_kernel": env.step_plan(obs=True)[0], "batches": rot.R})
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- e2e: public API, host buffers, WALL CLOCK around every call ----
    # (a) step_host: the kernel reads the pinned action buffer and writes reward / done
    #     into pinned host memory itself (zero-copy over PCIe); the call returns when
    #     they are visible to the host;
    # (b) step_physics + explicit cudaMemcpyAsync of the same buffers + stream sync.
    # Before each call the L2 is flushed and the device drained, outside the clock.
    host_actions = np.random.default_rng(5 + rank).integers(0, 9, (K, n)).astype(np.uint8)
    env.enable_host_io()
    timing, parts = np.zeros(2), np.zeros(2)
    for t in range(3):
        env.step_host(host_actions[t], auto_reset=True)
    barrier()
    checksum, e2e_s = 0.0, 0.0
    for t in range(K):
        if flush is not None:
            flush.fill_(t & 0xFF)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        reward_np, done_np = env.step_host(host_actions[t], auto_reset=True, timing=timing)
        checksum += float(reward_np[0]) + float(done_np[-1])   # the host consumes the result
        e2e_s += time.perf_counter() - t0
        parts += timing
    e2e_s = allmax(e2e_s)
    barrier()
    t0 = time.perf_counter()
    for t in range(K):                                            # the same, back to back
        reward_np, done_np = env.step_host(host_actions[t], auto_reset=True)
        checksum += float(reward_np[0]) + float(done_np[-1])
    e2e_b2b_s = allmax(time.perf_counter() - t0)

    # (s) a host session: one resident kernel holds the envs in shared memory; step_host
    #     posts the actions in pinned memory and spins on the result word -- no launch, no
    #     stream synchronisation, and no L2 to flush (the state never leaves the SMs)
    session = None
    try:
        env.begin_host_session()
        for t in range(3):
            env.step_host(host_actions[t])
        sparts = np.zeros(2)
        t0 = time.perf_counter()
        for t in range(K):
            reward_np, done_np = env.step_host(host_actions[t], timing=timing)
            checksum += float(reward_np[0]) + float(done_np[-1])
            sparts += timing
        sess_s = time.perf_counter() - t0
        env.end_host_session()
        sess_s = allmax(sess_s)
        session = {"seconds": sess_s, "post_us": sparts[0] / K, "wait_us": sparts[1] / K}
    except (ValueError, RuntimeError) as exc:      # a batch too large to be resident at once
        env.end_host_session()
        session = {"error": str(exc)}
    barrier()

    # (c) two groups of envs in flight on two streams (enable_host_groups): the host reads
    #     one group's results and submits its next actions while the device steps the other
    pipelined = None
    if n % 16 == 0:
        half = n // 2
        env.enable_host_groups(2)
        barrier()
        for gi in range(2):
            env.host_submit(gi, host_actions[0][gi * half:(gi + 1) * half], auto_reset=True)
        t0 = time.perf_counter()
        for t in range(1, K + 1):
            for gi in range(2):
                reward_np, done_np = env.host_wait(gi)
                checksum += float(reward_np[0]) + float(done_np[-1])
                if t < K:
                    env.host_submit(gi, host_actions[t][gi * half:(gi + 1) * half], auto_reset=True)
        pipelined = allmax(time.perf_counter() - t0)
        env.host_sync_groups()

    pinned_actions = torch.from_numpy(host_actions).pin_memory()
    host_reward = torch.empty(n, dtype=torch.float32).pin_memory()
    host_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    barrier()
    e2e_copy_s = 0.0
    for t in range(K):
        if flush is not None:
            flush.fill_(t & 0xFF)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = env.step_physics(pinned_actions[t], auto_reset=True)   # H2D inside
        host_reward.copy_(out["reward"], non_blocking=True)
        host_done.copy_(out["done"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        checksum += float(host_reward[0]) + float(host_done[-1])
        e2e_copy_s += time.perf_counter() - t0
    e2e_copy_s = allmax(e2e_copy_s)

    # ---- supplementary: the rollout regime, K steps back to back with the state
    # L2-resident (no flush, one event pair); not the headline (timing rules) ----
    l2_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    l2_ev[0].record()
    for t in range(K):
        step_kernels(Wm + t)
    l2_ev[1].record()
    barrier()
    l2_ms = l2_ev[0].elapsed_time(l2_ev[1])

    # ---- supplementary: the same steps captured once in a CUDA graph and replayed
    # (device-resident actions: a policy that lives on the GPU) ----
    try:
        graph_ms, graph_k = graph_replay(torch, env, lib, actions, args.kernel)
        graph = {"ms_per_step": graph_ms, "env_steps_per_s": n / (graph_ms * 1e-3), "steps_per_graph": graph_k,
                 "note": "supplementary: sl_env_step captured %d times in one CUDA graph, replayed 10 "
                         "times, no L2 flush; rank 0" % graph_k}
    except Exception as exc:   # noqa: BLE001
        graph = {"error": repr(exc)}
    barrier()

    # ---- supplementary: the same event pair around an EMPTY kernel of the step's
    # launch shape, after the same flush: the part of ms_per_step that is not kernel ----
    floor_ms = timed_steps(torch, lambda t: _lib.check(lib.sl_launch_floor(bref, stream)),
                           min(K, 200), flush)

    # ---- supplementary: the gym step as a user calls it -- physics + auto-reset +
    # observation in ONE launch (sl_env_step_obs), L2 flushed, device-timed ----
    obs_spec = ctypes.byref(env._obs_specs[0])
    obs_plan, obs_launches = env.step_plan(obs=True)

    def step_obs_of(e, t):
        _lib.check(lib.sl_env_step_obs(e._batch_ref, actions[t].data_ptr(),
                                       ctypes.byref(e._obs_specs[0]), args.kernel,
                                       _lib.SL_STEP_AUTO_RESET, stream))

    def step_obs(t):
        step_obs_of(env, t)
    for t in range(3):
        step_obs(t)
    barrier()
    step_obs_ms = allmax(rot.timed(torch, step_obs_of, min(K, 300), Wm))
    step_obs_flushed_ms = None
    if flush is not None:
        step_obs_flushed_ms = allmax(timed_steps(torch, step_obs, min(K, 200), flush, Wm))
    obs_only_bytes = n * (2 * H * W + 25 * 25 * len(TRAIN_CHANNELS))
    step_obs_bytes = bytes_per_launch + obs_only_bytes

    # ---- supplementary: the PUBLIC step() of the Python class, device-resident actions,
    # K calls queued back to back and one sync: per-step wall clock including the Python
    # layer (argument checks, the action copy, ctypes, output bookkeeping) ----
    Kp = min(K, 300)
    for t in range(3):
        env.step(actions[t])
    barrier()
    t0 = time.perf_counter()
    for t in range(Kp):
        env.step(actions[Wm + t])
    torch.cuda.synchronize()
    public_step_us = 1e6 * allmax(time.perf_counter() - t0) / Kp

    # ---- supplementary: PPO data collection with a stub CNN policy ----
    try:
        rollout = rollout_bench(torch, env) if rank == 0 else None
    except Exception as exc:   # noqa: BLE001
        rollout = {"error": repr(exc)}
    barrier()

    # ---- supplementary: SafeLifeEnv.get_obs alone (what reset() returns) ----
    env.get_obs()
    barrier()
    obs_ms = timed_steps(torch, lambda t: env.get_obs(), min(K, 100), flush)
    obs_bytes = n * (4 * H * W + 25 * 25 * len(TRAIN_CHANNELS))
    if sampler:
        sampler.mark()
    clocks = sampler.stop() if sampler else None

    configs = None
    if not args.no_configs:
        peak_now, _ = peaks()
        configs = other_configs(args, dev, world, rank, peak_now, barrier, allmax)

    if rank == 0:
        peak, peak_src = peaks()
        achieved = bytes_per_launch / (step_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": n * world / (step_ms * 1e-3), "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": {
                "workload": "%d envs/GPU x %dx%d %s procgen pool (%d levels), uniform random "
                            "actions, time_limit 1000, auto-reset" % (n, H, W, args.pool, len(levels)),
                "envs_per_gpu": n, "board": [H, W], "pool": args.pool, "kernel": args.kernel,
                "l2": rot.note(),
                "timing": "one CUDA-event pair on the launch stream around the K back-to-back "
                          "steps (after the warm-up and one untimed launch per batch), max over ranks",
                "parallelism": "dp%d (independent env shards, no collective)" % world,
            },
            "clocks": clocks,
            "e2e": e2e_line(n, world, K, e2e_s, parts, session),
            "e2e_launch_per_step": {
                    "value": n * world * K / e2e_s, "unit": "env-steps/s",
                    "h2d_bytes_per_step": n, "d2h_bytes_per_step": 8 * n,
                    "timing": "wall clock (time.perf_counter) around every step_host call incl. the "
                              "host's read of the result; L2 flushed and device drained before each "
                              "call, outside the clock; max over ranks",
                    "us_per_step": 1e6 * e2e_s / K,
                    "host_us": {"launch": parts[0] / K, "wait": parts[1] / K,
                                "note": "inside sl_env_step_host: validation + launch, then "
                                        "cudaStreamSynchronize"},
                    "api": "BatchedSafeLifeEnv.step_host(auto_reset=True) without a session: one "
                           "launch per step; uint8 actions in pinned "
                           "host memory read by the kernel, float32 reward + uint8 done/success/"
                           "times_up written by the kernel into pinned host memory as one packed "
                           "8-byte record per env (zero-copy over PCIe), stream synchronised, host "
                           "read every step",
                    "back_to_back": {"value": n * world * K / e2e_b2b_s, "us_per_step": 1e6 * e2e_b2b_s / K,
                                     "note": "the same calls with no flush between them (state stays in L2)"},
                    "pipelined": None if pipelined is None else {
                        "value": n * world * K / pipelined, "us_per_step": 1e6 * pipelined / K,
                        "api": "enable_host_groups(2) + host_wait / host_submit: two half batches on "
                               "two streams, the host's turn on one overlapping the device's on the "
                               "other; same pinned buffers and completion words; wall clock over %d "
                               "full-batch steps, no flush (the pipeline never drains)" % K},
                    "memcpy_variant": {
                        "value": n * world * K / e2e_copy_s, "us_per_step": 1e6 * e2e_copy_s / K,
                        "api": "step_physics + cudaMemcpyAsync H2D actions, D2H reward + done, "
                               "stream sync and host read every step; wall clock, flushed"}},
            "gpu_launches": launches_per_step * K,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(plan, n, H, W),
                "kernel": plan,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel_ms": step_ms, "dynamic_goal_envs": dyn,
            },
            "single_step_flushed": {
                "ms_per_step": flushed_ms,
                "env_steps_per_s": None if not flushed_ms else n * world / (flushed_ms * 1e-3),
                "frac": None if not flushed_ms else bytes_per_launch / (flushed_ms * 1e-3) / 1e9 / peak,
                "launch_floor_ms": floor_ms,
                "note": "supplementary (the round-1 headline): ONE batch, 256 MiB L2 flush before "
                        "every step (not timed), one event pair per step; each interval also holds "
                        "the launch latency of a lone kernel -- launch_floor_ms is an EMPTY kernel "
                        "of the same launch shape in the same event pair after the same flush"},
            "launch_floor_ms": floor_ms,
            "step_with_obs": {
                "ms": step_obs_ms, "env_steps_per_s": n * world / (step_obs_ms * 1e-3),
                "kernel": obs_plan, "launches_per_step": obs_launches,
                "bytes_per_launch": step_obs_bytes,
                "achieved_gbs": step_obs_bytes / (step_obs_ms * 1e-3) / 1e9,
                "frac": step_obs_bytes / (step_obs_ms * 1e-3) / 1e9 / peak,
                "traffic": measured_traffic(obs_plan, n, H, W),
                "note": "supplementary: SafeLifeEnv.step as the gym API returns it -- physics, "
                        "auto-reset and the uint8[N,25,25,15] training observation written by one "
                        "launch (sl_env_step_obs); bytes = step bytes + goals read + obs written; "
                        "timed like `value` (rotating batches, one event pair); max over ranks",
                "single_step_flushed_ms": step_obs_flushed_ms},
            "public_step": {
                "us_per_step": public_step_us, "env_steps_per_s": n * world / (public_step_us * 1e-6),
                "note": "supplementary: BatchedSafeLifeEnv.step(actions) -> (obs, reward, done, info) "
                        "called %d times back to back with device-resident actions, wall clock "
                        "incl. the Python layer, one sync at the end, no flush" % Kp},
            "obs": {
                "ms": obs_ms, "bytes_per_launch": obs_bytes,
                "achieved_gbs": obs_bytes / (obs_ms * 1e-3) / 1e9,
                "note": "supplementary: k_get_obs_warp (stand-alone, what reset() returns), "
                        "uint8[N,25,25,15] (read board + goals, write obs), L2 flushed; rank 0"},
            "back_to_back": {
                "ms_per_step": l2_ms / K, "env_steps_per_s": n * K / (l2_ms * 1e-3),
                "note": "supplementary: %d steps queued back to back, no L2 flush (state %s the "
                        "126 MB L2), one event pair; rank 0" % (
                            K, "fits" if n * 4 * H * W < 100e6 else "exceeds")},
            "graph_replay": graph,
            "rollout": rollout,
        }
        if configs is not None:
            line["configs"] = configs
        if not args.no_sweep and world == 1:
            line["sweep"] = sweep(args, levels, dev, peak)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"], _ = cpu_arm(levels, n, min(K, 100), 2, args.cpu_seconds)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    quiet_stdout()
    from safelife_b200.levels import load_levels
    levels = load_levels(os.path.join(POOL_DIR, args.pool + ".npz"))
    if args.impl == "reference":
        run_reference(args, levels)
    else:
        run_ours(args, levels)


if __name__ == "__main__":
    main()
