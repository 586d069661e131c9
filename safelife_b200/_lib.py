"""
Loader for libsafelife_b200.so (the C ABI declared in include/safelife_b200.h).

The library is built in-tree with nvcc for sm_100a (see build()); it is loaded
with plain ctypes.  There is no CPU fallback: if the library is missing it is
compiled, and if no CUDA device is usable the entry points return
SL_ERR_NO_DEVICE, which `check` raises as RuntimeError.
"""
import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libsafelife_b200.so")
_SRC = os.path.join(_PKG, "csrc")

SL_OK, SL_ERR_ARG, SL_ERR_CUDA, SL_ERR_NO_DEVICE = 0, 1, 2, 3
SL_RESET_NEXT_LEVEL, SL_RESET_KEEP_STATS = 1, 2
SL_STEP_AUTO_RESET = 1
SL_ABI_VERSION = 3
SL_POOL_UNIFORM_TABLE, SL_POOL_STATIC_GOALS = 1, 2

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]


def _sources():
    return sorted(os.path.join(_SRC, f) for f in os.listdir(_SRC))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(_ROOT, "include", "safelife_b200.h")]
    return any(os.path.getmtime(p) > built for p in deps)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a csrc/abi.cu -> libsafelife_b200.so"""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsafelife_b200.so (no CPU fallback)")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", LIB_PATH, os.path.join(_SRC, "abi.cu")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


class SlEnvBatch(ctypes.Structure):
    """Mirror of struct SlEnvBatch (include/safelife_b200.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n_envs", "H", "W", "max_exits", "time_limit", "n_levels", "level_stride", "pool_flags")] + \
               [(n, ctypes.c_void_p) for n in (
                   "board", "goals", "state", "reward", "done", "success", "times_up",
                   "pool_board", "pool_goals", "pool_loc", "pool_table", "pool_exits",
                   "pool_init_points", "pool_required", "pool_spawn_thr", "host_out")]


class SlObsSpec(ctypes.Structure):
    """Mirror of struct SlObsSpec (include/safelife_b200.h)."""
    _fields_ = [("obs", ctypes.c_void_p), ("vh", ctypes.c_int32), ("vw", ctypes.c_int32),
                ("n_channels", ctypes.c_int32), ("remove_white_goals", ctypes.c_int32),
                ("channels", ctypes.c_int32 * 32)]


class SlWrapSpec(ctypes.Structure):
    """Mirror of struct SlWrapSpec (include/safelife_b200.h)."""
    _fields_ = [("flags", ctypes.c_int32), ("movement_period", ctypes.c_int32),
                ("movement_bonus", ctypes.c_double), ("movement_table", ctypes.c_void_p),
                ("exit_bonus", ctypes.c_double), ("penalty_coef", ctypes.c_double),
                ("baseline", ctypes.c_void_p), ("wstate", ctypes.c_void_p),
                ("reward", ctypes.c_void_p), ("side_effect", ctypes.c_void_p),
                ("pool_required", ctypes.c_void_p)]


class SlHostSession(ctypes.Structure):
    """Mirror of struct SlHostSession (include/safelife_b200.h)."""
    _fields_ = [("host_actions", ctypes.c_void_p), ("dev_actions", ctypes.c_void_p),
                ("dev_ctrl", ctypes.c_void_p), ("host_ctrl", ctypes.c_void_p),
                ("host_out", ctypes.c_void_p), ("idle_timeout_ms", ctypes.c_uint32), ("seq", ctypes.c_uint32),
                ("n_envs", ctypes.c_int32), ("reserved", ctypes.c_int32)]


SL_SESSION_ENDED = 4
SL_SESSION_CTRL_WORDS = 32 + 32 * 256
SL_WRAP_MOVEMENT, SL_WRAP_MOVEMENT_PENALTY, SL_WRAP_EXIT_BONUS = 1, 2, 4
SL_WRAP_SIDE_EFFECT, SL_WRAP_IGNORE_REWARD_CELLS = 8, 16
WRAP_STATE_WORDS = 12   # int32 words per SlWrapState record (48 bytes)

# numpy mirror of struct SlEnvState (96 bytes per env)
import numpy as _np  # noqa: E402

ENV_STATE_DTYPE = _np.dtype([
    ("rng", "<u8", 4), ("loc", "<i4", 2), ("level_id", "<i4"), ("num_steps", "<i4"),
    ("old_value", "<i4"), ("ep_length", "<i4"), ("ep_reward", "<f4"), ("active", "u1"),
    ("static_goals", "u1"), ("empty_points", "u1"), ("reserved0", "u1"), ("spawn_thr", "<u8"),
    ("init_points", "<i4"), ("required", "<i4"), ("count_points", "<i4"), ("n_exits", "<i4"),
    ("exit0", "<i4"), ("exit1", "<i4"),
])
assert ENV_STATE_DTYPE.itemsize == 96
ENV_STATE_WORDS = 24   # int32 words per record


# name -> argtypes, exactly the prototypes of include/safelife_b200.h
_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
PROTOTYPES = {
    "sl_abi_version": [],
    "sl_last_error": [],
    "sl_device_info": [_vp, _vp, _vp, _vp],
    "sl_warp_kernel_supported": [_i, _i],
    "sl_advance_board": [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp],
    "sl_life_occupancy": [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp],
    "sl_side_effect_occupancy": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp],
    "sl_emd_host": [_vp, _vp, _vp, _i, ctypes.c_double, _vp],
    "sl_emd_host_batch": [_vp, _vp, _vp, _vp, _i, ctypes.c_double, _vp, _i],
    "sl_alive_counts": [_vp, _vp, _vp, _i, _i, _vp],
    "sl_execute_actions": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "sl_wrapped_label": [_vp, _vp, _i, _i, _i, _vp],
    "sl_advance_board_host": [_vp, _vp, _i, _i, _i, _f, _i, _vp],
    "sl_life_occupancy_host": [_vp, _vp, _i, _i, _i, _f, _i, _vp],
    "sl_alive_counts_host": [_vp, _vp, _vp, _i, _i],
    "sl_execute_actions_host": [_vp, _vp, _vp, _i, _i, _i, _i, _i],
    "sl_wrapped_label_host": [_vp, _vp, _i, _i, _i],
    "sl_env_step": [ctypes.POINTER(SlEnvBatch), _vp, _i, _i, _vp],
    "sl_env_step_sync": [ctypes.POINTER(SlEnvBatch), _vp, _i, _i, _vp],
    "sl_env_step_host": [ctypes.POINTER(SlEnvBatch), _vp, _i, _i, _vp, _vp],
    "sl_env_step_submit": [ctypes.POINTER(SlEnvBatch), _vp, _i, _i, _vp, _vp, ctypes.c_uint32],
    "sl_host_wait": [_vp, ctypes.c_uint32, _vp],
    "sl_host_session_alloc": [ctypes.POINTER(SlHostSession), _i, _i],
    "sl_host_session_free": [ctypes.POINTER(SlHostSession)],
    "sl_host_session_timeline": [ctypes.POINTER(SlHostSession), _vp],
    "sl_host_session_begin": [ctypes.POINTER(SlEnvBatch), ctypes.POINTER(SlHostSession), _i, _vp],
    "sl_host_session_step": [ctypes.POINTER(SlHostSession), _vp, _vp],
    "sl_host_session_end": [ctypes.POINTER(SlHostSession), _vp],
    "sl_env_step_obs": [ctypes.POINTER(SlEnvBatch), _vp, ctypes.POINTER(SlObsSpec), _i, _i, _vp],
    "sl_env_step_wrapped": [ctypes.POINTER(SlEnvBatch), _vp, ctypes.POINTER(SlObsSpec),
                            ctypes.POINTER(SlWrapSpec), _i, _i, _vp],
    "sl_wrap_reset": [ctypes.POINTER(SlEnvBatch), ctypes.POINTER(SlWrapSpec), _vp, _vp],
    "sl_env_step_plan": [ctypes.POINTER(SlEnvBatch), _i, _vp, _i, _vp],
    "sl_env_step_obs_plan": [ctypes.POINTER(SlEnvBatch), ctypes.POINTER(SlObsSpec), _i, _vp, _i, _vp],
    "sl_launch_floor": [ctypes.POINTER(SlEnvBatch), _vp],
    "sl_env_reset": [ctypes.POINTER(SlEnvBatch), _vp, _vp, _i, _vp],
    "sl_get_obs": [ctypes.POINTER(SlEnvBatch), _vp, _i, _i, _vp, _i, _i, _vp],
    "sl_get_obs_spec": [ctypes.POINTER(SlEnvBatch), ctypes.POINTER(SlObsSpec), _vp],
    "sl_side_effect_cells": [ctypes.POINTER(SlEnvBatch), _vp, _vp, _vp],
}

_lib = None


def lib():
    """The loaded library (built first if needed)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SAFELIFE_B200_LIB")   # A/B testing of kernel builds
        if not path:
            build()
            path = LIB_PATH
        handle = ctypes.CDLL(path)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_char_p if name == "sl_last_error" else ctypes.c_int
        _lib = handle
    return _lib


def check(status):
    """Turn a non-zero status into the exception the reference would raise."""
    if status == SL_OK:
        return
    msg = lib().sl_last_error().decode("utf-8", "replace")
    if status == SL_ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError("libsafelife_b200: " + msg)
