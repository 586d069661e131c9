"""
Import the UNMODIFIED Python reference from /root/reference in this container,
on top of the reference C extension compiled into oracle/_ref/ by
oracle/Makefile.  Used only by the fixture generators under tests/golden/ and
by tests that are skipped when the reference checkout is absent (it never
travels to the GPU box).

TEST INFRASTRUCTURE ONLY -- nothing under safelife_b200/ imports this.

Shims / hazard fixes, applied at run time without
touching the reference tree:
  * `gym` (absent here): Env, Wrapper with attribute forwarding, spaces, register.
  * `pyemd` (absent): an exact transport LP on scipy.optimize.linprog with the
    extra-mass penalty of pyemd 0.5.1 (parity unpinned at that boundary).
  * agent_locs made C-contiguous after every deserialize, because numpy >= 2
    drops the write-back the reference binding relies on (module.c:163-189).
  * every speedups call that may draw happens under safelife.random.set_rng.
"""
import importlib
import os
import sys
import types

import numpy as np

_REF_SO_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# The reference checkout, or -- on the GPU box, where it does not exist -- the verbatim
# copy of its Python package that __graft_entry__.build() leaves in the git-ignored
# baseline/_ref/ (used only by bench.py's cpu_baseline: BASELINE.md section 3.2).
_REF_COPY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REF_ROOT = os.environ.get("SAFELIFE_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "safelife")) and os.path.isdir(os.path.join(_REF_COPY, "safelife")):
    REF_ROOT = _REF_COPY


def have_ref_so():
    return os.path.isdir(_REF_SO_DIR) and any(
        f.startswith("speedups") and f.endswith(".so") for f in os.listdir(_REF_SO_DIR))


def have_reference():
    return have_ref_so() and os.path.isdir(os.path.join(REF_ROOT, "safelife"))


def load_ref_speedups():
    """The reference C extension alone (no reference Python needed)."""
    if "safelife.speedups" in sys.modules:
        return sys.modules["safelife.speedups"]
    if not have_ref_so():
        raise ImportError("oracle/_ref/speedups*.so not built (run `make -C oracle`)")
    so = [f for f in os.listdir(_REF_SO_DIR) if f.startswith("speedups") and f.endswith(".so")][0]
    spec = importlib.util.spec_from_file_location("speedups", os.path.join(_REF_SO_DIR, so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["safelife.speedups"] = mod
    return mod


def emd_lp(a, b, dist, extra_mass_penalty=-1.0):
    """Exact EMD with pyemd's convention: min-cost flow of min(sum a, sum b)
    mass plus extra_mass_penalty * |sum a - sum b| (max(dist) if negative)."""
    from scipy.optimize import linprog
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    n = len(a)
    if extra_mass_penalty < 0:
        extra_mass_penalty = dist.max()
    total = min(a.sum(), b.sum())
    if total <= 0:
        return extra_mass_penalty * abs(a.sum() - b.sum())
    A_ub = np.zeros((2 * n, n * n))
    for i in range(n):
        A_ub[i, i * n:(i + 1) * n] = 1
        A_ub[n + i, i::n] = 1
    res = linprog(dist.reshape(-1), A_ub=A_ub, b_ub=np.concatenate([a, b]),
                  A_eq=np.ones((1, n * n)), b_eq=[total], bounds=(0, None), method="highs")
    return float(res.fun + extra_mass_penalty * abs(a.sum() - b.sum()))


def _install_shims():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env(object):
            pass

        class Wrapper(Env):
            def __init__(self, env):
                self.env = env

            def __getattr__(self, name):
                if name.startswith("_"):
                    raise AttributeError(name)
                return getattr(self.env, name)

        spaces = types.ModuleType("gym.spaces")

        class Discrete(object):
            def __init__(self, n):
                self.n = n

        class Box(object):
            def __init__(self, low, high, shape, dtype):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        spaces.Discrete, spaces.Box = Discrete, Box
        gym.Env, gym.Wrapper, gym.spaces = Env, Wrapper, spaces
        gym.register = lambda **kw: None
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "pyemd" not in sys.modules:
        pyemd = types.ModuleType("pyemd")
        pyemd.emd = emd_lp
        sys.modules["pyemd"] = pyemd


_loaded = None


def load_reference():
    """Returns the imported `safelife` package (reference Python + _ref C)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not have_reference():
        raise ImportError("reference checkout or oracle/_ref not available")
    load_ref_speedups()
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import safelife
    from safelife import safelife_game

    # Hazard 1: keep agent_locs writable in place by the C extension.
    orig = safelife_game.GameState.deserialize

    def deserialize(self, data, as_initial_state=True):
        orig(self, data, as_initial_state)
        self.agent_locs = np.ascontiguousarray(self.agent_locs, dtype=np.int64)

    # GameWithGoals.deserialize calls super().deserialize -> patched base.
    safelife_game.GameState.deserialize = deserialize
    importlib.import_module("safelife.safelife_env")
    importlib.import_module("safelife.level_iterator")
    importlib.import_module("safelife.env_wrappers")
    _loaded = safelife
    return safelife
