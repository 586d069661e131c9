"""
BatchedSafeLifeEnv -- the reference's ``SafeLifeEnv`` gym API
(safelife/safelife_env.py:13-246) over a batch of N single-agent environments
whose whole state lives on one B200 and is stepped by libsafelife_b200.

    env = BatchedSafeLifeEnv(levels, num_envs=4096, view_shape=(25, 25))
    obs = env.reset()                              # uint8 [N, vh, vw, C] (cuda)
    obs, reward, done, info = env.step(actions)    # reward f32 [N], done bool [N]

Same keyword arguments as the reference (safelife_env.py:64-73): single_agent,
time_limit, remove_white_goals, view_shape, output_channels,
side_effect_weights, should_calculate_side_effects; unknown keys raise
ValueError like ``load_kwargs`` (helper_utils.py:78-87).  ``action_space`` is
Discrete(9) and ``observation_space`` a Box of the per-env observation shape.

Differences that come with batching, all deliberate:
  * `levels` is a list of level dicts / a path (see safelife_b200.levels), held
    on the device as a pool of L levels; env e starts on level
    (env_offset + e) % L and, when an episode ends, moves on by `level_stride`
    (mod L) -- the batched form of ``BaseAlgo.act_on_envs`` calling
    ``env.reset()`` on done (training/base_algo.py:231-236) with a level
    iterator behind it.  The stride is the smallest number >= total_envs % L
    that is coprime to L, so every env walks through the whole pool whatever
    the batch size (a stride of N % L would be 0 for 4096 envs on 1024 levels).
    `update_pool` swaps fresh levels into a live env (level_iterator.py:200-266
    streams new procgen levels into every reset).  With ``auto_reset`` the
    returned `obs` of a finished env is the first observation of its next
    episode, as in that loop.
  * outputs are double-buffered: the tensors returned by step t (obs, reward,
    done, info['times_up'], info['episode']['success']) stay valid through
    step t+1 and are overwritten by step t+2; ``step(..., obs_out=buf)`` writes
    the observation into caller storage (a rollout buffer) instead.
    info['board'/'goals'/'agent_locs'] and info['episode']['length'/'reward']
    are live views of the device state, like ``info['board'] is game.board`` in
    the reference.
  * each env owns a PCG64 stream (uint64[4]); the reference swaps a global
    generator in and out per game (safelife/random.py:19-32).
  * `info` holds batched tensors: board, goals, agent_locs, times_up, and
    episode = {length, reward, success} (safelife_env.py:177-201).  After an
    auto-reset the episode totals stay readable until that env's next step.
  * ``should_calculate_side_effects`` defaults to False (the reference: True,
    safelife_env.py:73).  Scoring needs the final board of a finished episode,
    so with it on every step ends in a host sync (``done.any()``) and finished
    envs are reset in a second launch; callers opt in.  When on,
    ``info['episode']['side_effects']`` maps env index -> the reference's score
    dictionary (string keys, 'total' with ``side_effect_weights``) on the step
    the episode ends (safelife_env.py:176-187).
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from .levels import LevelPool, load_levels


class Discrete(object):
    def __init__(self, n):
        self.n = n


class Box(object):
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


def _as_i16(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint16).view(np.int16))


def _as_i64(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64))


def seed_rng_states(seed, n):
    """
    PCG64 states of ``default_rng(s)`` for s in SeedSequence(seed).spawn(n) --
    how the reference seeds one game per level (level_iterator.py:218,
    safelife_game.py:175-180).  uint64[n,4] = state_hi, state_lo, inc_hi, inc_lo.
    """
    ss = seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(seed)
    out = np.zeros((n, 4), dtype=np.uint64)
    m = (1 << 64) - 1
    for i, child in enumerate(ss.spawn(n)):
        st = np.random.PCG64(child).state["state"]
        out[i] = [st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m]
    return out


def coprime_stride(n_envs, n_levels):
    """
    How far an env moves through the pool when its episode ends: the smallest
    s >= n_envs % n_levels (at least 1) with gcd(s, n_levels) = 1, so that env e
    visits every level -- (e + k * s) % L is a bijection of k mod L.
    """
    if n_levels <= 1:
        return 0
    s = n_envs % n_levels or 1
    while math.gcd(s, n_levels) != 1:
        s += 1
    return s % n_levels


def _address(arr):
    """Address of a numpy array's first byte (`arr.ctypes.data` costs 2.5 us, this 0.9)."""
    try:
        return ctypes.addressof(ctypes.c_char.from_buffer(arr))
    except (TypeError, ValueError):          # read-only or empty buffer
        return arr.ctypes.data


class _TensorTable(dict):
    """The env's device tensors by name; per-env scalars are views into `state`."""

    def __init__(self, owned, views):
        super().__init__(owned)
        self.views = views

    def __missing__(self, key):
        return self.views[key]


class GameView(object):
    """Host copy of one env with SafeLifeGame's attribute names (safelife_game.py:126-160)."""


class BatchedSafeLifeEnv(object):
    # defaults as in the reference (safelife_env.py:64-73)
    single_agent = True
    time_limit = 1000
    remove_white_goals = True
    view_shape = (15, 15)
    output_channels = tuple(range(16)) + (25, 26, 27)
    side_effect_weights = None
    should_calculate_side_effects = False
    side_effect_samples = 1000     # side_effect_score's num_samples default (side_effects.py:60)
    # keep the goal planes out of the step kernel's shared memory (SL_POOL_STATIC_GOALS):
    # None = if the pool's goals are static, True / False = force (results never change)
    goals_in_place = None

    def __init__(self, levels, num_envs, device=None, seed=None, auto_reset=True,
                 min_performance_fraction=1.0, kernel=0, env_offset=0, total_envs=None,
                 level_stride=None, **kwargs):
        for key, val in kwargs.items():
            if key.startswith("_") or not hasattr(type(self), key):
                raise ValueError("Unrecognized parameter: '%s'" % (key,))
            setattr(self, key, val)
        if not self.single_agent:
            raise ValueError("BatchedSafeLifeEnv supports single_agent=True only")
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedSafeLifeEnv needs a CUDA device: there is no CPU fallback")
        self._lib = _lib.lib()
        self.device = torch.device(device if device is not None else
                                   "cuda:%d" % torch.cuda.current_device())
        if isinstance(levels, str):
            levels = load_levels(levels)
        self.pool = levels if isinstance(levels, LevelPool) else LevelPool(
            levels, min_performance_fraction)
        self.num_envs = int(num_envs)
        self.auto_reset = bool(auto_reset)
        self.kernel = int(kernel)
        self.view_shape = tuple(int(v) for v in self.view_shape)
        self.action_space = Discrete(9)
        if self.output_channels is None:
            self.observation_space = Box(0, 2 ** 15, self.view_shape, np.uint32)
        else:
            self.output_channels = tuple(int(c) for c in self.output_channels)
            if len(self.output_channels) > 32 or any(c < 0 or c > 31 for c in self.output_channels):
                raise ValueError("output_channels must be at most 32 bit positions of the uint32 view")
            self.observation_space = Box(0, 1, self.view_shape + (len(self.output_channels),),
                                         np.uint8)
        N, H, W, dev = self.num_envs, self.pool.H, self.pool.W, self.device
        with torch.cuda.device(dev):
            p = self.pool
            self._pool_t = {
                "pool_board": _as_i16(p.board).to(dev), "pool_goals": _as_i16(p.goals).to(dev),
                "pool_loc": torch.from_numpy(p.loc).to(dev),
                "pool_table": torch.from_numpy(p.table).to(dev),
                "pool_exits": torch.from_numpy(p.exits).to(dev),
                "pool_init_points": torch.from_numpy(p.init_points).to(dev),
                "pool_required": torch.from_numpy(p.required).to(dev),
                "pool_spawn_thr": _as_i64(p.spawn_thr).to(dev),
            }
            z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
            # struct SlEnvState[N] as int32 words; the named entries below are views into it
            state = z((N, _lib.ENV_STATE_WORDS), torch.int32)
            flags = state[:, 15:16].view(torch.uint8)
            self._t = {
                "board": z((N, H, W), torch.int16), "goals": z((N, H, W), torch.int16),
                "state": state,
                "reward": z((N,), torch.float32), "done": z((N,), torch.uint8),
                "success": z((N,), torch.uint8), "times_up": z((N,), torch.uint8),
            }
            self._views = {
                "rng": state[:, 0:8].view(torch.int64), "loc": state[:, 8:10],
                "level_id": state[:, 10], "num_steps": state[:, 11], "old_value": state[:, 12],
                "ep_length": state[:, 13], "ep_reward": state[:, 14:15].view(torch.float32)[:, 0],
                "active": flags[:, 0], "static_goals": flags[:, 1],
                "init_points": state[:, 18], "required": state[:, 19],
                "count_points": state[:, 20],
            }
            self.env_offset = int(env_offset)
            self.total_envs = int(total_envs) if total_envs is not None else N
            self.level_stride = (coprime_stride(self.total_envs, len(p)) if level_stride is None
                                 else int(level_stride) % len(p))
            first = (torch.arange(N, dtype=torch.int64) + self.env_offset) % len(p)
            self._views["level_id"].copy_(first.to(torch.int32))
            self._views["rng"].copy_(_as_i64(seed_rng_states(seed, N)))
            obs_shape = (N,) + self.view_shape
            if self.output_channels is None:
                obs_dtype = torch.int32
            else:
                obs_shape, obs_dtype = obs_shape + (len(self.output_channels),), torch.uint8
            # per-step outputs, double-buffered: step t writes set t & 1; observations
            # have a ring of their own (step, reset and get_obs all produce one)
            names = ("reward", "done", "success", "times_up")
            self._out = [{k: self._t[k] for k in names},
                         {k: torch.zeros_like(self._t[k]) for k in names}]
            self._obs_bufs = [z(obs_shape, obs_dtype), z(obs_shape, obs_dtype)]
            # the views step() hands out, made once (a .view() costs more than the launch)
            self._out_views = [{k: o[k].view(torch.bool) for k in ("done", "success", "times_up")}
                               for o in self._out]
            self._board_u16 = self._t["board"].view(torch.uint16)
            self._goals_u16 = self._t["goals"].view(torch.uint16)
            self._agent_locs = self._views["loc"].reshape(N, 1, 2)
            self._actions = z((N,), torch.uint8)
        flags = self._pool_flags()
        self._batches, self._batch_refs, self._obs_specs = [], [], []
        for out in self._out:
            ptrs = {k: t.data_ptr() for k, t in self._t.items()}
            ptrs.update({k: out[k].data_ptr() for k in ("reward", "done", "success", "times_up")})
            b = _lib.SlEnvBatch(
                n_envs=N, H=H, W=W, max_exits=p.max_exits, time_limit=int(self.time_limit),
                n_levels=len(p), level_stride=self.level_stride, pool_flags=flags, **ptrs,
                **{k: t.data_ptr() for k, t in self._pool_t.items()})
            self._batches.append(b)
            self._batch_refs.append(ctypes.byref(b))
        self._obs_specs = [self._make_obs_spec(o) for o in self._obs_bufs]
        self._flip = self._obs_i = 0
        self._wrap = None       # wrappers.WrapRuntime once a training wrapper is stacked on top
        self._batch, self._batch_ref = self._batches[0], self._batch_refs[0]
        self._has_reset = False
        # `_t` also answers for the named views of the state record
        self._t = _TensorTable(self._t, self._views)

    def _make_obs_spec(self, obs):
        spec = _lib.SlObsSpec(obs=obs.data_ptr(), vh=self.view_shape[0], vw=self.view_shape[1],
                              n_channels=0 if self.output_channels is None else len(self.output_channels),
                              remove_white_goals=int(bool(self.remove_white_goals)))
        for k, c in enumerate(self.output_channels or ()):
            spec.channels[k] = c
        return spec

    def _next_outputs(self):
        """Flip to the other output set; returns (batch ref, outputs)."""
        self._flip ^= 1
        i = self._flip
        self._batch, self._batch_ref = self._batches[i], self._batch_refs[i]
        out = self._out[i]
        for k, v in out.items():
            dict.__setitem__(self._t, k, v)
        return self._batch_ref, out

    def _next_obs(self, out=None):
        """(obs spec, obs tensor) of the next observation: ring buffer or caller storage."""
        if out is not None:
            return self._make_obs_spec(self._check_obs_out(out)), out
        self._obs_i ^= 1
        return self._obs_specs[self._obs_i], self._obs_bufs[self._obs_i]

    def _pool_flags(self):
        """
        SL_POOL_* promises (include/safelife_b200.h) that hold for this pool: one
        shared points table; goals that are a fixed point of advance_board with no
        spawner (one stencil launch over the pool's goal planes decides).
        """
        p, flags = self.pool, 0
        if (p.table == p.table[0]).all():
            flags |= _lib.SL_POOL_UNIFORM_TABLE
        if self.goals_in_place is None:
            static = self._goals_static(self._pool_t["pool_goals"])
        else:
            static = bool(self.goals_in_place)
        if static:
            flags |= _lib.SL_POOL_STATIC_GOALS
        return flags

    def _goals_static(self, goals):
        """True if every goals plane is a fixed point of advance_board and has no spawner."""
        goals = goals.contiguous()
        nxt = torch.empty_like(goals)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sl_advance_board(
                goals.data_ptr(), nxt.data_ptr(), goals.shape[0], self.pool.H, self.pool.W, 0.0, 1,
                None, self._stream()))
        return bool((nxt == goals).all()) and not bool((goals & 128).any())

    # ------------------------------------------------------------ tensors
    @property
    def board(self):
        """uint16 [N,H,W] view of the device boards (SafeLifeGame.board layout)."""
        self.end_host_session()
        return self._board_u16

    @property
    def goals(self):
        self.end_host_session()
        return self._goals_u16

    @property
    def agent_locs(self):
        """int32 [N,1,2] (row, col), as SafeLifeGame.agent_locs."""
        return self._agent_locs

    @property
    def rng_state(self):
        """uint64[N,4] PCG64 words as a numpy array (host copy)."""
        self.end_host_session()
        return self._t["rng"].contiguous().cpu().numpy().view(np.uint64)

    def set_rng_state(self, states):
        self._t["rng"].copy_(_as_i64(states))

    def _stream(self):
        # every launch path comes through here: an open host session is closed first
        # (its kernel holds the envs in shared memory and the GPU with them)
        if getattr(self, "_session_open", False):
            self.end_host_session()
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -------------------------------------------------------------- gym API
    def reset(self, levels=None):
        """
        SafeLifeEnv.reset for every env (safelife_env.py:203-218).  `levels`
        optionally assigns a pool index per env (int array [N]).
        """
        with torch.cuda.device(self.device):
            new_level = None
            if levels is not None:
                new_level = torch.as_tensor(np.asarray(levels, dtype=np.int32)).to(self.device)
                if new_level.shape != (self.num_envs,):
                    raise ValueError("levels must have shape (num_envs,)")
                if int(new_level.min()) < 0 or int(new_level.max()) >= len(self.pool):
                    raise ValueError("level index out of range")
            _lib.check(self._lib.sl_env_reset(
                self._batch_ref, None, new_level.data_ptr() if new_level is not None else None,
                0, self._stream()))
            self._wrap_reset(None)
            self._has_reset = True
            self._scored = None
            return self.get_obs()

    def _view_of(self, obs):
        return obs if self.output_channels is not None else obs.view(torch.uint32)

    def get_obs(self, out=None):
        """
        SafeLifeEnv.get_obs (safelife_env.py:105-146) for all envs; written into
        `out` if given, else into the next buffer of the output ring.
        """
        spec, out = self._next_obs(out)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sl_get_obs_spec(self._batch_ref, ctypes.byref(spec), self._stream()))
        return self._view_of(out)

    def _check_obs_out(self, out):
        want = self._obs_bufs[0]
        if out.shape != want.shape or out.device != want.device or not out.is_contiguous() or \
                out.element_size() != want.element_size():
            raise ValueError("obs_out must be a contiguous %s tensor of shape %s on %s" % (
                want.dtype, tuple(want.shape), want.device))
        return out

    def step_physics(self, actions, auto_reset=False, obs_out=False):
        """
        The step path (safelife_env.py:148-175), plus the reset of finished envs in
        the same launch when `auto_reset` is set.  obs_out=False: no observation;
        None: the observation is written by the same launch into the output ring;
        a tensor: into that tensor.  Returns the output set of this step.
        """
        if not torch.is_tensor(actions):
            a = np.asarray(actions)
            if a.shape != (self.num_envs,):
                raise ValueError("actions must have shape (num_envs,)")
            if a.size and (a.min() < 0 or a.max() > 8):
                raise ValueError("actions must be in [0, 8]")
            actions = torch.from_numpy(a.astype(np.uint8))
        if actions.shape != (self.num_envs,):
            raise ValueError("actions must have shape (num_envs,)")
        if (actions.dtype == torch.uint8 and actions.is_cuda and actions.device == self.device
                and actions.is_contiguous()):
            # read in place by the kernel, in stream order like any torch op would (a
            # staging copy is a launch of its own and costs more host time than the step)
            aptr, self._held_actions = actions.data_ptr(), actions
        else:
            self._actions.copy_(actions, non_blocking=True)
            aptr = self._actions.data_ptr()
        bref, out = self._next_outputs()
        spec, obs = (None, None) if obs_out is False else self._next_obs(obs_out)
        spec = None if spec is None else ctypes.byref(spec)
        flags = _lib.SL_STEP_AUTO_RESET if auto_reset else 0
        if torch.cuda.current_device() == self.device.index:
            self._launch_step(bref, aptr, spec, flags)
        else:
            with torch.cuda.device(self.device):
                self._launch_step(bref, aptr, spec, flags)
        return dict(out, obs=obs)

    def _launch_step(self, bref, aptr, spec, flags):
        if self._wrap is None:
            status = self._lib.sl_env_step_obs(bref, aptr, spec, self.kernel, flags, self._stream())
        else:   # the training wrappers run as an epilogue of the same launch
            self._wrap.before_step()
            status = self._lib.sl_env_step_wrapped(bref, aptr, spec, self._wrap.fill_spec(self._flip),
                                                   self.kernel, flags, self._stream())
        if status:
            _lib.check(status)

    def _wrap_reset(self, mask):
        if self._wrap is not None:
            _lib.check(self._lib.sl_wrap_reset(
                self._batch_ref, self._wrap.fill_spec(self._flip),
                None if mask is None else mask.data_ptr(), self._stream()))

    def step_plan(self, obs=False):
        """
        (description, launches per step) of what sl_env_step (obs=False) or
        sl_env_step_obs (obs=True: step() with its observation) runs for this batch.
        """
        buf = ctypes.create_string_buffer(320)
        n = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sl_env_step_obs_plan(
                self._batch_ref, ctypes.byref(self._obs_specs[0]) if obs else None, self.kernel,
                buf, 320, ctypes.byref(n)))
        return buf.value.decode(), int(n.value)

    def enable_host_io(self):
        """
        Allocate pinned host buffers that the step kernel reads / writes in place
        (zero-copy over PCIe: pinned memory is device-addressable under UVA), for
        callers whose policy lives on the host: see step_host().  The outputs cross
        PCIe packed, 8 bytes per env (SlEnvBatch.host_out): the returned reward / done /
        success / times_up arrays are strided numpy views of that one buffer.
        """
        n = self.num_envs
        self._host = {
            "actions": torch.zeros(n, dtype=torch.uint8).pin_memory(),
            "out": torch.zeros(n, dtype=torch.int64).pin_memory(),
        }
        fields = {name: getattr(self._batches[0], name) for name, _ in _lib.SlEnvBatch._fields_}
        fields["host_out"] = self._host["out"].data_ptr()
        self._batch_host = _lib.SlEnvBatch(**fields)
        self._batch_host_ref = ctypes.byref(self._batch_host)
        packed = self._host["out"].numpy()
        self._host_np = {
            "actions": self._host["actions"].numpy(),
            "reward": packed.view(np.float32)[0::2], "done": packed.view(np.uint8)[4::8],
            "success": packed.view(np.uint8)[5::8], "times_up": packed.view(np.uint8)[6::8],
        }
        self._host_actions_ptr = self._host["actions"].data_ptr()
        self._host_stream = self._stream()
        return self._host_np

    def step_host(self, actions, auto_reset=None, timing=None):
        """
        One step with HOST buffers end to end: `actions` (array-like [N], values
        0..8) are placed in the pinned action buffer, the kernel reads them and
        writes reward / done / success / times_up (packed, 8 bytes per env) straight
        into pinned host memory, and the call returns once they are visible there
        (sl_env_step_host).
        Returns numpy views (reward float32[N], done uint8[N]) that the next call
        overwrites.  `timing`: optional float64[2] numpy array receiving the
        microseconds spent in the launch and in the wait.  Runs on the stream that
        was current when enable_host_io() was called.
        """
        if getattr(self, "_session_open", False):
            return self._session_step(actions, timing)
        if getattr(self, "_host", None) is None:
            self.enable_host_io()
        h = self._host_np
        if actions is not h["actions"]:
            np.copyto(h["actions"], actions, casting="unsafe")
        auto = self.auto_reset if auto_reset is None else auto_reset
        flags = _lib.SL_STEP_AUTO_RESET if auto else 0
        tptr = None if timing is None else _address(timing)
        # host-driven loop: keep the per-step Python overhead out of the way (the
        # device context manager costs more than the kernel launch)
        if torch.cuda.current_device() == self.device.index:
            status = self._lib.sl_env_step_host(self._batch_host_ref, self._host_actions_ptr,
                                                self.kernel, flags, self._host_stream, tptr)
        else:
            with torch.cuda.device(self.device):
                status = self._lib.sl_env_step_host(self._batch_host_ref, self._host_actions_ptr,
                                                    self.kernel, flags, self._host_stream, tptr)
        if status:
            _lib.check(status)
        return h["reward"], h["done"]

    # ---- host sessions: many host-driven steps on one resident kernel -------------
    def begin_host_session(self, auto_reset=None, idle_timeout_ms=200):
        """
        Open a host session (sl_host_session_begin): one kernel stays resident, keeps
        the envs in shared memory and steps them whenever step_host() posts actions,
        so a host-driven step costs a PCIe round trip instead of a kernel launch and a
        stream synchronisation.  Boards up to 32 x 32 and batches that fit the GPU at
        once (about 4700 envs); physics + auto-reset only.  Until end_host_session()
        the env's tensors are stale and only step_host() may be used; other CUDA work
        on the device waits.  The session ends by itself after `idle_timeout_ms`
        without a step and is reopened by the next step_host().
        """
        if getattr(self, "_host", None) is None:
            self.enable_host_io()
        if self._wrap is not None:
            raise ValueError("host sessions step the bare env (no training wrappers)")
        if getattr(self, "_session", None) is None:
            # the mailbox, allocated by the library (the action buffer write-combined)
            self._session = _lib.SlHostSession()
            self._session_ref = ctypes.byref(self._session)
            with torch.cuda.device(self.device):
                _lib.check(self._lib.sl_host_session_alloc(self._session_ref, self.num_envs, 1))
        self._session.idle_timeout_ms = int(idle_timeout_ms)
        auto = self.auto_reset if auto_reset is None else auto_reset
        self._session_flags = _lib.SL_STEP_AUTO_RESET if auto else 0
        self._session_args = (auto_reset, idle_timeout_ms)
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            _lib.check(self._lib.sl_host_session_begin(self._batch_host_ref, self._session_ref,
                                                       self._session_flags, self._host_stream))
        self._session_open = True

    def __del__(self):
        try:
            if getattr(self, "_session", None) is not None:
                self.end_host_session()
                self._lib.sl_host_session_free(self._session_ref)
                self._session = None
        except Exception:   # noqa: BLE001  (interpreter shutdown)
            pass

    def session_timeline(self):
        """globaltimer stamps (ns) the session kernel left for its last step (csrc/session.cuh)."""
        self.end_host_session()
        stamps = np.zeros(8, dtype=np.uint64)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sl_host_session_timeline(self._session_ref, stamps.ctypes.data))
        return stamps.astype(np.int64)

    def end_host_session(self):
        """Close the host session: the envs are written back and every method works again."""
        if getattr(self, "_session_open", False):
            self._session_open = False
            with torch.cuda.device(self.device):
                _lib.check(self._lib.sl_host_session_end(self._session_ref, self._host_stream))

    def _session_step(self, actions, timing):
        a = actions
        if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.flags.c_contiguous):
            a = np.ascontiguousarray(actions, dtype=np.uint8)
        if a.shape != (self.num_envs,):
            raise ValueError("actions must have shape (num_envs,)")
        tptr = None if timing is None else _address(timing)
        aptr = _address(a)
        status = self._lib.sl_host_session_step(self._session_ref, aptr, tptr)
        if status == _lib.SL_SESSION_ENDED:      # idle timeout: the state is back in HBM; reopen
            self.end_host_session()
            self.begin_host_session(*self._session_args)
            status = self._lib.sl_host_session_step(self._session_ref, aptr, tptr)
        if status:
            _lib.check(status)
        h = self._host_np
        return h["reward"], h["done"]

    # ---- host-driven loops with several groups of envs in flight -----------------
    def enable_host_groups(self, groups=2):
        """
        Split the batch into `groups` contiguous groups that step independently, each
        on a stream of its own, for host-side policies: while the host reads the
        results of one group and chooses its next actions, the device steps another
        (the asynchronous mode of vectorised env libraries).  Returns a list of dicts
        of numpy views (actions, reward, done, success, times_up) into pinned host
        memory, one per group; use host_submit(g) / host_wait(g).  Until
        host_sync_groups() the other methods of the env must not be used.
        """
        if getattr(self, "_host", None) is None:
            self.enable_host_io()
        n = self.num_envs
        if groups < 1 or n % (8 * groups) != 0:
            raise ValueError("num_envs must be a multiple of 8 * groups")
        per = n // groups
        cells = self.pool.H * self.pool.W
        self._flags = torch.zeros(16 * groups, dtype=torch.int32).pin_memory()
        self._groups = []
        with torch.cuda.device(self.device):
            for g in range(groups):
                lo = g * per
                fields = {name: getattr(self._batch_host, name) for name, _ in _lib.SlEnvBatch._fields_}
                fields["n_envs"] = per
                fields["board"] += lo * cells * 2
                fields["goals"] += lo * cells * 2
                fields["state"] += lo * 96
                fields["reward"] += lo * 4
                fields["host_out"] += lo * 8
                for k in ("done", "success", "times_up"):
                    fields[k] += lo
                batch = _lib.SlEnvBatch(**fields)
                stream = torch.cuda.Stream(device=self.device)
                self._groups.append({
                    "batch": batch, "ref": ctypes.byref(batch), "stream": stream,
                    "stream_ptr": ctypes.c_void_p(stream.cuda_stream), "seq": 0,
                    "flag": self._flags.data_ptr() + 64 * g,
                    "actions_ptr": self._host_actions_ptr + lo,
                    "views": {k: v[lo:lo + per] for k, v in self._host_np.items()},
                })
                stream.wait_stream(torch.cuda.current_stream(self.device))
        return [g["views"] for g in self._groups]

    def host_submit(self, g, actions=None, auto_reset=None):
        """Launch one step of group g (actions: array-like [N / groups], or already in the view)."""
        grp = self._groups[g]
        if actions is not None:
            np.copyto(grp["views"]["actions"], actions, casting="unsafe")
        auto = self.auto_reset if auto_reset is None else auto_reset
        grp["seq"] = (grp["seq"] + 1) & 0x7fffffff
        status = self._lib.sl_env_step_submit(grp["ref"], grp["actions_ptr"], self.kernel,
                                              _lib.SL_STEP_AUTO_RESET if auto else 0,
                                              grp["stream_ptr"], grp["flag"], grp["seq"])
        if status:
            _lib.check(status)

    def host_wait(self, g):
        """Wait for the step of group g; returns its (reward, done) numpy views."""
        grp = self._groups[g]
        status = self._lib.sl_host_wait(grp["flag"], grp["seq"], grp["stream_ptr"])
        if status:
            _lib.check(status)
        return grp["views"]["reward"], grp["views"]["done"]

    def host_sync_groups(self):
        """Drain every group's stream; afterwards the env can be used as a whole again."""
        for grp in getattr(self, "_groups", []):
            grp["stream"].synchronize()
            torch.cuda.current_stream(self.device).wait_stream(grp["stream"])

    def step(self, actions, obs_out=None):
        """
        SafeLifeEnv.step for every env: ONE launch writes the new state, reward / done
        and the observation (sl_env_step_obs).  Returns (obs, reward f32 [N], done bool
        [N], info); see the module docstring for how long the tensors stay valid.
        """
        assert self._has_reset, "Game state is not initialized."
        # side effects are scored on the final board, so the reset of finished
        # envs waits until after the score (safelife_env.py:176-187)
        score = bool(self.should_calculate_side_effects)
        out = self.step_physics(actions, self.auto_reset and not score, obs_out=obs_out)
        t = self._t
        obs = self._view_of(out["obs"])
        views = self._out_views[self._flip]            # of the output set this step wrote
        reward, done = out["reward"], views["done"]
        info = {
            "times_up": views["times_up"],
            "episode": {"length": t["ep_length"], "reward": t["ep_reward"],
                        "success": views["success"]},
        }
        if score:
            side_effects = self._episode_side_effects()
            if side_effects:
                info["episode"]["side_effects"] = side_effects
            if self.auto_reset:
                obs = self.reset_done(out=obs_out)
        info["board"], info["goals"], info["agent_locs"] = self.board, self.goals, self.agent_locs
        return obs, reward, done, info

    def _episode_side_effects(self):
        """
        {env index: side_effect_score(..., strkeys=True)} for the envs that finished in
        the last step, with the weighted 'total' entry of safelife_env.py:178-184.
        """
        from .side_effects import side_effect_score_batch
        # once per episode (`self.side_effects is None`, safelife_env.py:176): an env that
        # stays finished without a reset is not scored again
        if getattr(self, "_scored", None) is None:
            self._scored = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        todo = self._t["done"].bool() & ~self._scored
        if not bool(todo.any()):
            return {}
        self._scored |= todo
        out = side_effect_score_batch(self, mask=todo, num_samples=int(self.side_effect_samples),
                                      strkeys=True)
        if self.side_effect_weights is not None:
            for scores in out.values():
                total = np.zeros(2)
                for key, weight in self.side_effect_weights.items():
                    total += weight * np.array(scores.get(key, 0))
                scores["total"] = total.tolist()
        return out

    def reset_done(self, mask=None, out=None):
        """
        Reset the envs whose mask byte is set (default: the `done` flags of the last
        step) onto their next level, as BaseAlgo.act_on_envs does after a done
        (training/base_algo.py:231-236).  Episode totals stay readable.  Returns obs.
        """
        mask = self._t["done"] if mask is None else mask.to(torch.uint8)
        if getattr(self, "_scored", None) is not None:
            self._scored &= ~mask.bool()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sl_env_reset(
                self._batch_ref, mask.data_ptr(), None,
                _lib.SL_RESET_NEXT_LEVEL | _lib.SL_RESET_KEEP_STATS, self._stream()))
            self._wrap_reset(mask)
        return self.get_obs(out)

    # ------------------------------------------------------------ level supply
    def levels_in_use(self):
        """Pool slots that some env is playing right now (sorted numpy array; syncs)."""
        return torch.unique(self._t["level_id"]).cpu().numpy()

    def update_pool(self, slots, levels, force=False):
        """
        Replace the pool entries `slots` by `levels` (level dicts, same board shape)
        while the env is live: the batched counterpart of a level iterator that
        hands a fresh procgen level to every reset (safelife_env.py:203-218,
        level_iterator.py:200-266) -- a host worker pool generates levels and
        streams them into the slots the envs will reach next.  Stream-ordered
        with the steps; takes effect from the next reset onto those slots.  Slots
        being played are refused unless `force` (an env keeps its board, but
        levels with more than two exits or their own points table read both from
        the pool during the episode).
        """
        slots = np.asarray(slots, dtype=np.int64).ravel()
        if len(slots) != len(levels):
            raise ValueError("one level per slot")
        if len(slots) == 0:
            return
        L = len(self.pool)
        if slots.min() < 0 or slots.max() >= L or len(np.unique(slots)) != len(slots):
            raise ValueError("slots must be distinct pool indices")
        if not force and np.isin(slots, self.levels_in_use()).any():
            raise ValueError("some of these slots are being played (see levels_in_use)")
        frac = getattr(self.pool, "min_performance_fraction", 1.0)
        new = LevelPool(levels, frac)
        p = self.pool
        if (new.H, new.W) != (p.H, p.W):
            raise ValueError("new levels must have the pool's board shape")
        if new.max_exits > p.max_exits:
            raise ValueError("new levels have %d exits, the pool was built for %d" % (
                new.max_exits, p.max_exits))
        exits = np.full((len(slots), p.max_exits), -1, dtype=np.int32)
        exits[:, :new.max_exits] = new.exits
        host = {"pool_board": new.board, "pool_goals": new.goals, "pool_loc": new.loc,
                "pool_table": new.table, "pool_exits": exits, "pool_init_points": new.init_points,
                "pool_required": new.required, "pool_spawn_thr": new.spawn_thr}
        for name, arr in (("board", new.board), ("goals", new.goals), ("loc", new.loc),
                          ("table", new.table), ("exits", exits), ("init_points", new.init_points),
                          ("required", new.required), ("spawn_thr", new.spawn_thr)):
            getattr(p, name)[slots] = arr
        for i, lv in zip(slots, levels):
            p.levels[int(i)] = lv
        idx = torch.from_numpy(slots).to(self.device)
        with torch.cuda.device(self.device):
            for name, arr in host.items():
                if arr.dtype == np.uint16:
                    src = _as_i16(arr)
                elif arr.dtype == np.uint64:
                    src = _as_i64(arr)
                else:
                    src = torch.from_numpy(np.ascontiguousarray(arr))
                self._pool_t[name].index_copy_(0, idx, src.to(self.device))
            # the promises made for the old pool must still hold (include/safelife_b200.h)
            flags = self._batch.pool_flags
            if not (p.table == p.table[0]).all():
                flags &= ~_lib.SL_POOL_UNIFORM_TABLE
            if flags & _lib.SL_POOL_STATIC_GOALS and not self._goals_static(self._pool_t["pool_goals"][idx]):
                flags &= ~_lib.SL_POOL_STATIC_GOALS
            if self._wrap is not None:
                self._wrap._required_frac = None      # required points are per level
            for b in self._batches:
                b.pool_flags = flags
            if getattr(self, "_batch_host", None) is not None:
                self._batch_host.pool_flags = flags

    def side_effect_cells(self, baseline):
        """
        SimpleSideEffectPenalty's cell count against `baseline` (uint16/int16 [N,H,W]
        device tensor): env_wrappers.py:177-213 with ignore_reward_cells=False.
        """
        out = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sl_side_effect_cells(
                self._batch_ref, baseline.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def render(self, mode="ansi"):
        raise NotImplementedError("rendering is out of scope (SURVEY.md section 2)")

    def close(self):
        pass

    # ------------------------------------------------------------ host views
    def game_view(self, i):
        """Host copy of env i shaped like a SafeLifeGame (for loggers / side_effect_score)."""
        g = GameView()
        lvl = int(self._t["level_id"][i])
        g.board = self.board[i].cpu().numpy()
        g.goals = self.goals[i].cpu().numpy()
        loc = self._t["loc"][i].cpu().numpy()
        g.agent_locs = loc.reshape(1, 2).astype(np.int64) if loc[0] >= 0 else np.empty((0, 2), int)
        g.agent_names = np.array(["agent%i" % k for k in range(len(g.agent_locs))])
        level = self.pool.levels[lvl]
        g.spawn_prob = level["spawn_prob"]
        g.num_steps = int(self._t["num_steps"][i])
        g.points_table = np.asarray(level["points_table"]).reshape(1, 8, 9)
        g.min_performance = level["min_performance"]
        ex = self.pool.exits[lvl]
        ex = ex[ex >= 0]
        g.exit_locs = (ex // self.pool.W, ex % self.pool.W)
        g._init_data = {"board": level["board"], "goals": level["goals"]}
        return g

    def get_state(self):
        """All per-env tensors (clones), e.g. for checkpointing a rollout."""
        return {k: v.clone() for k, v in dict.items(self._t)}

    def set_state(self, state):
        for k, v in state.items():
            self._t[k].copy_(v)
        self._has_reset = True
