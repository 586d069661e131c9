"""
Drop-in mirror of the reference's ``safelife.speedups`` extension module
(safelife/speedups_src/module.c:515-556) on top of libsafelife_b200.

Same names, argument meaning, return conventions and error behaviour:

    advance_board(board, spawn_prob=0.3, n_step=1) -> new uint16 array
    life_occupancy(board, spawn_prob=0.3, n_step=1000) -> int32[H,W,8]
    alive_counts(board, goals) -> int64[8,9]
    execute_actions(board, locations, actions) -> None   (in place)
    wrapped_label(data) -> (int32 labels, num_labels)
    set_bit_generator(bitgen), seed(n)

Every call stages its numpy arguments through the GPU with the ``*_host`` C
entry points; this is plumbing for single boards (BASELINE config 1) and for
diffing against the reference, not the fast path -- that is
``safelife_b200.env.BatchedSafeLifeEnv``.  The ``*_batch`` functions take a
leading batch axis.

The reference draws from a process-global numpy bit generator
(speedups_src/random.c:21-39).  Here the generator handed to
``set_bit_generator`` must be a PCG64 (numpy's default); its (state, inc) is
read before a call that may draw and written back advanced by exactly the
number of doubles the reference would have consumed.
"""
import ctypes

import numpy as np

from . import _lib

_bitgen = None
_MASK64 = (1 << 64) - 1


def set_bit_generator(bitgen):
    """speedups.set_bit_generator (module.c:441-452, random.c:25-39)."""
    global _bitgen
    if bitgen.state["bit_generator"] != "PCG64":
        raise ValueError("safelife_b200 reproduces PCG64 streams only, got %s"
                         % bitgen.state["bit_generator"])
    _bitgen = bitgen


def seed(n):
    """speedups.seed (deprecated in the reference, module.c:428-438)."""
    set_bit_generator(np.random.default_rng(n if n > 0 else None).bit_generator)


def _rng_words(count=1):
    """uint64[count,4]; row 0 is the current global stream."""
    global _bitgen
    if _bitgen is None:
        seed(0)     # the reference seeds lazily from entropy (random.c:77-81)
    st = _bitgen.state["state"]
    s, inc = st["state"], st["inc"]
    words = np.zeros((count, 4), dtype=np.uint64)
    words[0] = [s >> 64, s & _MASK64, inc >> 64, inc & _MASK64]
    return words


def _store_rng(words):
    st = _bitgen.state
    st["state"]["state"] = (int(words[0, 0]) << 64) | int(words[0, 1])
    _bitgen.state = st


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _board(board):
    # PyArray_FROM_OTF(..., NPY_UINT16, IN_ARRAY | FORCECAST), module.c:28-35
    b = np.ascontiguousarray(np.asarray(board).astype(np.uint16, copy=False))
    if b.ndim != 2 or b.size == 0:
        raise ValueError("board must be a non-empty 2-dimensional array")
    return b


def advance_board(board, spawn_prob=0.3, n_step=1):
    """n_step Life-variant steps; returns a new array (module.c:20-49)."""
    b = _board(board)
    out = np.empty_like(b)
    rng = _rng_words()
    _lib.check(_lib.lib().sl_advance_board_host(
        _ptr(b), _ptr(out), 1, b.shape[0], b.shape[1], float(spawn_prob), int(n_step), _ptr(rng)))
    _store_rng(rng)
    return out


def life_occupancy(board, spawn_prob=0.3, n_step=1000):
    """Occupancy counts int32[H,W,8] over n_step steps (module.c:52-81)."""
    b = _board(board)
    counts = np.zeros(b.shape + (8,), dtype=np.int32)
    rng = _rng_words()
    _lib.check(_lib.lib().sl_life_occupancy_host(
        _ptr(b), _ptr(counts), 1, b.shape[0], b.shape[1], float(spawn_prob), int(n_step),
        _ptr(rng)))
    _store_rng(rng)
    return counts


def alive_counts(board, goals):
    """int64[8,9] histogram (module.c:99-132)."""
    b = np.ascontiguousarray(np.asarray(board).astype(np.uint16, copy=False))
    g = np.ascontiguousarray(np.asarray(goals).astype(np.uint16, copy=False))
    if b.size != g.size:
        raise ValueError("Board and goals must have same size.")
    out = np.zeros((8, 9), dtype=np.int64)
    if b.size:
        _lib.check(_lib.lib().sl_alive_counts_host(_ptr(b), _ptr(g), _ptr(out), 1, b.size))
    return out


def execute_actions(board, locations, actions):
    """In-place agent actions (module.c:155-202). Returns None."""
    if not isinstance(board, np.ndarray) or board.ndim != 2:
        raise ValueError("Board should be 2-dimensional.")
    if board.shape[0] < 3 or board.shape[1] < 3:
        raise ValueError("Board must be at least 3x3.")
    acts = np.ascontiguousarray(np.atleast_1d(np.asarray(actions)).astype(np.int64, copy=False))
    locations = np.asarray(locations) if not isinstance(locations, np.ndarray) else locations
    n_agents = locations.size // 2
    if acts.size != n_agents and acts.size != 1:
        raise ValueError("Locations should be shape (n_agent, 2).")
    # NPY_ARRAY_INOUT_ARRAY: operate on the caller's memory when it already is a
    # C-contiguous array of the right dtype, otherwise on a copy that is written
    # back (the reference relies on WRITEBACKIFCOPY for that, module.c:163-189).
    b = np.ascontiguousarray(board, dtype=np.uint16)
    loc = np.ascontiguousarray(locations, dtype=np.int64)
    _lib.check(_lib.lib().sl_execute_actions_host(
        _ptr(b), _ptr(loc), _ptr(acts), 1, b.shape[0], b.shape[1], n_agents,
        1 if acts.size == n_agents else 0))
    if b is not board:
        board[...] = b
    if loc is not locations:
        locations[...] = loc.reshape(locations.shape)
    return None


def wrapped_label(data):
    """(int32 labels, num_labels) with toroidal 8-connectivity (module.c:223-249)."""
    arr = np.array(data, dtype=np.int32, order="C")
    if arr.ndim != 2 or arr.size == 0:
        raise ValueError("data must be a non-empty 2-dimensional array")
    num = np.zeros(1, dtype=np.int32)
    _lib.check(_lib.lib().sl_wrapped_label_host(_ptr(arr), _ptr(num), 1, arr.shape[0], arr.shape[1]))
    return arr, int(num[0])


# ---- batched variants (leading axis n; rng is uint64[n,4], advanced in place) ----

def _boards(boards):
    b = np.ascontiguousarray(boards, dtype=np.uint16)
    if b.ndim != 3 or b.size == 0:
        raise ValueError("boards must be a non-empty [n,H,W] array")
    return b


def _rng_arg(rng, n):
    if rng is None:
        return None, ctypes.c_void_p(0)
    if rng.dtype != np.uint64 or rng.shape != (n, 4) or not rng.flags.c_contiguous:
        raise ValueError("rng must be a C-contiguous uint64[n,4] array")
    return rng, _ptr(rng)


def advance_board_batch(boards, spawn_prob=0.3, n_step=1, rng=None):
    b = _boards(boards)
    out = np.empty_like(b)
    _, rp = _rng_arg(rng, b.shape[0])
    _lib.check(_lib.lib().sl_advance_board_host(
        _ptr(b), _ptr(out), b.shape[0], b.shape[1], b.shape[2], float(spawn_prob), int(n_step), rp))
    return out


def life_occupancy_batch(boards, spawn_prob=0.3, n_step=1000, rng=None):
    b = _boards(boards)
    counts = np.zeros(b.shape + (8,), dtype=np.int32)
    _, rp = _rng_arg(rng, b.shape[0])
    _lib.check(_lib.lib().sl_life_occupancy_host(
        _ptr(b), _ptr(counts), b.shape[0], b.shape[1], b.shape[2], float(spawn_prob),
        int(n_step), rp))
    return counts


def alive_counts_batch(boards, goals):
    b = _boards(boards)
    g = np.ascontiguousarray(goals, dtype=np.uint16)
    if g.shape != b.shape:
        raise ValueError("Board and goals must have same size.")
    out = np.zeros((b.shape[0], 8, 9), dtype=np.int64)
    _lib.check(_lib.lib().sl_alive_counts_host(_ptr(b), _ptr(g), _ptr(out), b.shape[0],
                                                b.shape[1] * b.shape[2]))
    return out


def execute_actions_batch(boards, locations, actions):
    """boards uint16[n,H,W], locations int64[n,A,2], actions int64[n,A] or [n,1]; in place."""
    for a, dt in ((boards, np.uint16), (locations, np.int64)):
        if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
            raise ValueError("boards / locations must be C-contiguous uint16 / int64 arrays")
    if boards.ndim != 3 or locations.ndim != 3 or locations.shape[0] != boards.shape[0]:
        raise ValueError("Locations should be shape (n, n_agent, 2).")
    acts = np.ascontiguousarray(actions, dtype=np.int64)
    n, A = locations.shape[0], locations.shape[1]
    if acts.shape not in ((n, A), (n, 1)):
        raise ValueError("Locations should be shape (n_agent, 2).")
    _lib.check(_lib.lib().sl_execute_actions_host(
        _ptr(boards), _ptr(locations), _ptr(acts), n, boards.shape[1], boards.shape[2], A,
        1 if acts.shape[1] == A else 0))


def wrapped_label_batch(data):
    arr = np.array(data, dtype=np.int32, order="C")
    if arr.ndim != 3 or arr.size == 0:
        raise ValueError("data must be a non-empty [n,H,W] array")
    num = np.zeros(arr.shape[0], dtype=np.int32)
    _lib.check(_lib.lib().sl_wrapped_label_host(_ptr(arr), _ptr(num), arr.shape[0], arr.shape[1],
                                                 arr.shape[2]))
    return arr, num
