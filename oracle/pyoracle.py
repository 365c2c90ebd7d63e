"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module (see oracle/oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

GAME_TTT, GAME_C4 = 0, 1
EVAL_NET, EVAL_DET, EVAL_UNIFORM = 0, 1, 2
ONGOING, TIED, WON = 0, 1, 2
MAX_ACTIONS = 9
NUM_ACTIONS = {GAME_TTT: 9, GAME_C4: 7}
BOARD = {GAME_TTT: (3, 3), GAME_C4: (6, 7)}


class State(C.Structure):
    """Mirror of spb_state (include/selfplay_b200.h)."""

    _fields_ = [
        ("stones", C.c_uint64 * 2),
        ("current_player", C.c_uint8),
        ("num_actions_played", C.c_uint8),
        ("status", C.c_uint8),
        ("reserved", C.c_uint8 * 5),
    ]

    def key(self):
        return (int(self.stones[0]), int(self.stones[1]), int(self.current_player),
                int(self.num_actions_played), int(self.status))

    def copy(self):
        s = State()
        C.memmove(C.byref(s), C.byref(self), C.sizeof(State))
        return s

    def __repr__(self):
        return "State(x=%#x,o=%#x,p=%d,n=%d,st=%d)" % self.key()


STATE_DTYPE = np.dtype([("stones", "<u8", (2,)), ("current_player", "u1"), ("num_actions_played", "u1"),
                        ("status", "u1"), ("reserved", "u1", (5,))])
assert STATE_DTYPE.itemsize == C.sizeof(State) == 24


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("simulations", "evaluations", "terminal_leaves", "path_length_sum",
                                          "children_created", "nodes_live", "kernel_launches")] + \
               [("reserved", C.c_uint64 * 5)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float))


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("oracle.cc", "oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u32p, u8p, f32p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        sp = C.POINTER(State)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int32, C.c_uint32, C.c_float]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_void_p]
        L.orc_search.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_root_children.argtypes = [C.c_void_p, C.c_uint32, u8p, u32p, u32p, u32p]
        L.orc_root_policy.argtypes = [C.c_void_p, C.c_uint32, f32p]
        L.orc_use_subtree.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_get_state.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, sp]
        L.orc_arena_len.argtypes = [C.c_void_p, C.c_uint32, u32p]
        L.orc_node_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u32p, f32p, f32p, u32p, u32p]
        L.orc_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.orc_set_leaves_per_tree.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_next_state.argtypes = [C.c_int32, sp, C.c_uint8, sp]
        L.orc_valid_actions.restype = C.c_uint32
        L.orc_valid_actions.argtypes = [C.c_int32, sp]
        L.orc_encode.argtypes = [C.c_int32, sp, f32p]
        L.orc_mask_invalid_actions.argtypes = [C.c_int32, sp, f32p, f32p]
        L.orc_det_eval.argtypes = [C.c_int32, sp, f32p, f32p]
        L.orc_det_hash.restype = C.c_uint64
        L.orc_det_hash.argtypes = [C.c_int32, sp]
        L.orc_greedy_game.argtypes = [C.c_int32, sp, C.c_float, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p,
                                      u8p, u32p, u32p, u32p, u8p]
        L.orc_baseline_run.restype = C.c_uint64
        L.orc_baseline_run.argtypes = [C.c_int32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def states_array(states) -> np.ndarray:
    """list[State] | structured ndarray -> contiguous structured ndarray of STATE_DTYPE."""
    if isinstance(states, np.ndarray):
        assert states.dtype == STATE_DTYPE
        return np.ascontiguousarray(states)
    arr = np.zeros(len(states), dtype=STATE_DTYPE)
    for i, s in enumerate(states):
        arr[i]["stones"] = (s.stones[0], s.stones[1])
        arr[i]["current_player"] = s.current_player
        arr[i]["num_actions_played"] = s.num_actions_played
        arr[i]["status"] = s.status
    return arr


def state_from_record(rec) -> State:
    s = State()
    s.stones[0], s.stones[1] = int(rec["stones"][0]), int(rec["stones"][1])
    s.current_player = int(rec["current_player"])
    s.num_actions_played = int(rec["num_actions_played"])
    s.status = int(rec["status"])
    return s


def make_eval_callback(game: int, fn):
    """Wrap fn(enc: np.ndarray[n,3,R,C]) -> (probs[n,A], values[n]) as an orc_eval_fn."""
    R, Cc = BOARD[game]
    A = NUM_ACTIONS[game]

    def _cb(_user, enc_p, n, probs_p, values_p):
        enc = np.ctypeslib.as_array(enc_p, shape=(n, 3, R, Cc))
        probs, values = fn(enc)
        np.ctypeslib.as_array(probs_p, shape=(n, A))[...] = np.asarray(probs, dtype=np.float32).reshape(n, A)
        np.ctypeslib.as_array(values_p, shape=(n,))[...] = np.asarray(values, dtype=np.float32).reshape(n)

    return EVAL_FN(_cb)


# ---- State-trait helpers ---------------------------------------------------------------------
def next_state(game: int, s: State, action: int):
    out = State()
    rc = lib().orc_next_state(game, C.byref(s), action, C.byref(out))
    return out if rc == 0 else None


def valid_actions(game: int, s: State):
    m = lib().orc_valid_actions(game, C.byref(s))
    return [a for a in range(NUM_ACTIONS[game]) if m >> a & 1]


def encode(game: int, s: State) -> np.ndarray:
    R, Cc = BOARD[game]
    out = np.zeros((3, R, Cc), dtype=np.float32)
    lib().orc_encode(game, C.byref(s), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def mask_invalid_actions(game: int, s: State, probs) -> np.ndarray:
    A = NUM_ACTIONS[game]
    p = np.ascontiguousarray(probs, dtype=np.float32)
    out = np.zeros(A, dtype=np.float32)
    lib().orc_mask_invalid_actions(game, C.byref(s), p.ctypes.data_as(C.POINTER(C.c_float)),
                                   out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def det_eval(game: int, s: State):
    A = NUM_ACTIONS[game]
    p = np.zeros(A, dtype=np.float32)
    v = C.c_float()
    lib().orc_det_eval(game, C.byref(s), p.ctypes.data_as(C.POINTER(C.c_float)), C.byref(v))
    return p, v.value


def det_hash(game: int, s: State) -> int:
    return int(lib().orc_det_hash(game, C.byref(s)))


class Forest:
    """`Vec<Tree>` + `Mcts::search` of the reference, restated on the CPU."""

    def __init__(self, game: int, num_trees: int, c: float = 2.0, leaves_per_tree: int = 1):
        self.game, self.n, self.A = game, num_trees, NUM_ACTIONS[game]
        self._h = lib().orc_create(game, num_trees, c)
        if not self._h:
            raise ValueError("orc_create failed")
        if leaves_per_tree != 1 and lib().orc_set_leaves_per_tree(self._h, leaves_per_tree) != 0:
            raise ValueError("leaves_per_tree must be in 1..16")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    __del__ = close

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle error %d: %s" % (rc, lib().orc_last_error(self._h).decode()))

    def reset(self, roots=None, slots=None):
        n = self.n if slots is None else len(slots)
        sl = None if slots is None else (C.c_uint32 * n)(*slots)
        arr = None if roots is None else states_array(roots)
        self._chk(lib().orc_reset(self._h, sl, n, None if arr is None else arr.ctypes.data))

    def search(self, num_searches: int, evaluator: int = EVAL_DET, callback=None):
        self._chk(lib().orc_search(self._h, num_searches, evaluator,
                                   C.cast(callback, C.c_void_p) if callback is not None else None, None))

    def root_children(self, slot: int):
        a = (C.c_uint8 * MAX_ACTIONS)()
        cnt = (C.c_uint32 * MAX_ACTIONS)()
        ids = (C.c_uint32 * MAX_ACTIONS)()
        n = C.c_uint32()
        self._chk(lib().orc_root_children(self._h, slot, a, cnt, ids, C.byref(n)))
        k = n.value
        return list(a[:k]), list(cnt[:k]), list(ids[:k])

    def root_policy(self, slot: int) -> np.ndarray:
        p = np.zeros(self.A, dtype=np.float32)
        self._chk(lib().orc_root_policy(self._h, slot, p.ctypes.data_as(C.POINTER(C.c_float))))
        return p

    def use_subtree(self, slot: int, node_id: int):
        self._chk(lib().orc_use_subtree(self._h, slot, node_id))

    def get_state(self, slot: int, node_id: int) -> State:
        s = State()
        self._chk(lib().orc_get_state(self._h, slot, node_id, C.byref(s)))
        return s

    def arena_len(self, slot: int) -> int:
        n = C.c_uint32()
        self._chk(lib().orc_arena_len(self._h, slot, C.byref(n)))
        return n.value

    def node_stats(self, slot: int, node_id: int):
        n, fc, nc = C.c_uint32(), C.c_uint32(), C.c_uint32()
        w, p = C.c_float(), C.c_float()
        self._chk(lib().orc_node_stats(self._h, slot, node_id, C.byref(n), C.byref(w), C.byref(p), C.byref(fc), C.byref(nc)))
        return dict(visit_count=n.value, value_sum=w.value, prior=p.value, first_child=fc.value, n_children=nc.value)

    def counters(self) -> dict:
        c = Counters()
        self._chk(lib().orc_get_counters(self._h, C.byref(c)))
        return c.as_dict()


def greedy_game(game: int, num_searches: int, evaluator: int = EVAL_DET, root: State | None = None, c: float = 2.0,
                callback=None):
    acts = (C.c_uint8 * 64)()
    sizes = (C.c_uint32 * 64)()
    last = (C.c_uint32 * MAX_ACTIONS)()
    nl = C.c_uint32()
    fs = C.c_uint8()
    n = lib().orc_greedy_game(game, C.byref(root) if root is not None else None, c, num_searches, evaluator,
                              C.cast(callback, C.c_void_p) if callback is not None else None, None,
                              acts, sizes, last, C.byref(nl), C.byref(fs))
    return dict(actions=list(acts[:n]), arena_sizes=list(sizes[:n]), last_counts=list(last[:nl.value]),
                final_status=fs.value)


def baseline_run(game: int, roots, threads: int, games_per_thread: int, num_searches: int,
                 evaluator: int = EVAL_DET, c: float = 2.0, callback=None):
    """Returns (simulations, seconds) — the CPU baseline of bench.py."""
    arr = None if roots is None else states_array(roots)
    if arr is not None:
        assert len(arr) >= threads * games_per_thread
    sec = C.c_double()
    sims = lib().orc_baseline_run(game, None if arr is None else arr.ctypes.data, threads, games_per_thread, c,
                                  num_searches, evaluator,
                                  C.cast(callback, C.c_void_p) if callback is not None else None, None, C.byref(sec))
    return int(sims), float(sec.value)
