"""Chess rules on the device: ctypes wrappers of the spb_chess_* entry points (include/selfplay_b200.h).

Mirror of the reference's chess adapter, src/game/chess.rs: `get_valid_actions` / `get_status` (:150-166),
`get_next_state` (:112-148), `get_encoding` (:176-249), `Policy::get_channel` / `get_action` (:311-493).
A move is uint16 `from | to << 6 | promotion << 12`; squares are rank*8 + file.  `ChessEngine` adds the trees:
`Mcts::search` (mcts.rs:196-332), `Tree::use_subtree` (:161-192), `Model::predict` (model/mod.rs:36-98).  No CPU fallback:
every batched call runs kernels on the engine's device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import engine as E

MAX_MOVES, MAX_HISTORY, PLANES, POLICY_SIZE, NO_SQUARE = 256, 512, 19, 4672, 64
MOVE_NONE = 0xFFFF

CHESS_STATE_DTYPE = np.dtype([("piece", "<u8", (6,)), ("color", "<u8", (2,)), ("side", "u1"), ("castle", "u1"), ("ep", "u1"),
                              ("reserved0", "u1"), ("fifty", "<u2"), ("plies", "<u2"), ("hist_len", "<u4"), ("reserved1", "<u4")])
assert CHESS_STATE_DTYPE.itemsize == 80


def start_position() -> np.ndarray:
    """State::default() (chess.rs:94-102) -> array of one CHESS_STATE_DTYPE record."""
    s = np.zeros(1, CHESS_STATE_DTYPE)
    rc = E.load_library().spb_chess_start_position(s.ctypes.data)
    if rc != 0:
        raise E.EngineError(rc, "spb_chess_start_position")
    return s


def move_channel(side: int, move: int) -> int:
    return E.load_library().spb_chess_move_channel(side, move)


def policy_index(side: int, move: int) -> int:
    return E.load_library().spb_chess_policy_index(side, move)


def action(side: int, channel: int, row: int, col: int) -> int:
    return E.load_library().spb_chess_action(side, channel, row, col)


def _states(states) -> np.ndarray:
    a = np.ascontiguousarray(states, dtype=CHESS_STATE_DTYPE)
    return a.reshape(-1)


def _history(history, n):
    if history is None:
        return None
    h = np.ascontiguousarray(history, dtype=np.uint64)
    assert h.shape == (n, MAX_HISTORY), h.shape
    return h


class ChessEngine:
    """The chess engine: `Mcts<chess Net>` + its trees on one GPU (ref: src/mcts.rs:41-44 over src/game/chess.rs and
    src/model/chess.rs) — batched `State` methods, search, re-rooting, `Model::predict`."""

    def __init__(self, num_games=64, evaluator=E.EVAL_NET, c=2.0, device=0, max_nodes_per_tree=0, flags=0):
        L = E.load_library()
        cfg = E.Config()
        L.spb_default_config(C.byref(cfg))
        cfg.game, cfg.num_games, cfg.evaluator, cfg.c, cfg.device = E.GAME_CHESS, num_games, evaluator, c, device
        cfg.max_nodes_per_tree, cfg.flags, cfg.leaves_per_tree = max_nodes_per_tree, flags, 1
        h = C.c_void_p()
        rc = L.spb_chess_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise E.EngineError(rc, L.spb_chess_last_error(None).decode())
        self._h, self._L, self.G = h, L, num_games

    def close(self):
        if getattr(self, "_h", None):
            self._L.spb_chess_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise E.EngineError(rc, self._L.spb_chess_last_error(self._h).decode())

    # ---- State trait, batched ----------------------------------------------------------------------------------------
    def legal_moves(self, states, history=None):
        """-> (moves[n,256] u16 sorted by (from,to,promotion), counts[n], policy_index[n,256], status[n], repetitions[n])."""
        st = _states(states)
        n = len(st)
        h = _history(history, n)
        moves = np.zeros((n, MAX_MOVES), np.uint16)
        pidx = np.zeros((n, MAX_MOVES), np.uint16)
        counts, reps, status = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint8)
        self._chk(self._L.spb_chess_legal_moves(self._h, st.ctypes.data, None if h is None else h.ctypes.data, n, moves.ctypes.data,
                                                counts.ctypes.data, pidx.ctypes.data, status.ctypes.data, reps.ctypes.data))
        return moves, counts, pidx, status, reps

    def next_states(self, states, history, moves):
        """get_next_state for n states -> (out_states, history (updated copy), err[n])."""
        st = _states(states)
        n = len(st)
        h = np.array(_history(history, n), copy=True)
        mv = np.ascontiguousarray(moves, dtype=np.uint16)
        out = np.zeros(n, CHESS_STATE_DTYPE)
        err = np.zeros(n, np.int32)
        self._chk(self._L.spb_chess_next_states(self._h, st.ctypes.data, h.ctypes.data, mv.ctypes.data, n, out.ctypes.data, err.ctypes.data))
        return out, h, err

    def encode(self, states, history=None) -> np.ndarray:
        st = _states(states)
        n = len(st)
        h = _history(history, n)
        out = np.zeros((n, PLANES, 8, 8), np.float32)
        self._chk(self._L.spb_chess_encode(self._h, st.ctypes.data, None if h is None else h.ctypes.data, n, out.ctypes.data))
        return out

    def perft(self, state, depth: int) -> int:
        st = _states(state)
        assert len(st) == 1
        nodes = C.c_uint64()
        self._chk(self._L.spb_chess_perft(self._h, st.ctypes.data, depth, C.byref(nodes)))
        return int(nodes.value)

    # ---- weights / Model::predict ------------------------------------------------------------------------------------
    def load_weights(self, blob: bytes):
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._chk(self._L.spb_chess_load_weights(self._h, C.cast(buf, C.c_void_p), len(blob)))

    def predict(self, states, history=None, want_logits=False):
        st = _states(states)
        n = len(st)
        h = _history(history, n)
        pol = np.zeros((n, POLICY_SIZE), np.float32)
        val = np.zeros(n, np.float32)
        lg = np.zeros((n, POLICY_SIZE), np.float32) if want_logits else None
        self._chk(self._L.spb_chess_predict(self._h, st.ctypes.data, None if h is None else h.ctypes.data, n, pol.ctypes.data, val.ctypes.data,
                                            None if lg is None else lg.ctypes.data))
        return (pol, val, lg) if want_logits else (pol, val)

    # ---- trees -------------------------------------------------------------------------------------------------------
    def reset_games(self, roots=None, history=None, slots=None):
        n = self.G if slots is None else len(slots)
        sl = None if slots is None else np.ascontiguousarray(slots, dtype=np.uint32)
        st = None if roots is None else _states(roots)
        h = None if history is None else _history(history, n)
        if st is not None:
            assert len(st) == n
        self._chk(self._L.spb_chess_reset_games(self._h, None if sl is None else sl.ctypes.data, n, None if st is None else st.ctypes.data,
                                                None if h is None else h.ctypes.data))

    def search(self, num_searches: int):
        self._chk(self._L.spb_chess_search(self._h, num_searches))

    def last_search_ms(self) -> float:
        ms = C.c_float()
        self._chk(self._L.spb_chess_last_search_ms(self._h, C.byref(ms)))
        return ms.value

    def root_children(self, slot: int):
        mv, cnt, ids = np.zeros(MAX_MOVES, np.uint16), np.zeros(MAX_MOVES, np.uint32), np.zeros(MAX_MOVES, np.uint32)
        n = C.c_uint32()
        self._chk(self._L.spb_chess_root_children(self._h, slot, mv.ctypes.data, cnt.ctypes.data, ids.ctypes.data, C.byref(n)))
        k = n.value
        return mv[:k].tolist(), cnt[:k].tolist(), ids[:k].tolist()

    def root_children_all(self, out=None):
        """`out`: the four arrays of an earlier call (e.g. in pinned host memory), written in place."""
        if out is not None:
            mv, cnt, ids, n = out
            assert mv.shape == (self.G, MAX_MOVES) and mv.dtype == np.uint16 and cnt.dtype == np.uint32 and ids.dtype == np.uint32 and n.shape == (self.G,)
            assert all(x.flags.c_contiguous for x in out)
        else:
            mv = np.zeros((self.G, MAX_MOVES), np.uint16)
            cnt = np.zeros((self.G, MAX_MOVES), np.uint32)
            ids = np.zeros((self.G, MAX_MOVES), np.uint32)
            n = np.zeros(self.G, np.uint32)
        self._chk(self._L.spb_chess_root_children_all(self._h, mv.ctypes.data, cnt.ctypes.data, ids.ctypes.data, n.ctypes.data))
        return mv, cnt, ids, n

    def root_policy(self, slot: int) -> np.ndarray:
        p = np.zeros(POLICY_SIZE, np.float32)
        self._chk(self._L.spb_chess_root_policy(self._h, slot, p.ctypes.data))
        return p

    def advance(self, child_ids, slots=None) -> np.ndarray:
        ids = np.ascontiguousarray(child_ids, dtype=np.uint32)
        sl = None if slots is None else np.ascontiguousarray(slots, dtype=np.uint32)
        out = np.zeros(len(ids), CHESS_STATE_DTYPE)
        self._chk(self._L.spb_chess_advance(self._h, None if sl is None else sl.ctypes.data, ids.ctypes.data, len(ids), out.ctypes.data))
        return out

    def get_state(self, slot: int, node_id: int) -> np.ndarray:
        out = np.zeros(1, CHESS_STATE_DTYPE)
        self._chk(self._L.spb_chess_get_state(self._h, slot, node_id, out.ctypes.data))
        return out[0]

    def arena_len(self, slot: int) -> int:
        n = C.c_uint32()
        self._chk(self._L.spb_chess_arena_len(self._h, slot, C.byref(n)))
        return n.value

    def node_stats(self, slot: int, node_id: int) -> dict:
        n, fc, nc = C.c_uint32(), C.c_uint32(), C.c_uint32()
        w, p = C.c_float(), C.c_float()
        mv, st = C.c_uint16(), C.c_uint8()
        self._chk(self._L.spb_chess_node_stats(self._h, slot, node_id, C.byref(n), C.byref(w), C.byref(p), C.byref(fc), C.byref(nc),
                                               C.byref(mv), C.byref(st)))
        return dict(visit_count=n.value, value_sum=w.value, prior=p.value, first_child=fc.value, n_children=nc.value, move=mv.value,
                    status=st.value)

    def time_conv(self, iters=20):
        """-> (avg launch ms, positions in the batch, FLOPs per launch, FLOPs per evaluated position) of the residual conv kernel
        re-run on the last evaluator batch (spb_chess_time_conv)."""
        ms, n, fl, fp = C.c_float(), C.c_uint32(), C.c_double(), C.c_double()
        self._chk(self._L.spb_chess_time_conv(self._h, iters, C.byref(ms), C.byref(n), C.byref(fl), C.byref(fp)))
        return ms.value, n.value, fl.value, fp.value

    def counters(self) -> dict:
        c = E.Counters()
        self._chk(self._L.spb_chess_get_counters(self._h, C.byref(c)))
        d = c.as_dict()
        d["largest_arena"] = int(c.reserved[0])
        return d

    def reset_counters(self):
        self._chk(self._L.spb_chess_reset_counters(self._h))


def check_weights(blob: bytes):
    """Host-only validation of a chess checkpoint -> (code, message)."""
    L = E.load_library()
    buf = (C.c_char * len(blob)).from_buffer_copy(blob)
    err = C.create_string_buffer(512)
    rc = L.spb_chess_check_weights(C.cast(buf, C.c_void_p), len(blob), err, 512)
    return rc, err.value.decode()
