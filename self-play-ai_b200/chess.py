"""Chess rules on the device: ctypes wrappers of the spb_chess_* entry points (include/selfplay_b200.h).

Mirror of the reference's chess adapter, src/game/chess.rs: `get_valid_actions` / `get_status` (:150-166),
`get_next_state` (:112-148), `get_encoding` (:176-249), `Policy::get_channel` / `get_action` (:311-493).
A move is uint16 `from | to << 6 | promotion << 12`; squares are rank*8 + file.  No CPU fallback: every batched call
runs a kernel on the engine's device.
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


class ChessRules:
    """Batched chess `State` methods on one engine's device and stream."""

    def __init__(self, engine: E.Engine):
        self.e = engine

    def legal_moves(self, states, history=None):
        """-> (moves[n,256] u16 sorted by (from,to,promotion), counts[n], policy_index[n,256], status[n], repetitions[n])."""
        st = _states(states)
        n = len(st)
        h = _history(history, n)
        moves = np.zeros((n, MAX_MOVES), np.uint16)
        pidx = np.zeros((n, MAX_MOVES), np.uint16)
        counts, reps, status = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint8)
        self.e._chk(self.e._L.spb_chess_legal_moves(self.e._h, st.ctypes.data, None if h is None else h.ctypes.data, n, moves.ctypes.data,
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
        self.e._chk(self.e._L.spb_chess_next_states(self.e._h, st.ctypes.data, h.ctypes.data, mv.ctypes.data, n, out.ctypes.data, err.ctypes.data))
        return out, h, err

    def encode(self, states, history=None) -> np.ndarray:
        st = _states(states)
        n = len(st)
        h = _history(history, n)
        out = np.zeros((n, PLANES, 8, 8), np.float32)
        self.e._chk(self.e._L.spb_chess_encode(self.e._h, st.ctypes.data, None if h is None else h.ctypes.data, n, out.ctypes.data))
        return out

    def perft(self, state, depth: int) -> int:
        st = _states(state)
        assert len(st) == 1
        nodes = C.c_uint64()
        self.e._chk(self.e._L.spb_chess_perft(self.e._h, st.ctypes.data, depth, C.byref(nodes)))
        return int(nodes.value)
