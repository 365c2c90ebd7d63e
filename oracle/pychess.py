"""ctypes binding of the chess part of the CPU oracle (oracle/chess_oracle.cc).

TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/oracle.h).  PARITY UNPINNED BY THE REFERENCE: src/game/chess.rs delegates
move generation (and the order of the legal moves) to the un-vendored crate chess 3.2.0; the legal-move SETS are pinned by
the published perft counts below, everything chess.rs computes itself is restated from its text.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import pyoracle as O

MAX_MOVES, MAX_HISTORY, PLANES, POLICY_SIZE = 256, 512, 19, 4672
KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
# chessprogramming.org/Perft_Results — the standard move-generator test positions and their leaf counts by depth
PERFT = {
    None: [20, 400, 8902, 197281, 4865609],
    KIWIPETE: [48, 2039, 97862, 4085603],
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1": [14, 191, 2812, 43238, 674624],
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1": [6, 264, 9467, 422333],
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8": [44, 1486, 62379, 2103487],
    "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10": [46, 2079, 89890, 3894594],
}


class ChessState(C.Structure):
    """Mirror of spb_chess_state (include/selfplay_b200.h)."""
    _fields_ = [("piece", C.c_uint64 * 6), ("color", C.c_uint64 * 2), ("side", C.c_uint8), ("castle", C.c_uint8), ("ep", C.c_uint8),
                ("reserved0", C.c_uint8), ("fifty", C.c_uint16), ("plies", C.c_uint16), ("hist_len", C.c_uint32), ("reserved1", C.c_uint32)]


CHESS_STATE_DTYPE = np.dtype([("piece", "<u8", (6,)), ("color", "<u8", (2,)), ("side", "u1"), ("castle", "u1"), ("ep", "u1"),
                              ("reserved0", "u1"), ("fifty", "<u2"), ("plies", "<u2"), ("hist_len", "<u4"), ("reserved1", "<u4")])
assert CHESS_STATE_DTYPE.itemsize == C.sizeof(ChessState) == 80

CHESS_EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float))
_bound = False


def _lib():
    global _bound
    L = O.lib()
    if not _bound:
        vp, u16p, f32p = C.c_void_p, C.POINTER(C.c_uint16), C.POINTER(C.c_float)
        L.orc_chess_new.restype, L.orc_chess_new.argtypes = vp, [C.c_char_p]
        L.orc_chess_free.argtypes = [vp]
        L.orc_chess_clone.restype, L.orc_chess_clone.argtypes = vp, [vp]
        L.orc_chess_legal_moves.restype, L.orc_chess_legal_moves.argtypes = C.c_int32, [vp, u16p]
        L.orc_chess_make_move.restype, L.orc_chess_make_move.argtypes = C.c_int32, [vp, C.c_uint16]
        L.orc_chess_status.restype, L.orc_chess_status.argtypes = C.c_int32, [vp]
        L.orc_chess_repetitions.restype, L.orc_chess_repetitions.argtypes = C.c_int32, [vp]
        L.orc_chess_value.restype, L.orc_chess_value.argtypes = C.c_float, [vp]
        L.orc_chess_side.restype, L.orc_chess_side.argtypes = C.c_int32, [vp]
        L.orc_chess_perft.restype, L.orc_chess_perft.argtypes = C.c_uint64, [vp, C.c_int32]
        L.orc_chess_encode.argtypes = [vp, f32p]
        L.orc_chess_channel.restype, L.orc_chess_channel.argtypes = C.c_int32, [C.c_int32, C.c_uint16]
        L.orc_chess_action.restype, L.orc_chess_action.argtypes = C.c_uint16, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.orc_chess_export.argtypes = [vp, C.POINTER(ChessState), C.POINTER(C.c_uint64)]
        L.orc_chess_det_hash.restype, L.orc_chess_det_hash.argtypes = C.c_uint64, [vp]
        u32p = C.POINTER(C.c_uint32)
        L.orc_chess_forest_new.restype, L.orc_chess_forest_new.argtypes = vp, [C.c_uint32, C.c_float]
        L.orc_chess_forest_free.argtypes = [vp]
        L.orc_chess_forest_reset.argtypes = [vp, C.c_uint32, vp]
        L.orc_chess_forest_search.argtypes = [vp, C.c_uint32, C.c_int32, CHESS_EVAL_FN, vp]
        L.orc_chess_forest_root_children.restype, L.orc_chess_forest_root_children.argtypes = C.c_int32, [vp, C.c_uint32, u16p, u32p, u32p]
        L.orc_chess_forest_arena_len.restype, L.orc_chess_forest_arena_len.argtypes = C.c_uint32, [vp, C.c_uint32]
        L.orc_chess_forest_node.restype = C.c_int32
        L.orc_chess_forest_node.argtypes = [vp, C.c_uint32, C.c_uint32, u32p, f32p, f32p, u32p, u32p, u16p]
        L.orc_chess_forest_use_subtree.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.orc_chess_forest_state.restype, L.orc_chess_forest_state.argtypes = vp, [vp, C.c_uint32, C.c_uint32]
        L.orc_chess_forest_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
        _bound = True
    return L


class Game:
    """`State` of chess.rs: position + transposition_table + fifty-move counter."""

    def __init__(self, fen: str | None = None, _h=None):
        self._h = _h if _h is not None else _lib().orc_chess_new(fen.encode() if fen else None)
        if not self._h:
            raise ValueError("bad FEN")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:           # module globals are gone at interpreter shutdown
            try:
                _lib().orc_chess_free(self._h)
            except Exception:
                pass
            self._h = None

    def clone(self):
        return Game(_h=_lib().orc_chess_clone(self._h))

    def legal_moves(self):
        buf = (C.c_uint16 * MAX_MOVES)()
        n = _lib().orc_chess_legal_moves(self._h, buf)
        return list(buf[:n])

    def make_move(self, m: int) -> int:
        return _lib().orc_chess_make_move(self._h, m)

    def status(self) -> int:
        return _lib().orc_chess_status(self._h)

    def repetitions(self) -> int:
        return _lib().orc_chess_repetitions(self._h)

    def value(self) -> float:
        return _lib().orc_chess_value(self._h)

    def side(self) -> int:
        return _lib().orc_chess_side(self._h)

    def perft(self, depth: int) -> int:
        return int(_lib().orc_chess_perft(self._h, depth))

    def encode(self) -> np.ndarray:
        out = np.zeros((PLANES, 8, 8), np.float32)
        _lib().orc_chess_encode(self._h, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def export(self):
        """-> (ChessState, history[MAX_HISTORY] u64) in the C ABI's form."""
        s = ChessState()
        hist = np.zeros(MAX_HISTORY, np.uint64)
        _lib().orc_chess_export(self._h, C.byref(s), hist.ctypes.data_as(C.POINTER(C.c_uint64)))
        return s, hist


class Forest:
    """`Mcts::search` over chess trees (src/mcts.rs restated in oracle/chess_oracle.cc)."""

    def __init__(self, n: int, c: float = 2.0):
        self._h = _lib().orc_chess_forest_new(n, c)
        self.n = n

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            try:
                _lib().orc_chess_forest_free(self._h)
            except Exception:
                pass
            self._h = None

    def reset(self, slot: int, game: "Game"):
        _lib().orc_chess_forest_reset(self._h, slot, game._h)

    def search(self, num_searches: int, evaluator: int, net_fn=None):
        """evaluator: 1 DetEval, 2 uniform, 0 network via net_fn(encodings[n,19,8,8]) -> (softmax probs[n,4672], values[n])."""
        if net_fn is None:
            cb = CHESS_EVAL_FN()
        else:
            def _cb(user, enc, n, probs, values):
                e = np.ctypeslib.as_array(enc, shape=(n, PLANES, 8, 8))
                p, v = net_fn(e)
                np.ctypeslib.as_array(probs, shape=(n, POLICY_SIZE))[:] = p
                np.ctypeslib.as_array(values, shape=(n,))[:] = v
            cb = CHESS_EVAL_FN(_cb)
        _lib().orc_chess_forest_search(self._h, num_searches, evaluator, cb, None)

    def root_children(self, slot: int):
        mv, cnt, ids = (C.c_uint16 * MAX_MOVES)(), (C.c_uint32 * MAX_MOVES)(), (C.c_uint32 * MAX_MOVES)()
        n = _lib().orc_chess_forest_root_children(self._h, slot, mv, cnt, ids)
        return list(mv[:n]), list(cnt[:n]), list(ids[:n])

    def arena_len(self, slot: int) -> int:
        return _lib().orc_chess_forest_arena_len(self._h, slot)

    def node(self, slot: int, node_id: int) -> dict:
        n, fc, nc = C.c_uint32(), C.c_uint32(), C.c_uint32()
        w, p = C.c_float(), C.c_float()
        mv = C.c_uint16()
        rc = _lib().orc_chess_forest_node(self._h, slot, node_id, C.byref(n), C.byref(w), C.byref(p), C.byref(fc), C.byref(nc), C.byref(mv))
        assert rc == 0
        return dict(visit_count=n.value, value_sum=w.value, prior=p.value, first_child=fc.value, n_children=nc.value, move=mv.value)

    def use_subtree(self, slot: int, node_id: int):
        _lib().orc_chess_forest_use_subtree(self._h, slot, node_id)

    def state(self, slot: int, node_id: int) -> "Game":
        return Game(_h=_lib().orc_chess_forest_state(self._h, slot, node_id))

    def counters(self) -> dict:
        v = (C.c_uint64 * 5)()
        _lib().orc_chess_forest_counters(self._h, v)
        return dict(zip(("simulations", "evaluations", "terminal_leaves", "path_length_sum", "children_created"), (int(x) for x in v)))


def det_hash(game: "Game") -> int:
    return int(_lib().orc_chess_det_hash(game._h))


def channel(player: int, move: int) -> int:
    return _lib().orc_chess_channel(player, move)


def action(player: int, ch: int, row: int, col: int) -> int:
    return _lib().orc_chess_action(player, ch, row, col)


def move_str(m: int) -> str:
    f, t, p = m & 63, (m >> 6) & 63, (m >> 12) & 7
    return "%s%d%s%d%s" % ("abcdefgh"[f % 8], f // 8 + 1, "abcdefgh"[t % 8], t // 8 + 1, ["", "n", "b", "r", "q"][p])
