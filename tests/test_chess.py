"""Chess rules (BASELINE config 5): the oracle against the published perft counts and hand-made known answers (CPU), and
the device rules (csrc/chess.cuh, bitboards) against the oracle (array board) through the C ABI (GPU).

PARITY UNPINNED BY THE REFERENCE where the un-vendored crate `chess 3.2.0` decides (the order of the legal moves); what
src/game/chess.rs computes itself — repetition on legal-move lists (:51-62), the reversible-move counter (:124-143), get_status
(:154-166), +1.0 for Won (:168-174), the 19x8x8 encoding (:176-249), the 73 move planes (:311-493) — is checked line by line.
"""
import os

import numpy as np
import pytest

import selfplay_b200 as S
from oracle import pychess as P

CH = S.chess


def mv(s):
    """'e2e4' / 'a7a8q' -> move code."""
    f = (ord(s[0]) - 97) + 8 * (int(s[1]) - 1)
    t = (ord(s[2]) - 97) + 8 * (int(s[3]) - 1)
    p = {"": 0, "n": 1, "b": 2, "r": 3, "q": 4}[s[4:]]
    return f | (t << 6) | (p << 12)


def play(moves, fen=None):
    g = P.Game(fen)
    for m in moves.split():
        assert g.make_move(mv(m)) == 0, m
    return g


def random_games(n_games, seed, max_plies=120):
    """Random playouts -> list of (Game after every ply) including finished positions."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_games):
        g = P.Game()
        out.append(g.clone())
        for _ in range(int(rng.integers(1, max_plies))):
            if g.status() != S.ONGOING:
                break
            lm = g.legal_moves()
            assert g.make_move(int(lm[rng.integers(len(lm))])) == 0
            out.append(g.clone())
    return out


# ---- CPU: the oracle ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("fen", list(P.PERFT))
def test_oracle_perft_matches_published_counts(fen):
    g = P.Game(fen)
    for depth, want in enumerate(P.PERFT[fen][:3], 1):
        assert g.perft(depth) == want, (fen, depth)


def test_oracle_perft_depth4_start_and_kiwipete():
    assert P.Game().perft(4) == 197281
    assert P.Game(P.KIWIPETE).perft(3) == 97862


def test_legal_moves_are_sorted_by_from_to_promotion():
    for g in random_games(5, seed=1):
        lm = g.legal_moves()
        keys = [((m & 63), (m >> 6) & 63, m >> 12) for m in lm]
        assert keys == sorted(keys) and len(set(lm)) == len(lm)


def test_fools_mate_is_won_with_value_plus_one():
    g = play("f2f3 e7e5 g2g4 d8h4")
    assert g.status() == S.WON and g.legal_moves() == []
    assert g.value() == 1.0                                   # chess.rs:172: +1.0, unlike connect_four.rs:236 (-1.0)
    assert g.make_move(mv("a2a3")) == -4                      # chess.rs:113-115 "Game is already over"


def test_stalemate_is_tied():
    g = P.Game("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1")
    assert g.legal_moves() == [] and g.status() == S.TIED and g.value() == 0.0


def test_repetition_counts_legal_move_lists():
    # knights out and back: the start position's legal-move list occurs for the 3rd time after 8 plies (chess.rs:51-62)
    g = P.Game()
    seq = "g1f3 g8f6 f3g1 f6g8".split()
    reps = [g.repetitions()]
    for m in seq + seq:
        assert g.make_move(mv(m)) == 0
        reps.append(g.repetitions())
    assert reps == [1, 1, 1, 1, 2, 2, 2, 2, 3]
    assert g.status() == S.TIED
    assert len(g.legal_moves()) == 20                         # get_valid_actions does not look at the status (chess.rs:150-152)
    assert g.make_move(mv("e2e4")) == -4


def test_fifty_move_counter_follows_chess_rs():
    g = play("g1f3 g8f6")
    s, _ = g.export()
    assert s.fifty == 2
    g.make_move(mv("e2e4"))                                   # pawn move resets (chess.rs:131)
    assert g.export()[0].fifty == 0
    g = play("e2e4 d7d5 e4d5")                                # capture resets (:132)
    assert g.export()[0].fifty == 0
    g = play("e2e4 e7e5 e1e2")                                # king move loses castle rights: not reversible (:133-134)
    assert g.export()[0].fifty == 0
    g = play("e2e4 e7e5 e1e2 e8e7 e2e1 e7e8")                 # both sides' rights already gone: reversible again
    assert g.export()[0].fifty == 2
    # counter >= 100 -> Tied (chess.rs:160)
    g = P.Game("4k3/8/8/8/8/8/8/R3K3 w - - 98 60")
    assert g.export()[0].fifty == 98 and g.export()[0].plies == 118
    assert g.make_move(mv("a1b1")) == 0 and g.status() == S.ONGOING
    assert g.make_move(mv("e8f8")) == 0
    assert g.export()[0].fifty == 100 and g.status() == S.TIED and g.repetitions() == 1 and g.value() == 0.0


def test_repetition_ignores_the_opponents_pieces():
    # The reference compares the legal-move list of the SIDE TO MOVE only (chess.rs:51-62), so a position whose other
    # side's pieces stand elsewhere still counts: the black king walks an 8-ply cycle while the white rook never returns
    # to a square, and the game is Tied when black's list at e8 occurs for the third time (plies 1, 17, 33).
    g = P.Game("4k3/8/8/8/8/8/8/R3K3 w - - 0 1")
    squares_w = ["a1", "b1", "c1", "d1", "d2", "c2", "b2", "a2", "a3", "b3", "c3", "d3", "d4", "c4", "b4", "a4", "a5", "b5", "c5", "d5"]
    path_w = [a + b for a, b in zip(squares_w, squares_w[1:])]
    path_b = ["e8f8", "f8g8", "g8h8", "h8h7", "h7g7", "g7f7", "f7e7", "e7e8"]
    n = 0
    while g.status() == S.ONGOING:
        m = path_w[n // 2] if n % 2 == 0 else path_b[(n // 2) % len(path_b)]
        assert g.make_move(mv(m)) == 0, (n, m)
        n += 1
    assert n == 33 and g.side() == 1 and g.repetitions() == 3 and g.status() == S.TIED and g.export()[0].fifty == 33


def test_encoding_planes_follow_chess_rs():
    g = play("e2e4")                                          # black to move: rows flipped, files not (chess.rs:184-187)
    enc = g.encode()
    assert enc.shape == (19, 8, 8)
    # own (black) pawns on rank 7 -> row 7-6 = 1; opponent's e-pawn on e4 (rank 3) -> row 4, col 4
    assert enc[0, 1].tolist() == [1.0] * 8 and enc[0].sum() == 8
    assert enc[6, 4, 4] == 1.0 and enc[6, 6].sum() == 7 and enc[6].sum() == 8
    assert enc[5, 0, 4] == 1.0 and enc[11, 7, 4] == 1.0        # own king e8 -> row 0; white king e1 -> row 7
    for p in (12, 13, 14, 15):
        assert (enc[p] == 1.0).all()
    assert (enc[16] == 1.0).all() and (enc[17] == 0.0).all() and (enc[18] == 0.0).all()
    g = play("g1f3 g8f6 f3g1 f6g8")
    enc = g.encode()
    assert (enc[16] == 2.0).all() and (enc[17] == np.float32(4) / np.float32(100)).all()
    assert (enc[18] == np.float32(2) / np.float32(50)).all()   # 4 plies -> 2 full moves (chess.rs:240-244)


def test_move_planes_round_trip_and_the_knight_promotion_slip():
    seen = set()
    for g in random_games(30, seed=7) + [P.Game(f) for f in P.PERFT if f]:
        side = g.side()
        for m in g.legal_moves():
            ch = P.channel(side, m)
            assert 0 <= ch < 73
            seen.add(ch)
            f = m & 63
            row = 7 - (f >> 3) if side else (f >> 3)
            back = P.action(side, ch, row, f & 7)
            promo = m >> 12
            if promo == 4:
                assert back == (m & 0xFFF)                     # queen promotions travel on the ordinary planes (no piece in get_action)
            elif promo == 1:
                # chess.rs:442 compares with KNIGHT_MOVE_START_IDX: the knight-promotion planes decode to file differences +2..+4
                want_file = (f & 7) + (ch - 3 - 1)
                assert back == (P.MOVE_NONE if hasattr(P, "MOVE_NONE") else 0xFFFF) or ((back >> 6) & 7) == want_file
            else:
                assert back == m, (P.move_str(m), ch)
    assert len(seen) > 60


def test_host_move_plane_functions_equal_the_oracle():
    for side in (0, 1):
        for f in range(64):
            for t in range(64):
                if f == t:
                    continue
                dr, df = abs((t >> 3) - (f >> 3)), abs((t & 7) - (f & 7))
                if not (dr == 0 or df == 0 or dr == df or {dr, df} == {1, 2}):
                    continue
                for promo in (0, 1, 2, 3, 4):
                    if promo and not (df <= 1 and ((t >> 3) - (f >> 3)) * (1 if side == 0 else -1) == 1):
                        continue
                    m = f | (t << 6) | (promo << 12)
                    assert CH.move_channel(side, m) == P.channel(side, m)
                    row = 7 - (f >> 3) if side else (f >> 3)
                    assert CH.policy_index(side, m) == P.channel(side, m) * 64 + row * 8 + (f & 7)
        for ch in range(73):
            for row in range(8):
                for col in range(8):
                    assert CH.action(side, ch, row, col) == P.action(side, ch, row, col)


def test_start_position_equals_the_oracle_export():
    s, _ = P.Game().export()
    got = CH.start_position()[0]
    assert [int(x) for x in got["piece"]] == [int(x) for x in s.piece]
    assert [int(x) for x in got["color"]] == [int(x) for x in s.color]
    assert (got["side"], got["castle"], got["ep"], got["fifty"], got["plies"], got["hist_len"]) == (0, 15, 64, 0, 0, 0)


# ---- GPU: device rules vs the oracle through the C ABI ---------------------------------------------------------------

def export_all(games):
    st = np.zeros(len(games), S.CHESS_STATE_DTYPE)
    hist = np.zeros((len(games), CH.MAX_HISTORY), np.uint64)
    for i, g in enumerate(games):
        s, h = g.export()
        st[i] = np.frombuffer(bytes(s), S.CHESS_STATE_DTYPE)[0]
        hist[i] = h
    return st, hist


@pytest.fixture(scope="module")
def rules():
    with S.ChessEngine(num_games=1, evaluator=S.EVAL_DET) as e:
        yield e


@pytest.mark.gpu
@pytest.mark.parametrize("fen", list(P.PERFT))
def test_device_perft(rules, fen):
    g = P.Game(fen)
    st, _ = export_all([g])
    for depth, want in enumerate(P.PERFT[fen], 1):
        assert rules.perft(st, depth) == want, (fen, depth)
    assert rules.perft(st, 0) == 1


@pytest.mark.gpu
def test_device_legal_moves_status_and_repetitions_match_oracle(rules):
    games = random_games(40, seed=11, max_plies=200)
    games += [play("f2f3 e7e5 g2g4 d8h4"), P.Game("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1"), play("g1f3 g8f6 f3g1 f6g8 g1f3 g8f6 f3g1 f6g8")]
    games += [P.Game(f) for f in P.PERFT if f]
    st, hist = export_all(games)
    moves, counts, pidx, status, reps = rules.legal_moves(st, hist)
    for i, g in enumerate(games):
        lm = g.legal_moves()
        assert counts[i] == len(lm) and moves[i, :len(lm)].tolist() == lm, i
        assert (moves[i, len(lm):] == 0xFFFF).all()
        assert status[i] == g.status() and reps[i] == g.repetitions(), i
        side = g.side()
        assert pidx[i, :len(lm)].tolist() == [P.channel(side, m) * 64 + ((7 - ((m & 63) >> 3)) if side else ((m & 63) >> 3)) * 8 + (m & 7)
                                              for m in lm]
        assert len(set(pidx[i, :len(lm)].tolist())) == len(lm)          # no two legal moves share a policy cell
    assert (status != S.ONGOING).sum() >= 3


@pytest.mark.gpu
def test_device_next_states_match_oracle_including_illegal_moves(rules):
    rng = np.random.default_rng(5)
    games = random_games(25, seed=13, max_plies=150)
    st, hist = export_all(games)
    chosen, want = [], []
    for g in games:
        lm = g.legal_moves()
        r = rng.random()
        if r < 0.15 or not lm:
            m = int(rng.integers(0, 1 << 15))                            # mostly illegal
        else:
            m = int(lm[rng.integers(len(lm))])
        h = g.clone()
        rc = h.make_move(m)
        chosen.append(m)
        want.append((rc, h))
    out, hist2, err = rules.next_states(st, hist, np.array(chosen, np.uint16))
    for i, (rc, h) in enumerate(want):
        assert err[i] == rc, (i, chosen[i])
        ws, wh = h.export()
        assert out[i].tobytes() == bytes(ws), i
        assert (hist2[i] == wh).all(), i
    assert (err != 0).sum() > 10 and (err == 0).sum() > 100


@pytest.mark.gpu
def test_device_encoding_matches_oracle(rules):
    games = random_games(20, seed=17, max_plies=160) + [play("g1f3 g8f6 f3g1 f6g8 g1f3")]
    st, hist = export_all(games)
    enc = rules.encode(st, hist)
    for i, g in enumerate(games):
        assert np.array_equal(enc[i], g.encode()), i


@pytest.mark.gpu
def test_device_handles_empty_and_single_batches(rules):
    st, hist = export_all([P.Game()])
    assert rules.legal_moves(st[:0], hist[:0])[1].shape == (0,)
    moves, counts, _, status, reps = rules.legal_moves(st, None)
    assert counts[0] == 20 and status[0] == S.ONGOING and reps[0] == 1


# ---- CPU: the product's shared rule code (csrc/chess.cuh is __host__ __device__) compiled for the host ------------------

@pytest.fixture(scope="module")
def host_rules(tmp_path_factory):
    import ctypes as C
    import shutil
    import subprocess
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path_factory.mktemp("chess_host") / "libchess_host.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--fmad=false", "-shared", "-Xcompiler", "-fPIC",
                           "-o", out, os.path.join(here, "chess_host_check.cu")])
    L = C.CDLL(out)
    L.host_perft.restype = C.c_uint64
    L.host_list_hash.restype = C.c_uint64
    L.host_det_hash.restype = C.c_uint64
    L.host_encode.restype = C.c_float
    L.host_det_raw_prob.restype, L.host_det_raw_prob.argtypes = C.c_float, [C.c_uint64, C.c_int]
    L.host_det_value.restype, L.host_det_value.argtypes = C.c_float, [C.c_uint64]
    return L


def test_shared_rule_code_matches_oracle_on_the_host(host_rules):
    import ctypes as C
    L = host_rules
    buf = (C.c_uint16 * 256)()
    for fen in P.PERFT:
        s, _ = P.Game(fen).export()
        assert L.host_perft(C.byref(s), 3) == P.PERFT[fen][2], fen
    for g in random_games(12, seed=3, max_plies=150) + [P.Game(f) for f in P.PERFT if f]:
        s, hist = g.export()
        lm = g.legal_moves()
        n = L.host_legal(C.byref(s), buf)
        assert list(buf[:n]) == lm
        assert L.host_det_hash(C.byref(s)) == P.det_hash(g)
        for m in lm[:6]:
            h = g.clone()
            assert h.make_move(m) == 0 or g.status() != S.ONGOING
            if g.status() == S.ONGOING:
                out = P.ChessState()
                L.host_apply(C.byref(s), m, C.byref(out))
                want, wh = h.export()
                out.hist_len = want.hist_len
                assert bytes(out) == bytes(want), P.move_str(m)
                arr = (C.c_uint16 * len(lm))(*lm)
                assert L.host_list_hash(arr, len(lm)) == int(wh[want.hist_len - 1])
        enc = g.encode()
        reps = g.repetitions()
        got = np.array([[[L.host_encode(C.byref(s), reps, p, r, c) for c in range(8)] for r in range(8)] for p in range(19)], np.float32)
        assert np.array_equal(got, enc)
        if not lm:
            assert bool(L.host_in_check(C.byref(s))) == (g.status() == S.WON)      # mate vs stalemate
