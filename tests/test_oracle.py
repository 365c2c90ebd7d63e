"""CPU tests: the oracle against the survey's independent KATs, a second Python restatement,
and invariants.  No GPU, no /root/reference."""
import json
import os
import random

import numpy as np
import pytest

import pyref
from oracle import pyoracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "survey_kats.json")))


def _search(game, S, ev):
    f = O.Forest(game, 1)
    f.search(S, ev)
    return f


def test_det_hash_check_value():
    assert O.det_hash(O.GAME_C4, O.State()) == int(KAT["det_hash_empty_c4"], 16)


@pytest.mark.parametrize("S", sorted(KAT["ttt_uniform"], key=int))
def test_ttt_uniform_kats(S):
    f = _search(O.GAME_TTT, int(S), O.EVAL_UNIFORM)
    assert f.root_children(0)[1] == KAT["ttt_uniform"][S]
    if S == "600":
        assert f.arena_len(0) == KAT["ttt_uniform_arena_600"]


@pytest.mark.parametrize("S", sorted(KAT["c4_uniform"], key=int))
def test_c4_uniform_kats(S):
    f = _search(O.GAME_C4, int(S), O.EVAL_UNIFORM)
    assert f.root_children(0)[1] == KAT["c4_uniform"][S]
    if S in KAT["c4_uniform_arena"]:
        assert f.arena_len(0) == KAT["c4_uniform_arena"][S]
    if S == "800":
        c = f.counters()
        assert c["path_length_sum"] == KAT["c4_uniform_800_path_sum"]   # mean path 3.60
        assert c["terminal_leaves"] == 0


@pytest.mark.parametrize("S", sorted(KAT["c4_det_root_w"], key=int))
def test_c4_det_kats(S):
    f = _search(O.GAME_C4, int(S), O.EVAL_DET)
    assert f.node_stats(0, 0)["value_sum"] == KAT["c4_det_root_w"][S]
    if S in KAT["c4_det"]:
        assert f.root_children(0)[1] == KAT["c4_det"][S]
    if S in KAT["c4_det_arena"]:
        assert f.arena_len(0) == KAT["c4_det_arena"][S]
    if S == "800":
        c = f.counters()
        assert c["path_length_sum"] == KAT["c4_det_800_path_sum"]       # mean path 4.175
        assert c["terminal_leaves"] == KAT["c4_det_800_terminal_leaves"]


def test_c4_greedy_games():
    g = O.greedy_game(O.GAME_C4, 100, O.EVAL_UNIFORM)
    assert g["actions"] == KAT["c4_uniform_greedy_100_actions"] and g["final_status"] == O.TIED
    g = O.greedy_game(O.GAME_C4, 800, O.EVAL_DET)
    k = KAT["c4_det_greedy_800"]
    assert g["actions"] == k["actions"] and g["final_status"] == k["final_status"]
    assert g["last_counts"] == k["last_counts"] and g["arena_sizes"] == k["arena_sizes"]
    assert sum(g["last_counts"]) == 1132
    g = O.greedy_game(O.GAME_C4, 200, O.EVAL_DET)
    assert g["actions"] == KAT["c4_det_greedy_200"]["actions"]
    f = _search(O.GAME_C4, 200, O.EVAL_DET)
    assert f.root_children(0)[1] == KAT["c4_det_greedy_200"]["first_counts"]


def test_c4_antidiagonal_not_a_win():
    """connect_four.rs:163-176 only checks the (row+i, col+i) diagonal."""
    s = O.State()
    for a in KAT["c4_antidiagonal_moves"]:
        s = O.next_state(O.GAME_C4, s, a)
    x = s.stones[0]
    for (r, c) in [(0, 3), (1, 2), (2, 1), (3, 0)]:
        assert x >> (c * 7 + r) & 1
    assert s.status == O.ONGOING


def test_c4_main_diagonal_row_col_wins():
    def play(moves):
        s = O.State()
        for a in moves:
            s = O.next_state(O.GAME_C4, s, a)
        return s
    assert play([0, 0, 1, 1, 2, 2, 3]).status == O.WON                   # row
    assert play([0, 1, 0, 1, 0, 1, 0]).status == O.WON                   # column
    assert play([0, 1, 1, 2, 2, 3, 2, 3, 3, 6, 3]).status == O.WON       # main diagonal (0,0)..(3,3)
    s = play([0, 0, 1, 1, 2, 2, 3])
    assert O.next_state(O.GAME_C4, s, 4) is None                          # game has already ended
    assert O.valid_actions(O.GAME_C4, s) == []
    s = play([0] * 6)
    assert O.next_state(O.GAME_C4, s, 0) is None                          # column already filled
    assert O.valid_actions(O.GAME_C4, s) == [1, 2, 3, 4, 5, 6]


def _py_state(game):
    return pyref.C4() if game == O.GAME_C4 else pyref.TTT()


def _same(ps, os_):
    x, o = ps.stones()
    return (x, o, ps.player, ps.n, ps.status) == os_.key()


@pytest.mark.parametrize("game", [O.GAME_TTT, O.GAME_C4])
def test_game_rules_vs_python_restatement(game):
    rng = random.Random(1234 + game)
    A = O.NUM_ACTIONS[game]
    for _ in range(300):
        ps, os_ = _py_state(game), O.State()
        while True:
            assert _same(ps, os_)
            assert ps.valid_actions() == O.valid_actions(game, os_)
            assert np.array_equal(ps.encoding(), O.encode(game, os_))
            a = rng.randrange(A)                                         # includes illegal moves
            pn, on = ps.next_state(a), O.next_state(game, os_, a)
            assert (pn is None) == (on is None)
            if ps.status != 0:
                break
            if pn is not None:
                ps, os_ = pn, on


def test_ttt_exhaustive_state_space():
    """All 5,478 reachable tic-tac-toe positions: both restatements agree on each."""
    seen = {}
    stack = [(pyref.TTT(), O.State())]
    while stack:
        ps, os_ = stack.pop()
        k = os_.key()
        if k in seen:
            continue
        seen[k] = True
        assert _same(ps, os_)
        assert ps.valid_actions() == O.valid_actions(O.GAME_TTT, os_)
        for a in range(9):
            pn, on = ps.next_state(a), O.next_state(O.GAME_TTT, os_, a)
            assert (pn is None) == (on is None)
            if pn is not None:
                stack.append((pn, on))
    boards = {(k[0], k[1]) for k in seen}
    assert len(boards) == 5478


@pytest.mark.parametrize("game,ev,S", [(O.GAME_C4, O.EVAL_DET, 60), (O.GAME_C4, O.EVAL_UNIFORM, 40),
                                       (O.GAME_TTT, O.EVAL_DET, 80), (O.GAME_TTT, O.EVAL_UNIFORM, 30)])
def test_search_vs_python_restatement(game, ev, S):
    """Visit counts, value sums and arena layout agree node for node, incl. after use_subtree."""
    pe = pyref.det_eval if ev == O.EVAL_DET else pyref.uniform_eval
    pt = pyref.Tree(_py_state(game))
    f = O.Forest(game, 1)
    for move in range(3):
        pyref.search([pt], S, pe)
        f.search(S, ev)
        assert f.arena_len(0) == len(pt.arena)
        for i, node in enumerate(pt.arena):
            st = f.node_stats(0, i)
            assert st["visit_count"] == node["N"] and st["value_sum"] == float(node["W"])
            assert st["n_children"] == len(node["children"])
            if node["children"]:
                assert st["first_child"] == node["children"][0]
                assert node["children"] == list(range(node["children"][0], node["children"][0] + len(node["children"])))
            if node["prior"] is not None:
                assert st["prior"] == float(node["prior"])
        acts, counts, ids = f.root_children(0)
        if not ids:
            break
        best = max(range(len(ids)), key=lambda i: (counts[i], i))       # last max
        pt.use_subtree(ids[best])
        f.use_subtree(0, ids[best])
        assert _same(pt.arena[0]["state"], f.get_state(0, 0))
        if pt.arena[0]["state"].status != 0:
            break


def test_mask_invalid_actions_sum_order():
    rng = np.random.default_rng(0)
    for game in (O.GAME_TTT, O.GAME_C4):
        A = O.NUM_ACTIONS[game]
        for _ in range(200):
            ps, os_ = _py_state(game), O.State()
            for _ in range(rng.integers(0, 5)):
                va = ps.valid_actions()
                if not va:
                    break
                a = int(rng.choice(va))
                ps, os_ = ps.next_state(a), O.next_state(game, os_, a)
            if ps.status != 0:
                continue
            logits = rng.normal(size=A).astype(np.float32)
            p = np.exp(logits - logits.max()).astype(np.float32)
            p = (p / p.sum(dtype=np.float32)).astype(np.float32)
            want = np.array(pyref.mask(ps, list(p)), dtype=np.float32)
            got = O.mask_invalid_actions(game, os_, p)
            assert np.array_equal(want, got)


def test_invariants_and_batch_independence():
    """N_root = sum(child N) + 1 on a fresh tree; trees in one forest do not interact."""
    roots = []
    s = O.State()
    for a in [3, 3, 4]:
        roots.append(s.copy())
        s = O.next_state(O.GAME_C4, s, a)
    roots.append(s.copy())
    f = O.Forest(O.GAME_C4, len(roots))
    f.reset(roots)
    f.search(150, O.EVAL_DET)
    for i, r in enumerate(roots):
        g = O.Forest(O.GAME_C4, 1)
        g.reset([r])
        g.search(150, O.EVAL_DET)
        assert g.root_children(0) == f.root_children(i)
        assert sum(f.root_children(i)[1]) == 149 and f.node_stats(i, 0)["visit_count"] == 150
        pol = f.root_policy(i)
        assert abs(float(pol.sum()) - 1.0) < 1e-6
    c = f.counters()
    assert c["simulations"] == 150 * len(roots)
    assert c["evaluations"] + c["terminal_leaves"] == c["simulations"]


def test_callback_evaluator_matches_builtin():
    """SPB_EVAL_NET through the callback (encodings -> probs) reproduces DetEval when the callback
    recomputes the hash from the encoding planes."""
    def fn(enc):
        n = enc.shape[0]
        probs = np.zeros((n, 7), np.float32)
        vals = np.zeros(n, np.float32)
        for i in range(n):
            mine = opp = 0
            for r in range(6):
                for c in range(7):
                    if enc[i, 0, r, c]:
                        mine |= 1 << (c * 7 + r)
                    if enc[i, 1, r, c]:
                        opp |= 1 << (c * 7 + r)
            h = pyref.splitmix64(mine ^ pyref.splitmix64(opp))
            probs[i] = [(1 + ((h >> (4 * a)) & 7)) / 64 for a in range(7)]
            vals[i] = (((h >> 40) & 0xFF) - 128) / 128
        return probs, vals
    cb = O.make_eval_callback(O.GAME_C4, fn)
    f = O.Forest(O.GAME_C4, 2)
    f.search(100, O.EVAL_NET, cb)
    assert f.root_children(0)[1] == KAT["c4_det"]["100"] == f.root_children(1)[1]


def test_committed_golden_vectors_come_from_the_generator():
    """tests/golden/pyref_kats.json is the output of tools/gen_golden.py (the pure-Python restatement); every vector that
    SURVEY.md §8(c) lists (survey_kats.json) equals it, and a re-run of one generator case reproduces the committed value."""
    gen = json.load(open(os.path.join(HERE, "golden", "pyref_kats.json")))
    for k, v in KAT.items():
        if k != "_source":
            assert gen[k] == v, k
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(os.path.dirname(HERE), "tools", "gen_golden.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    assert g.counts(g.R.C4, 100, g.R.det_eval).root_counts() == gen["c4_det"]["100"]
    assert g.greedy(g.R.C4, 200, g.R.det_eval)["actions"] == gen["c4_det_greedy_200"]["actions"]
