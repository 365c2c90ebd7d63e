"""EXTENSION (BASELINE.json config 4): K in-flight leaves per tree with virtual loss.  The reference has no such
search (one leaf per tree per step, mcts.rs:236-252), so parity is pinned only by our own oracle of the same
definition (oracle/oracle.cc search_vl); K = 1 must stay on the reference algorithm bit for bit."""
import json
import os

import numpy as np
import pytest

import selfplay_b200 as S
from helpers import synthetic_roots
from oracle import pyoracle as O

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "survey_kats.json")))


@pytest.mark.parametrize("K", [2, 5, 16])
def test_oracle_virtual_loss_invariants(K):
    S_ = 203                                                     # not a multiple of K
    f = O.Forest(O.GAME_C4, 3, leaves_per_tree=K)
    f.search(S_, O.EVAL_DET)
    for slot in range(3):
        root = f.node_stats(slot, 0)
        assert root["visit_count"] == S_                         # every simulation visits the root exactly once
        n = f.arena_len(slot)
        # no virtual loss left behind: for an expanded node N = own evaluations/terminal visits + sum of children N
        for i in range(n):
            st = f.node_stats(slot, i)
            if st["n_children"]:
                kids = sum(f.node_stats(slot, st["first_child"] + j)["visit_count"] for j in range(st["n_children"]))
                assert st["visit_count"] >= kids + 1
            assert abs(st["value_sum"]) <= st["visit_count"] + 1e-3
    c = f.counters()
    assert c["simulations"] == 3 * S_
    assert c["evaluations"] + c["terminal_leaves"] <= c["simulations"]      # duplicates share an evaluation


def test_oracle_k1_is_the_reference_algorithm():
    f = O.Forest(O.GAME_C4, 1, leaves_per_tree=1)
    f.search(800, O.EVAL_DET)
    assert f.root_children(0)[1] == KAT["c4_det"]["800"]


@pytest.mark.gpu
@pytest.mark.parametrize("game,K,ev", [(S.GAME_C4, 16, S.EVAL_DET), (S.GAME_C4, 4, S.EVAL_UNIFORM), (S.GAME_C4, 3, S.EVAL_DET),
                                       (S.GAME_TTT, 8, S.EVAL_DET)])
def test_multi_leaf_search_matches_oracle_bit_for_bit(game, K, ev):
    G = 24
    roots = synthetic_roots(game, G, start=300, max_ply=21 if game == S.GAME_C4 else 4)
    f = O.Forest(game, G, leaves_per_tree=K)
    f.reset(roots)
    with S.Engine(game=game, num_games=G, evaluator=ev, leaves_per_tree=K) as e:
        e.reset_games(roots)
        for sims in (1, 50, 203):
            e.search(sims)
            f.search(sims, ev)
            for slot in range(G):
                n = f.arena_len(slot)
                assert e.arena_len(slot) == n
                assert e.root_children(slot) == f.root_children(slot)
                ids = range(n) if slot < 3 else list(range(min(n, 30))) + list(range(max(0, n - 30), n))
                for i in ids:
                    assert e.node_stats(slot, i) == f.node_stats(slot, i), (slot, i)
        ce, cf = e.counters(), f.counters()
        for k in ("simulations", "evaluations", "terminal_leaves", "path_length_sum", "children_created", "nodes_live"):
            assert ce[k] == cf[k], k
        # subtree reuse works on top of it
        acts, counts, ids = f.root_children(0)
        best = max(range(len(ids)), key=lambda j: (counts[j], j))
        e.advance([ids[best]], slots=[0])
        f.use_subtree(0, ids[best])
        e.search(40)
        f.search(40, ev)
        assert e.root_children(0) == f.root_children(0) and e.arena_len(0) == f.arena_len(0)


@pytest.mark.gpu
def test_multi_leaf_search_with_network():
    from oracle import torch_net
    G, K, sims = 32, 16, 160
    net = torch_net.make_net(S.GAME_C4, seed=0)
    roots = synthetic_roots(S.GAME_C4, G, start=50)
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, leaves_per_tree=K) as e:
        e.load_weights(torch_net.to_safetensors_tch(net))
        e.reset_games(roots)
        e.search(sims)
        a, c, i, n = e.root_children_all()
        for slot in range(G):
            assert e.node_stats(slot, 0)["visit_count"] == sims
            assert 0 < int(c[slot].sum()) <= sims - 1
        ctr = e.counters()
        assert ctr["simulations"] == G * sims and ctr["evaluations"] <= G * sims
        first = c.copy()
        e.reset_games(roots)
        e.search(sims)
        assert np.array_equal(e.root_children_all()[1], first)           # deterministic run to run
