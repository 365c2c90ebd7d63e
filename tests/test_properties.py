"""Property tests (hypothesis) of the search invariants, on the oracle (SURVEY.md §4 test pyramid)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from helpers import synthetic_root
from oracle import pyoracle as O


@settings(max_examples=25, deadline=None)
@given(game=st.sampled_from([O.GAME_TTT, O.GAME_C4]), g=st.integers(0, 5000), sims=st.integers(1, 120), carry=st.booleans(),
       ev=st.sampled_from([O.EVAL_DET, O.EVAL_UNIFORM]))
def test_tree_invariants(game, g, sims, carry, ev):
    root = synthetic_root(game, g, max_ply=21 if game == O.GAME_C4 else 5)
    f = O.Forest(game, 1)
    f.reset([root])
    f.search(sims, ev)
    total = sims
    if carry:
        acts, counts, ids = f.root_children(0)
        if ids:
            best = max(range(len(ids)), key=lambda j: (counts[j], j))
            kept = counts[best]
            f.use_subtree(0, ids[best])
            assert f.node_stats(0, 0)["visit_count"] == kept           # statistics survive the re-root (mcts.rs:161-192)
            f.search(sims, ev)
            total = kept + sims
    n = f.arena_len(0)
    root_st = f.node_stats(0, 0)
    assert root_st["visit_count"] == total
    seen_children = 0
    for i in range(n):
        s = f.node_stats(0, i)
        if s["n_children"]:
            fc = s["first_child"]
            assert fc > i and fc + s["n_children"] <= n                # children contiguous, after the parent (BFS / append order)
            kids = [f.node_stats(0, fc + j) for j in range(s["n_children"])]
            child_sum = sum(k["visit_count"] for k in kids)
            # every visit of an expanded node either expanded it (once) or went through a child
            assert s["visit_count"] == child_sum + 1 or (i == 0 and carry and s["visit_count"] >= child_sum)
            assert abs(sum(k["prior"] for k in kids) - 1.0) < 1e-5      # masked + renormalised priors
            state = f.get_state(0, i)
            assert [O.next_state(game, state, a).key() for a in O.valid_actions(game, state)] == \
                   [f.get_state(0, fc + j).key() for j in range(s["n_children"])]      # children in get_valid_actions order
            seen_children += s["n_children"]
        assert abs(s["value_sum"]) <= s["visit_count"] + 1e-4
    assert seen_children == n - 1
    pol = f.root_policy(0)
    if root_st["n_children"] and sum(f.root_children(0)[1]) > 0:
        assert abs(float(pol.sum()) - 1.0) < 1e-5 and np.all(pol >= 0)
    elif root_st["n_children"]:
        assert np.all(np.isnan(pol))       # the reference's normalize() divides by a zero sum here too (mcts.rs:328)
