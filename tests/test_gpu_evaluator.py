"""GPU parity: the fused bf16 evaluator vs the torch fp32 restatement of the reference net.
Tolerance (north_star): 1e-2 relative on policy logits and value."""
import numpy as np
import pytest

import selfplay_b200 as S
from oracle import torch_net
from helpers import random_states, synthetic_roots
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-2


def _check(game, flags, n, seed, blob_fn):
    net = torch_net.make_net(game, seed=seed)
    states = random_states(game, n, seed=seed + 1, include_terminal=False)
    enc = np.stack([O.encode(game, s) for s in states])
    probs_ref, v_ref, logit_ref = torch_net.forward_probs(net, enc)
    with S.Engine(game=game, num_games=4, evaluator=S.EVAL_NET, flags=flags) as e:
        e.load_weights(blob_fn(net))
        pol, v, lg = e.predict(states, want_logits=True)
    scale = float(np.abs(logit_ref).max())
    assert np.allclose(lg, logit_ref, rtol=RTOL, atol=RTOL * scale), float(np.abs(lg - logit_ref).max())
    assert np.allclose(v, v_ref, rtol=RTOL, atol=RTOL), float(np.abs(v - v_ref).max())
    want_pol = np.stack([O.mask_invalid_actions(game, s, p) for s, p in zip(states, probs_ref)])
    assert np.allclose(pol, want_pol, rtol=5 * RTOL, atol=1e-3)
    assert np.allclose(pol.sum(1), 1.0, atol=1e-5)
    for s, p in zip(states, pol):
        legal = set(O.valid_actions(game, s))
        assert all((p[a] > 0) == (a in legal) or p[a] == 0 for a in range(len(p)))
        assert all(p[a] == 0 for a in range(len(p)) if a not in legal)
    return lg, v


@pytest.mark.parametrize("game", [S.GAME_C4, S.GAME_TTT], ids=["c4", "ttt"])
def test_simt_evaluator_matches_torch(game):
    _check(game, S.FLAG_EVAL_SIMT, 96, 0, torch_net.to_safetensors_tch)


@pytest.mark.parametrize("game", [S.GAME_C4, S.GAME_TTT], ids=["c4", "ttt"])
@pytest.mark.parametrize("n", [1, 9, 10, 300, 2000])
def test_tcgen05_evaluator_matches_torch(game, n):
    _check(game, 0, n, 2, torch_net.to_safetensors_tch)


@pytest.mark.parametrize("game", [S.GAME_C4, S.GAME_TTT], ids=["c4", "ttt"])
@pytest.mark.parametrize("n", [1, 19, 2000])
def test_first_tcgen05_kernel_matches_torch(game, n):
    """SPB_FLAG_EVAL_V1: one MMA group per tap (N = 64), kept as a cross-check of the default kx-pair kernel."""
    _check(game, S.FLAG_EVAL_V1, n, 2, torch_net.to_safetensors_tch)


@pytest.mark.parametrize("game", [S.GAME_C4, S.GAME_TTT], ids=["c4", "ttt"])
@pytest.mark.parametrize("n", [1, 2, 19, 37, 2000])
def test_cta_pair_kernel_is_bit_identical_to_the_default(game, n):
    """SPB_FLAG_EVAL_PAIR2 (tcgen05 cta_group::2: two CTAs share every B operand) performs the same arithmetic in the same
    order as the default kernel: identical logits, values and policies for every batch shape (odd counts leave the peer
    CTA of a pair with fewer or no boards)."""
    net = torch_net.make_net(game, seed=5)
    states = random_states(game, n, seed=11, include_terminal=False)
    outs = []
    for flags in (0, S.FLAG_EVAL_PAIR2):
        with S.Engine(game=game, num_games=4, evaluator=S.EVAL_NET, flags=flags) as e:
            e.load_weights(torch_net.to_safetensors_tch(net))
            outs.append(e.predict(states, want_logits=True))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def test_tcgen05_and_simt_agree_closely():
    lg0, v0 = _check(S.GAME_C4, 0, 257, 4, torch_net.to_safetensors_explicit)
    lg1, v1 = _check(S.GAME_C4, S.FLAG_EVAL_SIMT, 257, 4, torch_net.to_safetensors_explicit)
    scale = float(np.abs(lg1).max())
    assert np.abs(lg0 - lg1).max() <= 2e-3 * scale and np.abs(v0 - v1).max() <= 2e-3


@pytest.mark.parametrize("flags", [0, S.FLAG_NO_GRAPH], ids=["graph", "nograph"])
def test_network_search_invariants(flags):
    G, sims = 64, 120
    net = torch_net.make_net(S.GAME_C4, seed=0)
    roots = synthetic_roots(S.GAME_C4, G, start=50)
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags) as e:
        e.load_weights(torch_net.to_safetensors_tch(net))
        e.reset_games(roots)
        e.search(sims)
        a, c, i, n = e.root_children_all()
        for slot in range(G):
            assert int(c[slot].sum()) == sims - 1 and e.node_stats(slot, 0)["visit_count"] == sims
            assert n[slot] == len(O.valid_actions(O.GAME_C4, roots[slot]))
        ctr = e.counters()
        assert ctr["simulations"] == G * sims and ctr["evaluations"] + ctr["terminal_leaves"] == G * sims
        # priors of the root children = masked softmax of the evaluator for the root position
        pol, _ = e.predict(roots[:8])
        for slot in range(8):
            fc = e.node_stats(slot, 0)["first_child"]
            for j in range(int(n[slot])):
                assert e.node_stats(slot, fc + j)["prior"] == pol[slot, a[slot, j]]


def test_network_search_close_to_oracle_with_torch_evaluator():
    """Same search, oracle + torch fp32 evaluator vs GPU + bf16 evaluator: root visit distributions agree
    closely (not bit-exact: the evaluators differ by ~1e-3)."""
    G, sims = 16, 200
    net = torch_net.make_net(S.GAME_C4, seed=0)
    roots = synthetic_roots(S.GAME_C4, G, start=50)
    cb = O.make_eval_callback(O.GAME_C4, lambda enc: torch_net.forward_probs(net, enc)[:2])
    f = O.Forest(O.GAME_C4, G)
    f.reset(roots)
    f.search(sims, O.EVAL_NET, cb)
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET) as e:
        e.load_weights(torch_net.to_safetensors_tch(net))
        e.reset_games(roots)
        e.search(sims)
        tv = 0.0
        for slot in range(G):
            tv += np.abs(e.root_policy(slot) - f.root_policy(slot)).sum() / 2
    assert tv / G < 0.05, tv / G
