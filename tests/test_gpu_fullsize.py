"""GPU parity at BASELINE.json's full size (configs[1]: 4,096 concurrent Connect4 games x 800 simulations per move).

The oracle cannot replay 3.3 M simulations per case in seconds, so full-size parity rests on properties that do not depend on
the size: trees are independent (a sample of them is replayed by the oracle bit for bit, and the rest must agree between
the fused and the lock-step pipelines), visit counts conserve the simulation budget (mcts.rs:161-192: every simulation
adds one visit to exactly one root child except the first, which expands the root), a search is a pure function of
(roots, evaluator), and a finished trajectory replays through the oracle's rules move by move.
"""
import hashlib

import numpy as np
import pytest

import selfplay_b200 as S
from helpers import synthetic_roots
from oracle import pyoracle as O
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

pytestmark = pytest.mark.gpu
G, SIMS = 4096, 800


def _digest(e):
    a, c, i, n = e.root_children_all()
    h = hashlib.sha256()
    for arr in (a, c, i, n):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest(), c, n


def _conservation(e, counts, n_children, sims):
    # every root here is an ongoing position, so each tree spent one simulation on expanding its root
    assert (counts.sum(axis=1) == sims - 1).all()
    assert (n_children >= 1).all() and (n_children <= 7).all()
    ctr = e.counters()
    assert ctr["simulations"] == e.G * sims
    assert ctr["evaluations"] + ctr["terminal_leaves"] == e.G * sims
    return ctr


def test_full_size_deterministic_evaluator_sampled_oracle_and_pipeline_agreement():
    sample = list(range(0, G, 171))                       # 24 trees spread over the batch
    f = O.Forest(O.GAME_C4, len(sample))
    digests = []
    for flags in (0, S.FLAG_FORCE_SPLIT, S.FLAG_FORCE_SPLIT | S.FLAG_LOCKSTEP):       # fused, asynchronous, lock-step
        with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_DET, flags=flags) as e:
            roots = synthetic_roots_device(e, G)
            e.reset_games(roots)
            e.reset_counters()
            e.search(SIMS)
            d, counts, n = _digest(e)
            ctr = _conservation(e, counts, n, SIMS)
            digests.append((d, ctr["path_length_sum"], ctr["children_created"], ctr["nodes_live"]))
            if flags == 0:
                want_roots = synthetic_roots(O.GAME_C4, G)      # oracle-built roots == device-built roots
                for g in sample:
                    assert (int(roots[g]["stones"][0]), int(roots[g]["stones"][1])) == tuple(want_roots[g].stones)
                f.reset([want_roots[g] for g in sample])
                f.search(SIMS, O.EVAL_DET)
            for k, g in enumerate(sample):
                assert e.root_children(g) == f.root_children(k), g
                assert e.arena_len(g) == f.arena_len(k)
                assert e.node_stats(g, 0) == f.node_stats(k, 0)
                last = f.arena_len(k) - 1
                assert e.node_stats(g, last) == f.node_stats(k, last)
    assert digests[0] == digests[1] == digests[2]          # all 4,096 trees: fused == asynchronous == lock-step, bit for bit


def test_full_size_network_search_is_conserving_and_reproducible():
    """The tcgen05 evaluator reduces in a fixed order per board, so the result of a search does not depend on how
    leaves were batched: two searches from the same roots through the asynchronous pipeline, the lock-step CUDA-graph
    pipeline and the lock-step direct-launch pipeline give identical visit counts for all 4,096 trees."""
    blob = random_checkpoint(1, 0)
    seen = []
    for flags in (0, S.FLAG_LOCKSTEP, S.FLAG_LOCKSTEP | S.FLAG_NO_GRAPH):
        with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags) as e:
            e.load_weights(blob)
            roots = synthetic_roots_device(e, G)
            for _ in range(2):
                e.reset_games(roots)
                e.reset_counters()
                e.search(SIMS)
                d, counts, n = _digest(e)
                _conservation(e, counts, n, SIMS)
                seen.append(d)
    assert len(set(seen)) == 1


def test_full_size_virtual_loss_16_leaves_conserves_the_budget():
    """configs[3]: 16 leaves in flight per tree (EXTENSION, oracle.cc search_vl).  Virtual loss must be fully reverted:
    the root has seen every simulation, and the children all but the 16 of the first step, which all selected the
    still unexpanded root.  A sample of trees is replayed by the oracle's own K-leaf restatement."""
    K = 16
    sample = list(range(5, G, 341))
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_DET, leaves_per_tree=K) as e:
        roots = synthetic_roots_device(e, G)
        e.reset_games(roots)
        e.reset_counters()
        e.search(SIMS)
        _, counts, n = _digest(e)
        assert (counts.sum(axis=1) == SIMS - K).all()
        ctr = e.counters()
        assert ctr["simulations"] == G * SIMS
        f = O.Forest(O.GAME_C4, len(sample), leaves_per_tree=K)
        f.reset([O.state_from_record(roots[g]) for g in sample])
        f.search(SIMS, O.EVAL_DET)
        for k, g in enumerate(sample):
            assert e.node_stats(g, 0)["visit_count"] == SIMS
            assert e.root_children(g) == f.root_children(k), g
            assert e.arena_len(g) == f.arena_len(k)
            assert e.node_stats(g, 0) == f.node_stats(k, 0)


def test_full_size_selfplay_trajectories_replay_through_the_oracle_rules():
    """4,096 games played to the end on the device (greedy last-max rule, subtree reuse).  Every emitted trajectory
    must replay through the oracle's Connect4 rules: position k+1 = next_state(position k, last-max action of the stored
    visit counts), the game ends exactly at the last stored ply, and the value targets follow learner_concurrent.rs:214-226."""
    sims = 48
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_DET) as e:
        roots = synthetic_roots_device(e, G)
        e.reset_games(roots)
        finished = 0
        for _ in range(42):
            e.search(sims)
            finished += e.selfplay_step(S.MOVE_GREEDY_LAST_MAX)
            if finished == G:
                break
        assert finished == G
        pos, gids = e.drain_trajectories()
    assert sorted(set(int(g) for g in gids)) == list(range(G))
    bounds = np.flatnonzero(np.diff(gids.astype(np.int64))) + 1
    games = np.split(np.arange(len(pos)), bounds)
    assert len(games) == G
    for idx in games:
        g = int(gids[idx[0]])
        s = O.state_from_record(roots[g])
        final_mover = None
        for k, i in enumerate(idx):
            r = pos[i]
            assert int(r["ply"]) == k
            assert (int(r["stones"][0]), int(r["stones"][1]), int(r["current_player"])) == (s.stones[0], s.stones[1], s.current_player)
            counts = [int(c) for c in r["visit_counts"]]
            legal = O.valid_actions(O.GAME_C4, s)
            assert all(c == 0 for a, c in enumerate(counts) if a not in legal)
            assert sum(counts) >= sims - 1                      # the carried subtree only adds visits
            best = max(legal, key=lambda a: (counts[a], a))     # main.rs:108-112 over children in action order
            final_mover = s.current_player
            s = O.next_state(O.GAME_C4, s, best)
            assert (s.status == O.ONGOING) == (k + 1 < len(idx))
        value = 1 if s.status == O.WON else 0                    # from the side that made the last move
        for i in idx:
            want = value if int(pos[i]["current_player"]) == final_mover else -value
            assert int(pos[i]["outcome"]) == want


def test_full_size_tictactoe_sampled_oracle_and_pipeline_agreement():
    """configs[0] at scale: 4,096 tic-tac-toe games x 200 simulations — the trees exhaust the game below most roots, so
    terminal-leaf backups dominate; sampled trees bit-exact vs the oracle, fused == lock-step for all of them."""
    sims = 200
    sample = list(range(3, G, 293))
    f = O.Forest(O.GAME_TTT, len(sample))
    digests = []
    for flags in (0, S.FLAG_FORCE_SPLIT, S.FLAG_FORCE_SPLIT | S.FLAG_LOCKSTEP):
        with S.Engine(game=S.GAME_TTT, num_games=G, evaluator=S.EVAL_DET, flags=flags) as e:
            roots = synthetic_roots_device(e, G, max_ply=5)
            e.reset_games(roots)
            e.reset_counters()
            e.search(sims)
            d, counts, n = _digest(e)
            assert (counts.sum(axis=1) == sims - 1).all()
            ctr = e.counters()
            assert ctr["simulations"] == G * sims and ctr["evaluations"] + ctr["terminal_leaves"] == G * sims
            digests.append((d, ctr["path_length_sum"], ctr["children_created"], ctr["terminal_leaves"]))
            if flags == 0:
                f.reset([O.state_from_record(roots[g]) for g in sample])
                f.search(sims, O.EVAL_DET)
            for k, g in enumerate(sample):
                assert e.root_children(g) == f.root_children(k), g
                assert e.arena_len(g) == f.arena_len(k)
                assert e.node_stats(g, 0) == f.node_stats(k, 0)
    assert digests[0] == digests[1] == digests[2]


def test_full_size_network_with_16_leaves_in_flight_conserves_the_budget():
    """configs[3] with the real evaluator: 4,096 trees x 16 leaves per step = up to 65,536 positions per evaluator launch."""
    K = 16
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, leaves_per_tree=K) as e:
        e.load_weights(random_checkpoint(1, 0))
        roots = synthetic_roots_device(e, G)
        seen = []
        for _ in range(2):
            e.reset_games(roots)
            e.reset_counters()
            e.search(SIMS)
            d, counts, n = _digest(e)
            assert (counts.sum(axis=1) == SIMS - K).all()
            ctr = e.counters()
            assert ctr["simulations"] == G * SIMS
            assert e.node_stats(17, 0)["visit_count"] == SIMS
            seen.append(d)
        assert seen[0] == seen[1]
