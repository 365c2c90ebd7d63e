"""Regenerates the known-answer vectors of tests/golden/survey_kats.json from the pure-Python restatement (tests/pyref.py)
and writes them to tests/golden/pyref_kats.json.  tests/test_oracle.py::test_committed_golden_vectors_come_from_the_generator
checks that both files agree key by key, so the golden fixture is the output of a committed script (and equal to the vectors
SURVEY.md §8(c) lists), not a transcription.  Not produced by the reference (no rustc here): parity unpinned, DESIGN.md §2.
Run from the repo root: python tools/gen_golden.py        (pure Python: a few minutes)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pyref as R   # noqa: E402


def counts(game_cls, sims, ev):
    t = R.Tree(game_cls())
    R.search([t], sims, ev)
    return t


def path_stats(game_cls, sims, ev):
    """Path-length sum / max and terminal leaves of one search, counted beside pyref.search's loop."""
    t = R.Tree(game_cls())
    psum = pmax = term = 0
    for _ in range(sims):
        nid, d = 0, 0
        while t.arena[nid]["children"]:
            nid = t.select(nid)
            d += 1
        psum += d
        pmax = max(pmax, d)
        st = t.arena[nid]["state"]
        if st.status == 2:
            t.backprop(nid, R.F(-1)); term += 1
        elif st.status == 1:
            t.backprop(nid, R.F(0)); term += 1
        else:
            probs, v = ev(st)
            t.expand(nid, R.mask(st, probs))
            t.backprop(nid, v)
    return t, psum, pmax, term


def greedy(game_cls, sims, ev, max_plies=64):
    """main.rs:106-114: search, pick the LAST child with the maximal visit count, use_subtree, until the game ends."""
    t = R.Tree(game_cls())
    actions, arena_sizes, last, first = [], [], None, None
    while t.arena[0]["state"].status == 0 and len(actions) < max_plies:
        R.search([t], sims, ev)
        arena_sizes.append(len(t.arena))
        ch = t.arena[0]["children"]
        last = [t.arena[c]["N"] for c in ch]
        first = first or last
        best = max(range(len(ch)), key=lambda i: (last[i], i))
        actions.append(t.arena[ch[best]]["action"])
        t.use_subtree(ch[best])
    return dict(actions=actions, final_status=t.arena[0]["state"].status, last_counts=last, arena_sizes=arena_sizes, first_counts=first)


def main():
    out = {"_source": "tools/gen_golden.py: tests/pyref.py (pure-Python restatement of src/mcts.rs + src/game/*.rs). Fresh tree, c=2.0, counts in child order."}
    st = R.C4().stones()
    out["det_hash_empty_c4"] = "0x%016x" % R.splitmix64(st[0] ^ R.splitmix64(st[1]))
    out["ttt_uniform"] = {str(s): counts(R.TTT, s, R.uniform_eval).root_counts() for s in (2, 10, 11, 100, 600)}
    out["ttt_uniform_arena_600"] = len(counts(R.TTT, 600, R.uniform_eval).arena)
    out["c4_uniform"] = {str(s): counts(R.C4, s, R.uniform_eval).root_counts() for s in (8, 9, 100)}
    out["c4_uniform_arena"] = {"100": len(counts(R.C4, 100, R.uniform_eval).arena)}
    t, psum, _, term = path_stats(R.C4, 800, R.uniform_eval)
    out["c4_uniform"]["800"] = t.root_counts()
    out["c4_uniform_arena"]["800"] = len(t.arena)
    out["c4_uniform_800_path_sum"] = psum
    assert term == 0
    out["c4_uniform_greedy_100_actions"] = greedy(R.C4, 100, R.uniform_eval)["actions"]
    out["c4_det"], out["c4_det_root_w"], out["c4_det_arena"] = {}, {}, {}
    for s in (1, 8, 100):
        t = counts(R.C4, s, R.det_eval)
        out["c4_det_root_w"][str(s)] = float(t.arena[0]["W"])
        if s > 1:
            out["c4_det"][str(s)] = t.root_counts()
        if s == 100:
            out["c4_det_arena"]["100"] = len(t.arena)
    t, psum, pmax, term = path_stats(R.C4, 800, R.det_eval)
    out["c4_det"]["800"] = t.root_counts()
    out["c4_det_root_w"]["800"] = float(t.arena[0]["W"])
    out["c4_det_arena"]["800"] = len(t.arena)
    out["c4_det_800_path_sum"], out["c4_det_800_max_path"], out["c4_det_800_terminal_leaves"] = psum, pmax, term
    g = greedy(R.C4, 800, R.det_eval)
    out["c4_det_greedy_800"] = {k: g[k] for k in ("actions", "final_status", "last_counts", "arena_sizes")}
    g = greedy(R.C4, 200, R.det_eval)
    out["c4_det_greedy_200"] = {"actions": g["actions"], "first_counts": g["first_counts"]}
    out["c4_antidiagonal_moves"] = [3, 2, 2, 1, 1, 0, 1, 0, 0, 6, 0]
    s = R.C4()
    for a in out["c4_antidiagonal_moves"]:
        s = s.next_state(a)
    assert s.status == 0                                              # connect_four.rs:163-176: the anti-diagonal is not checked
    with open(os.path.join(ROOT, "tests", "golden", "pyref_kats.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote tests/golden/pyref_kats.json")


if __name__ == "__main__":
    main()
