"""Writes tests/golden/chess_kats.json: known answers of the chess search, produced by the pure-Python restatement of
mcts.rs (tests/pyref.py's Tree driven by tests/test_chess_search.py's py_search) — NOT by the C++ oracle they pin, and not by
the reference (no rustc in this image: parity unpinned).  Run from the repo root: python tools/gen_chess_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_chess_search as T   # noqa: E402
from test_chess import play     # noqa: E402

CASES = [
    ("start_det_120", None, "", 120, 1),
    ("start_uniform_100", None, "", 100, 2),
    ("kiwipete_det_60", T.P.KIWIPETE, "", 60, 1),
    ("midgame_det_100", T.MIDGAME, "", 100, 1),
    ("mate_in_one_uniform_150", T.MATE_IN_ONE, "", 150, 2),
    ("after_knight_shuffle_uniform_150", None, "g1f3 g8f6 f3g1 f6g8 g1f3 g8f6 f3g1", 150, 2),
    ("promotion_race_det_80", "8/P6k/8/8/8/8/6Kp/8 w - - 0 1", "", 80, 1),
]


def main():
    out = {"generator": "tools/gen_chess_golden.py (tests/pyref.py Tree + py_search)", "cases": []}
    for name, fen, moves, sims, ev in CASES:
        g = play(moves, fen)
        t = T.py_tree(g)
        T.py_search(t, sims, ev)
        root = t.arena[0]
        out["cases"].append({
            "name": name, "fen": fen, "moves": moves, "sims": sims, "evaluator": ev,
            "root_moves": [t.arena[c]["action"] for c in root["children"]],
            "root_counts": [t.arena[c]["N"] for c in root["children"]],
            "arena_len": len(t.arena), "root_value_sum": float(root["W"]), "det_hash": "%016x" % T.py_det_hash(g),
        })
        print(name, out["cases"][-1]["root_counts"], out["cases"][-1]["arena_len"])
    with open(os.path.join(ROOT, "tests", "golden", "chess_kats.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
