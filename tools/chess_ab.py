"""A/B of the chess network pipeline: two half-loops on two streams (default) against one loop over all trees
(SPB_FLAG_LOCKSTEP), with a digest of every root's children.  usage: python tools/chess_ab.py [games] [sims]"""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200.engine as E
if os.environ.get("SPB_LIB"):
    E._LIB = os.path.abspath(os.environ["SPB_LIB"])
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_chess_roots_device
from selfplay_b200.weights_init import random_chess_checkpoint

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 24
digests = []
for tag, flags in (("two half-loops", 0), ("one loop", S.FLAG_LOCKSTEP)):
    with S.ChessEngine(num_games=n, evaluator=S.EVAL_NET, max_nodes_per_tree=(0 if sims > 60 else 2048), flags=flags) as e:
        e.load_weights(random_chess_checkpoint(0))
        st, hist = synthetic_chess_roots_device(e, n)
        ms = []
        for rep in range(3):
            e.reset_games(st, hist)
            t0 = time.perf_counter()
            e.search(sims)
            ms.append("%.1f/%.1f" % ((time.perf_counter() - t0) * 1e3, e.last_search_ms()))
        mv, cnt, ids, nn = e.root_children_all()
        h = hashlib.sha256(mv.tobytes() + cnt.tobytes() + nn.tobytes()).hexdigest()[:16]
        digests.append(h)
        best = min(float(x.split("/")[1]) for x in ms)
        print("%-15s wall/device ms %s  best %.2f ms per simulation step = %.0f k sims/s  digest %s" % (tag, ms, best / sims, n * sims / best, h), flush=True)
print("identical results:", digests[0] == digests[1])
