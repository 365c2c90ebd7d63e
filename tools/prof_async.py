"""Driver for ncu: one warm-up search and one profiled search at the bench workload.
usage: python tools/prof_async.py [async|lockstep] [games] [sims]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

mode = sys.argv[1] if len(sys.argv) > 1 else "async"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 800
flags = 0 if mode == "async" else (S.FLAG_LOCKSTEP | S.FLAG_NO_GRAPH)
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags, max_nodes_per_tree=8192) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    for _ in range(2):
        e.reset_games(roots)
        e.search(sims)
        print(mode, "search ms", e.last_search_timing()[0], flush=True)
