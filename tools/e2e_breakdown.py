"""Wall-clock breakdown of one end-to-end step of bench.py (host buffers through the C ABI): reset_games / search / root_children_all."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

G, sims = 4096, 800
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, max_nodes_per_tree=8192) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    for rep in range(4):
        t0 = time.perf_counter(); e.reset_games(roots)
        t1 = time.perf_counter(); e.search(sims)
        t2 = time.perf_counter(); out = e.root_children_all()
        t3 = time.perf_counter()
        print("reset %.3f ms | search wall %.3f ms, device %.3f ms | root_children_all %.3f ms" % (
            (t1 - t0) * 1e3, (t2 - t1) * 1e3, e.last_search_timing()[0], (t3 - t2) * 1e3), flush=True)
