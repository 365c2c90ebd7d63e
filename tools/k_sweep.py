import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
G, sims = 4096, 800
blob = random_checkpoint(1, 0)
for K in (1, 2, 4, 8, 16):
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, leaves_per_tree=K, max_nodes_per_tree=16384) as e:
        e.load_weights(blob)
        roots = synthetic_roots_device(e, G)
        for rep in range(2):
            e.reset_games(roots); e.reset_counters()
            e.search(sims)
        ms = e.last_search_timing()[0]
        c = e.counters()
        ev_ms, n, fl = e.time_evaluator(5)
        print("K=%2d: %.1f ms per 800-sim search, %.1f M sims/s, evals %d (%.0f%% of sims), last eval launch %.0f us for %d positions = %.0f TFLOP/s" % (
            K, ms, G * sims / ms / 1e3, c["evaluations"], 100.0 * c["evaluations"] / c["simulations"], ev_ms * 1e3, n, fl * n / ev_ms / 1e9), flush=True)
