"""A/B timing of evaluator build variants: SPB_LIB=<path to .so> python tools/variant_bench.py [games] [sims]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200.engine as E
if os.environ.get("SPB_LIB"):
    E._LIB = os.path.abspath(os.environ["SPB_LIB"])
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 200
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    best = 1e9
    for _ in range(4):
        e.reset_games(roots)
        e.search(sims)
        best = min(best, e.last_search_timing()[0])
    ms, n, fl = e.time_evaluator(30)
    import hashlib, numpy as np
    a, c, i, nn = e.root_children_all()
    print("%-28s search %.2f ms (%.1f us/step, %.2f M sims/s) | evaluator %.1f us for %d positions = %.0f TFLOP/s | counts sha %s" % (
        os.path.basename(os.environ.get("SPB_LIB", "default")), best, best * 1e3 / sims, G * sims / best / 1e3, ms * 1e3, n, fl * n / ms / 1e9,
        hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()[:12]), flush=True)
