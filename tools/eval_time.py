import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for flags, tag in ((0, "default"), (S.FLAG_EVAL_DX, "dx-sharing")):
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags | S.FLAG_NO_GRAPH) as e:
        e.load_weights(random_checkpoint(1, 0))
        roots = synthetic_roots_device(e, G)
        e.reset_games(roots)
        e.search(sims)
        ms, n, fl = e.time_evaluator(20)
        print("%s: search %.2f ms (%.1f us/step) | evaluator %.1f us for %d positions = %.0f TFLOP/s" % (tag, e.last_search_timing()[0], e.last_search_timing()[0] * 1e3 / sims, ms * 1e3, n, fl * n / ms / 1e9), flush=True)
