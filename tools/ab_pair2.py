"""End-to-end A/B: 800-simulation search of 4,096 games with the default evaluator and the CTA-pair variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
G, sims = 4096, 800
for name, flags in (("default", 0), ("cta-pair", S.FLAG_EVAL_PAIR2), ("default", 0), ("cta-pair", S.FLAG_EVAL_PAIR2)):
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags) as e:
        e.load_weights(random_checkpoint(1, 0))
        roots = synthetic_roots_device(e, G)
        best = 1e9
        for _ in range(3):
            e.reset_games(roots); e.search(sims); best = min(best, e.last_search_timing()[0])
        print("%-9s %.2f ms  %.2f M sims/s" % (name, best, G * sims / best / 1e3), flush=True)
