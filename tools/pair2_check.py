"""Correctness + timing of the CTA-pair evaluator (SPB_FLAG_EVAL_PAIR2) against the default kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

for game, tag in ((S.GAME_C4, "c4"), (S.GAME_TTT, "ttt")):
    blob = random_checkpoint(game, 0)
    for n in (1, 2, 9, 19, 37, 700, 3000):
        outs = {}
        for flags, name in ((0, "pair"), (S.FLAG_EVAL_PAIR2, "pair2")):
            with S.Engine(game=game, num_games=64, evaluator=S.EVAL_NET, flags=flags) as e:
                e.load_weights(blob)
                roots = synthetic_roots_device(e, n, max_ply=21 if game == S.GAME_C4 else 5)
                outs[name] = e.predict(roots, want_logits=True)
        dl = np.abs(outs["pair2"][2] - outs["pair"][2]).max()
        dv = np.abs(outs["pair2"][1] - outs["pair"][1]).max()
        print("%s n=%4d pair2 vs pair: max |dlogit| %.3e  max |dvalue| %.3e" % (tag, n, dl, dv), flush=True)

G, sims = 4096, 200
for flags, name in ((0, "pair"), (S.FLAG_EVAL_PAIR2, "pair2")):
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags) as e:
        e.load_weights(random_checkpoint(1, 0))
        roots = synthetic_roots_device(e, G)
        best = 1e9
        for _ in range(3):
            e.reset_games(roots)
            e.search(sims)
            best = min(best, e.last_search_timing()[0])
        ms, n, fl = e.time_evaluator(30)
        print("%-5s search %.2f ms (%.1f us/step, %.2f M sims/s) | evaluator %.1f us for %d positions = %.0f TFLOP/s" % (
            name, best, best * 1e3 / sims, G * sims / best / 1e3, ms * 1e3, n, fl * n / ms / 1e9), flush=True)
