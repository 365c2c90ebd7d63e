"""Evaluator time vs number of full batches per CTA (n = 148 * 9 * k positions): T(k) = F + k * B."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200.engine as E
if os.environ.get("SPB_LIB"):
    E._LIB = os.path.abspath(os.environ["SPB_LIB"])
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
blob = random_checkpoint(1, 0)
for name, flags in (("pair (default)", 0), ("pair2", S.FLAG_EVAL_PAIR2)):
    res = []
    for n in (148 * 2, 148 * 4, 148 * 9, 148 * 11, 148 * 13, 148 * 18, 148 * 21, 148 * 27, 148 * 36, 148 * 72):
        with S.Engine(game=S.GAME_C4, num_games=n, evaluator=S.EVAL_NET, flags=flags | S.FLAG_NO_GRAPH) as e:
            e.load_weights(blob)
            e.reset_games(synthetic_roots_device(e, n))
            e.search(1)
            ms, npos, fl = e.time_evaluator(30)
            assert npos == n
            res.append((n // 148, ms * 1e3, fl * n / ms / 1e9))
    print(name, " | ".join("%d boards/CTA: %.1f us (%.0f TF)" % r for r in res), flush=True)
