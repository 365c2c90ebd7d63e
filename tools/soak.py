"""Determinism soak: repeated searches from the same roots must give the same visit counts every time, in the asynchronous
and the lock-step pipeline alike (a rare race in the evaluator's mbarrier pipeline or in the rings would show up as a
different digest).  usage: python tools/soak.py [repetitions]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
CASES = (("c4 4096 x 800", S.GAME_C4, 4096, 800), ("c4 300 x 200 (ragged batches)", S.GAME_C4, 300, 200), ("c4 20 x 100 (few leaves in flight)", S.GAME_C4, 20, 100),
         ("ttt 4096 x 300", S.GAME_TTT, 4096, 300))
for name, game, G, sims in CASES:
    digests = {}
    for mode, flags in (("async", 0), ("lock-step", S.FLAG_LOCKSTEP)):
        with S.Engine(game=game, num_games=G, evaluator=S.EVAL_NET, flags=flags) as e:
            e.load_weights(random_checkpoint(game, 0))
            roots = synthetic_roots_device(e, G)
            seen = set()
            for r in range(reps if mode == "async" else max(2, reps // 5)):
                e.reset_games(roots)
                e.search(sims)
                a, c, i, n = e.root_children_all()
                seen.add(hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()[:16])
            digests[mode] = seen
    ok = len(digests["async"]) == 1 and digests["async"] == digests["lock-step"]
    print("%-36s async %d distinct digest(s), lock-step %d, equal: %s  %s" % (name, len(digests["async"]), len(digests["lock-step"]), ok, sorted(digests["async"])), flush=True)
