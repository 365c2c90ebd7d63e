"""Determinism soak: repeated 800-simulation searches of 4,096 games from the same roots must give the same visit counts
every time (a rare race in the evaluator's mbarrier pipeline would show up as a different digest)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for name, flags in (("default", 0), ("default no-graph", S.FLAG_NO_GRAPH), ("cta-pair", S.FLAG_EVAL_PAIR2), ("first kernel", S.FLAG_EVAL_V1)):
    with S.Engine(game=S.GAME_C4, num_games=4096, evaluator=S.EVAL_NET, flags=flags) as e:
        e.load_weights(random_checkpoint(1, 0))
        roots = synthetic_roots_device(e, 4096)
        seen = set()
        for r in range(reps):
            e.reset_games(roots)
            e.search(800)
            a, c, i, n = e.root_children_all()
            seen.add(hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()[:16])
        print("%-18s %d searches -> %d distinct digest(s): %s" % (name, reps, len(seen), sorted(seen)), flush=True)
