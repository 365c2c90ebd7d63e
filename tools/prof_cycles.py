import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=S.FLAG_NO_GRAPH) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    e.reset_games(roots)
    e.search(sims)
    L = S.load_library()
    lay = np.zeros((148, 24), np.uint64)
    L.spb_debug_eval_profile_layers(C.c_void_p(lay.ctypes.data), 148, 1)
    for dbg in (0,):
        L.spb_debug_set(dbg)
        ms, n, fl = e.time_evaluator(10)
        buf = np.zeros((148, 8), np.uint64)
        L.spb_debug_eval_profile(C.c_void_p(buf.ctypes.data), 148)
        b = buf.astype(np.float64)
        print("dbg %d: eval ms %.3f positions %d | MMA warp total %.0f, wait epi %.0f (%.0f%%), wait weights %.0f (%.0f%%) | epi total %.0f wait MMA %.0f (%.0f%%)" % (
            dbg, ms, n, b[:, 0].mean(), b[:, 1].mean(), 100 * b[:, 1].mean() / b[:, 0].mean(), b[:, 5].mean(), 100 * b[:, 5].mean() / b[:, 0].mean(),
            b[:, 3].mean(), b[:, 4].mean(), 100 * b[:, 4].mean() / b[:, 3].mean()), flush=True)
        L.spb_debug_eval_profile_layers(C.c_void_p(lay.ctypes.data), 148, 1)
        lm = lay.astype(np.float64).mean(0) / 12.0   # 12 launches (2 warm-up + 10)
        print("per-launch MMA-warp wait for activations by layer:", [int(x) for x in lm[:10]])
        print("per-launch MMA-warp wait for weights by layer:    ", [int(x) for x in lm[10:20]])
