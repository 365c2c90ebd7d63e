"""Cycle attribution of the default tcgen05 evaluator (build with `make -C self-play-ai_b200/csrc PROFILE=1`)."""
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
    L.spb_debug_eval_profile_layers_v1(C.c_void_p(lay.ctypes.data), 148, 1)
    iters = 10
    ms, n, fl = e.time_evaluator(iters)
    buf = np.zeros((148, 8), np.uint64)
    L.spb_debug_eval_profile_v1(C.c_void_p(buf.ctypes.data), 148)
    L.spb_debug_eval_profile_layers_v1(C.c_void_p(lay.ctypes.data), 148, 1)
    b = buf.astype(np.float64)
    print("eval ms %.3f positions %d (%.0f TFLOP/s) | MMA warp total %.0f cycles, wait act %.0f (%.0f%%), wait weights %.0f (%.0f%%) | epi total %.0f wait MMA %.0f (%.0f%%) | batches %.1f" % (
        ms, n, fl * n / ms / 1e9, b[:, 0].mean(), b[:, 1].mean(), 100 * b[:, 1].mean() / b[:, 0].mean(), b[:, 5].mean(), 100 * b[:, 5].mean() / b[:, 0].mean(),
        b[:, 3].mean(), b[:, 4].mean(), 100 * b[:, 4].mean() / b[:, 3].mean(), b[:, 2].mean()))
    l = lay.astype(np.float64).mean(axis=0) / (iters + 1)
    print("per-launch MMA-warp wait for activations by layer:", [int(x) for x in l[:10]])
    print("per-launch MMA-warp wait for weights by layer:    ", [int(x) for x in l[10:20]])
