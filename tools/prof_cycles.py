import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=S.FLAG_NO_GRAPH | S.FLAG_EVAL_DX) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    e.reset_games(roots)
    e.search(sims)
    L = S.load_library()
    for dbg in (0,):
      L.spb_debug_set(dbg)
      ms, n, fl = e.time_evaluator(10)
      print('dbg', dbg)
      buf = np.zeros((148, 8), np.uint64)
      L.spb_debug_eval_profile(C.c_void_p(buf.ctypes.data), 148)
      b = buf.astype(np.float64)
      print("eval ms %.3f positions %d | MMA warp total %.0f, wait epi %.0f (%.0f%%), wait weights %.0f (%.0f%%) | epi total %.0f wait MMA %.0f (%.0f%%)" % (
        ms, n, b[:, 0].mean(), b[:, 1].mean(), 100 * b[:, 1].mean() / b[:, 0].mean(), b[:, 5].mean(), 100 * b[:, 5].mean() / b[:, 0].mean(),
        b[:, 3].mean(), b[:, 4].mean(), 100 * b[:, 4].mean() / b[:, 3].mean()), flush=True)
      print("epilogue layer bodies total %.0f ; head+stage+finish total %.0f ; batches %.0f" % (b[:, 6].mean(), b[:, 7].mean(), b[:, 2].mean()))
