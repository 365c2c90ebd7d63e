"""Timeline of CTA 0 of the evaluator on a static list (profile build): when the MMA warp issues each (batch, layer, tile) and when
the epilogue handles it.  SPB_LIB=variants/lib_prof.so python tools/trace_v2.py"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200.engine as E
if os.environ.get("SPB_LIB"):
    E._LIB = os.path.abspath(os.environ["SPB_LIB"])
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
G = 4096
VER = "v2"
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=S.FLAG_NO_GRAPH | S.FLAG_LOCKSTEP) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    e.reset_games(roots)
    e.search(60)
    L = S.load_library()
    tr = np.zeros((4, 512), np.uint64)
    getattr(L, "spb_debug_trace_" + VER)(C.c_void_p(tr.ctypes.data), 1)
    ms, n, fl = e.time_evaluator(3)
    getattr(L, "spb_debug_trace_" + VER)(C.c_void_p(tr.ctypes.data), 0)
    tr = tr.astype(np.int64)
    t0 = tr[0][tr[0] > 0].min()
    print("evaluator %.1f us, %d positions; times in cycles from the first MMA issue of CTA 0" % (ms * 1e3, n))
    for b in range(4):
        for l in range(10):
            row = []
            for t in range(4):
                i = (b * 10 + l) * 4 + t
                if tr[0][i] == 0:
                    continue
                row.append("t%d mma %6d..%6d epi %6d..%6d" % (t, tr[0][i] - t0, tr[1][i] - t0, tr[2][i] - t0, tr[3][i] - t0))
            if row:
                print("b%d l%d | " % (b, l) + " | ".join(row))
    for w, off in (("warp 2 (output 0)", 480), ("warp 9 (value)", 488)):
        s = tr[3][off:off + 6]
        print("linear_heads(b0) %s: start %d | weights +%d | board loop +%d | cross-warp sum +%d | softmax +%d | sync +%d" % (
            w, s[0] - t0, s[1] - s[0], s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4]))
