"""Kernel-to-kernel timeline of the lock-step pipeline (trace build: make -B VARIANT=-DSPB_TRACE OUT=...): globaltimer stamps of
tree-step block 0 (entry, after griddepcontrol.wait) and evaluator CTA 0 (entry, after wait, exit) for consecutive steps."""
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
flags = int(os.environ.get("SPB_FLAGS", "0"))
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    e.reset_games(roots)
    e.search(400)
    L = S.load_library()
    tt = np.zeros((64, 8), np.uint64); et = np.zeros((64, 4), np.uint64)
    L.spb_debug_pdl_trace(C.c_void_p(tt.ctypes.data)); L.spb_debug_eval_times_v2(C.c_void_p(et.ctypes.data))
    e.search(40)
    print("search %.2f ms for 40 steps = %.1f us/step" % (e.last_search_timing()[0], e.last_search_timing()[0] * 1e3 / 40))
    L.spb_debug_pdl_trace(C.c_void_p(tt.ctypes.data)); L.spb_debug_eval_times_v2(C.c_void_p(et.ctypes.data))
    tt = tt.astype(np.int64); et = et.astype(np.int64)
    t0 = tt[10][0]
    for i in range(10, 18):
        print("step %2d: tree entry %8d wait-done %8d | eval entry %8d wait-done %8d exit %8d   (ns)" % (
            i, tt[i][0] - t0, tt[i][1] - t0, et[i][0] - t0, et[i][1] - t0, et[i][2] - t0))
    wt = np.zeros((G, 6), np.uint64)
    L.spb_debug_warp_trace(C.c_void_p(wt.ctypes.data), G)
    wt = wt.astype(np.int64)
    base = wt[:, 1].min()
    def pct(a): return "min %6d p50 %6d p90 %6d p99 %6d max %6d" % (a.min(), np.percentile(a, 50), np.percentile(a, 90), np.percentile(a, 99), a.max())
    print("last tree step, per tree (ns): block entry -> wait done   ", pct(wt[:, 1] - wt[:, 0]))
    print("  finish (expand + backup)                                 ", pct(wt[:, 2] - wt[:, 1]))
    print("  descend                                                  ", pct(wt[:, 3] - wt[:, 2]))
    print("  terminal backup / path store / list append               ", pct(wt[:, 4] - wt[:, 3]))
    print("  counters                                                 ", pct(wt[:, 5] - wt[:, 4]))
    print("  warp start after first warp                              ", pct(wt[:, 1] - base))
    print("  warp end after first warp start                          ", pct(wt[:, 5] - base))
