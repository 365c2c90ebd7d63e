"""Cycle attribution of the FIRST tcgen05 evaluator (evaluator_umma_v1.cu, SPB_FLAG_EVAL_V1; build with
`make -B -C self-play-ai_b200/csrc PROFILE=1`).  The default kernel is profiled with tools/trace_v2.py (VARIANT=-DSPB_TRACE)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import selfplay_b200.engine as E
if os.environ.get("SPB_LIB"):
    E._LIB = os.path.abspath(os.environ["SPB_LIB"])
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint
VER = "v1"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=S.FLAG_NO_GRAPH | S.FLAG_EVAL_V1) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    e.reset_games(roots)
    e.search(sims)
    L = S.load_library()
    lay = np.zeros((148, 24), np.uint64)
    getattr(L, "spb_debug_eval_profile_layers_" + VER)(C.c_void_p(lay.ctypes.data), 148, 1)
    iters = 10
    getattr(L, "spb_debug_set_" + VER)(int(os.environ.get("SPB_DBG", "0")))
    ms, n, fl = e.time_evaluator(iters)
    buf = np.zeros((148, 8), np.uint64)
    getattr(L, "spb_debug_eval_profile_" + VER)(C.c_void_p(buf.ctypes.data), 148)
    getattr(L, "spb_debug_eval_profile_layers_" + VER)(C.c_void_p(lay.ctypes.data), 148, 1)
    b = buf.astype(np.float64)
    print("dbg=%s " % os.environ.get("SPB_DBG", "0") + "eval ms %.3f positions %d (%.0f TFLOP/s) | MMA warp total %.0f cycles, wait act %.0f (%.0f%%), wait weights %.0f (%.0f%%) | epi total %.0f wait MMA %.0f (%.0f%%) | batches %.1f" % (
        ms, n, fl * n / ms / 1e9, b[:, 0].mean(), b[:, 1].mean(), 100 * b[:, 1].mean() / b[:, 0].mean(), b[:, 5].mean(), 100 * b[:, 5].mean() / b[:, 0].mean(),
        b[:, 3].mean(), b[:, 4].mean(), 100 * b[:, 4].mean() / b[:, 3].mean(), b[:, 2].mean()))
    print("conv-layer epilogue body: %.0f cycles per tile (warp 2 lane 0), %.0f bodies per launch" % (b[:, 6].sum() / max(1, b[:, 7].sum()), b[:, 7].mean()))
    l = lay.astype(np.float64).mean(axis=0) / (iters + 1)
    print("per-launch MMA-warp wait for activations by layer:", [int(x) for x in l[:10]])
    print("per-launch MMA-warp wait for weights by layer:    ", [int(x) for x in l[10:20]])
    L2 = lay.astype(np.float64).sum(axis=0)
    if L2[23] > 0:
        print("residual-layer tile issue cycles (4-tile batches, weight waits excluded): first %.0f  middle %.0f  last %.0f  => %.1f / %.1f / %.1f cycles per MMA" % (
            L2[20] / L2[23], L2[21] / (2 * L2[23]), L2[22] / L2[23], L2[20] / L2[23] / 36, L2[21] / (2 * L2[23]) / 36, L2[22] / L2[23] / 36))
