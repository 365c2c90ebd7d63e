"""Timeline of evaluator CTA 0 inside the asynchronous pipeline (trace build): a window of 12 consecutive batches in the
steady state — when the MMA warp issues each (batch, layer), when the epilogue finishes it, when the stager claims and
stages.  tools/build_variant.sh trace -DSPB_TRACE; SPB_LIB=variants/lib_trace.so python tools/trace_async.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200.engine as E
if os.environ.get("SPB_LIB"):
    E._LIB = os.path.abspath(os.environ["SPB_LIB"])
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

G = 4096
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, max_nodes_per_tree=8192) as e:
    e.load_weights(random_checkpoint(1, 0))
    roots = synthetic_roots_device(e, G)
    e.reset_games(roots)
    e.search(800)
    print("search %.2f ms" % e.last_search_timing()[0])
    L = S.load_library()
    tr = np.zeros((4, 512), np.uint64)
    L.spb_debug_trace_v2(C.c_void_p(tr.ctypes.data), 0)
    st = np.zeros((16, 4), np.uint64)
    L.spb_debug_trace_stager(C.c_void_p(st.ctypes.data))
    tr = tr.astype(np.int64)
    st = st.astype(np.int64)
    t0 = tr[0][tr[0] > 0].min()
    prev = None
    for b in range(12):
        i0 = (b * 10) * 4
        if tr[0][i0] == 0:
            continue
        starts = [tr[0][(b * 10 + l) * 4] - t0 for l in range(10)]
        mma_end = [max(tr[1][(b * 10 + l) * 4:(b * 10 + l) * 4 + 4]) - t0 for l in range(10)]
        epi_end = [max(tr[3][(b * 10 + l) * 4:(b * 10 + l) * 4 + 4]) - t0 for l in range(10)]
        per = "" if prev is None else " period %6d" % (starts[0] - prev)
        prev = starts[0]
        print("b%02d stem issue @%7d%s | layer issue starts (rel) %s" % (b, starts[0], per, " ".join("%5d" % (x - starts[0]) for x in starts)))
        print("      mma issue ends (rel)  %s" % " ".join("%5d" % (x - starts[0]) for x in mma_end))
        print("      epilogue ends (rel)   %s" % " ".join("%5d" % (x - starts[0]) for x in epi_end))
        print("      stager: claim %d..%d  act0_free %d  staged %d (rel to this batch's stem issue)" % tuple(int(x - t0 - starts[0]) for x in st[b]))
    b = 5
    base = tr[0][(b * 10) * 4]
    print("per tile of batch %d (cycles from its stem issue): mma issue begin..end, epilogue begin..end" % b)
    for l in range(10):
        row = []
        for t in range(4):
            i = (b * 10 + l) * 4 + t
            row.append("t%d mma %6d..%6d epi %6d..%6d" % (t, tr[0][i] - base, tr[1][i] - base, tr[2][i] - base, tr[3][i] - base))
        print("l%d | " % l + " | ".join(row))
    s = tr[3][480:486]
    print("linear_heads(first traced batch) warp 2: start %d | weights +%d | board loop +%d | cross-warp sum +%d | softmax +%d | sync +%d" % (
        s[0] - t0, s[1] - s[0], s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4]))
    for w, name in ((0, "first epilogue warp (half 0)"), (8, "fifth epilogue warp (half 1)")):
        s = tr[3][496 + w:496 + w + 6]
        if s[0]:
            print("conv epilogue of (batch +2, layer 4, tile 1), %s: wait acc_full +%d | tcgen05.ld x6 + wait + drained +%d | 64 shuffles + adds +%d | skip, bias, ReLU, 4 stores +%d | fence + arrive +%d" % (
                name, s[1] - s[0], s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4]))
