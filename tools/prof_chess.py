"""Driver for ncu / timing of the chess network: evaluates a batch of synthetic chess positions a few times.
usage: python tools/prof_chess.py [positions] [repeats]
Under ncu: ncu --set full --clock-control none --import-source on -k regex:k_conv -s 3 -c 1 -o gpurun_out/r02_chess_conv python tools/prof_chess.py 4096 1"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_chess_roots_device
from selfplay_b200.weights_init import random_chess_checkpoint

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with S.ChessEngine(num_games=n, evaluator=S.EVAL_NET, max_nodes_per_tree=1024) as e:
    e.load_weights(random_chess_checkpoint(0))
    st, hist = synthetic_chess_roots_device(e, n)
    e.reset_games(st, hist)
    for _ in range(reps):
        t0 = time.perf_counter()
        e.search(1)                                   # one simulation per tree = one network evaluation of n root positions
        dt = time.perf_counter() - t0
        print("search(1): %.3f ms wall, %.3f ms device" % (dt * 1e3, e.last_search_ms()), flush=True)
        e.reset_games(st, hist)
    e.search(1)
    ms, npos, fl, fp = e.time_conv(iters=20)
    print("k_conv<8,9,256>: %.3f ms per launch, %d positions, %.1f TFLOP/s (algorithmic), whole net %.3f GFLOP per position" % (
        ms, npos, fl / (ms * 1e-3) / 1e12, fp / 1e9), flush=True)
