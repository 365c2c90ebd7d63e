"""One worker PROCESS of the CPU arm of bench.py (test infrastructure: see oracle/oracle.h).

The reference runs its self-play as independent worker threads, each with its own `Mcts`, its own copy of the net and
its own batch of trees (ref: src/main.rs:169, src/learner_concurrent.rs:244-290).  Rust threads share nothing on that
path, so the faithful way to time it from Python is one PROCESS per worker: threads of one Python process would
serialise on the interpreter lock inside the evaluator callback (that is why round 1's thread-based arm did not scale
from 16 to 32 host threads).

Protocol (stdin/stdout, one line each way):
    "step <num_searches>"  -> runs Mcts::search(num_searches) on the worker's trees (the oracle port of src/mcts.rs with
                              the torch CPU fp32 net, 1 intra-op thread), answers
                              "<simulations> <evaluations> <terminal leaves> <seconds>"
    "reset"                -> Tree::with_root_state for every tree, answers "ok"
    "quit"
usage: python cpu_worker.py <game> <worker index> <games per worker> <checkpoint file>
"""
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import pyoracle as O  # noqa: E402
from oracle import torch_net  # noqa: E402
from helpers import synthetic_roots  # noqa: E402


def main():
    game, widx, games = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    blob = open(sys.argv[4], "rb").read()
    torch.set_num_threads(1)
    net = torch_net.load_tch_safetensors(blob, game)
    max_ply = 21 if game == O.GAME_C4 else 5
    roots = synthetic_roots(game, games, start=widx * games, max_ply=max_ply)

    def fn(enc):
        p, v, _ = torch_net.forward_probs(net, np.array(enc, copy=True))
        return p, v

    cb = O.make_eval_callback(game, fn)
    forest = O.Forest(game, games)
    forest.reset(roots)
    fn(np.zeros((games, 3) + O.BOARD[game], np.float32))       # first-call set-up of the torch kernels, outside the timing
    out = sys.stdout
    out.write("ready\n")
    out.flush()
    last = forest.counters()
    for line in sys.stdin:
        cmd = line.split()
        if not cmd:
            continue
        if cmd[0] == "quit":
            break
        if cmd[0] == "reset":
            forest.reset(roots)
            out.write("ok\n")
        elif cmd[0] == "step":
            t0 = time.perf_counter()
            forest.search(int(cmd[1]), O.EVAL_NET, cb)
            dt = time.perf_counter() - t0
            c = forest.counters()
            out.write("%d %d %d %.6f\n" % (c["simulations"] - last["simulations"], c["evaluations"] - last["evaluations"],
                                           c["terminal_leaves"] - last["terminal_leaves"], dt))
            last = c
        out.flush()


if __name__ == "__main__":
    main()
