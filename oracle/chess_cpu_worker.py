"""One worker PROCESS of the CPU arm of `bench.py --game chess` (test infrastructure: see oracle/oracle.h).

Same shape as oracle/cpu_worker.py (one process per worker, like the reference's share-nothing SelfPlayWorker threads,
src/main.rs:169): the oracle port of src/mcts.rs over src/game/chess.rs (oracle/chess_oracle.cc) with the torch CPU fp32
restatement of src/model/chess.rs, 1 intra-op thread.

Protocol (stdin/stdout): "step <num_searches>" -> "<simulations> <evaluations> <terminal leaves> <seconds>"; "reset" -> "ok";
"quit".   usage: python chess_cpu_worker.py <worker index> <games per worker> <checkpoint file> <max_ply>
"""
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import pychess as P  # noqa: E402
from oracle import torch_net  # noqa: E402

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def synthetic_chess_root(g, max_ply=41):
    """The oracle-side twin of selfplay_b200.synth.synthetic_chess_roots_device."""
    while True:
        r = splitmix64(0xC4E55000 + g)
        plies = r % max_ply
        game = P.Game()
        ok = True
        for _ in range(plies):
            if game.status() != 0:
                ok = False
                break
            lm = game.legal_moves()
            r = splitmix64(r)
            game.make_move(lm[r % len(lm)])
        if ok and game.status() == 0:
            return game
        g += 1 << 32


def main():
    widx, games, max_ply = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[4])
    blob = open(sys.argv[3], "rb").read()
    torch.set_num_threads(1)
    net = torch_net.load_chess_tch_safetensors(blob)
    roots = [synthetic_chess_root(widx * games + i, max_ply) for i in range(games)]

    def fn(enc):
        p, v, _ = torch_net.chess_forward(net, np.array(enc, copy=True))
        return p, v

    forest = P.Forest(games)

    def reset():
        for i, g in enumerate(roots):
            forest.reset(i, g)

    reset()
    fn(np.zeros((1, 19, 8, 8), np.float32))
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
            reset()
            out.write("ok\n")
        elif cmd[0] == "step":
            t0 = time.perf_counter()
            forest.search(int(cmd[1]), 0, fn)
            dt = time.perf_counter() - t0
            c = forest.counters()
            out.write("%d %d %d %.6f\n" % (c["simulations"] - last["simulations"], c["evaluations"] - last["evaluations"],
                                           c["terminal_leaves"] - last["terminal_leaves"], dt))
            last = c
        out.flush()


if __name__ == "__main__":
    main()
