"""Per-position error map of the chess network against torch fp32 (debugging aid for k_conv changes).
usage: python tools/chess_net_errmap.py [positions]   — prints, per position, max |logit error| / scale and the worst channel"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.abspath("tests"))   # under tools/run_variant.sh __file__ is not this file
import selfplay_b200 as S
import test_chess_search as T
from oracle import torch_net

n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
net = torch_net.make_chess_net(seed=0)
games = T.random_games(8, seed=31, max_plies=90)
games = [g for g in games if g.status() == S.ONGOING]
games = (games * (n // max(1, len(games)) + 1))[:n]
st, hist = T.export_all(games)
enc = np.stack([g.encode() for g in games])
_, v_ref, logit_ref = torch_net.chess_forward(net, enc)
scale = float(np.abs(logit_ref).max())
with S.ChessEngine(num_games=max(64, n), evaluator=S.EVAL_NET) as e:
    e.load_weights(torch_net.chess_to_safetensors_tch(net))
    for rep in range(2):
        _, v, lg = e.predict(st, hist, want_logits=True)
        err = np.abs(np.nan_to_num(lg, nan=1e30, posinf=1e30, neginf=-1e30) - logit_ref).reshape(n, 73, 64)
        bad = [(i, float(err[i].max() / scale), int(err[i].max(axis=1).argmax()), int((err[i] > 0.01 * scale).sum())) for i in range(n)]
        nb = [b for b in bad if b[1] > 0.01]
        print("rep %d: %d positions, %d bad; value err %.2e" % (rep, n, len(nb), float(np.abs(v - v_ref).max())))
        for b in nb[:40]:
            print("  position %d (tiles %d..%d): err/scale %.3g worst channel %d bad cells %d" % (b[0], b[0] * 9 // 16, (b[0] * 9 + 8) // 16, b[1], b[2], b[3]))
