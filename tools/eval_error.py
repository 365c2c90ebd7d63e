"""Error of the evaluator kernels against the torch fp32 network over several nets: fresh (random init, randomised BatchNorm)
and trained-like (BatchNorm gammas 0.5..2, shifted statistics, larger heads).  Reported per net: max |dlogit| / max |logit|
(the bound of tests/test_gpu_evaluator.py), the worst per-position relative L2 error of the logit vector, the largest change of
a softmax probability, max |dvalue|.  Also the chess 10x256 network.  Needs a B200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

import selfplay_b200 as S  # noqa: E402
from helpers import random_states  # noqa: E402
from oracle import pyoracle as O, torch_net  # noqa: E402


def report(tag, lg, v, logit_ref, v_ref):
    scale = float(np.abs(logit_ref).max())
    rel_l2 = np.linalg.norm(lg - logit_ref, axis=1) / np.maximum(np.linalg.norm(logit_ref, axis=1), 1e-6)
    sm = lambda x: np.exp(x - x.max(1, keepdims=True)) / np.exp(x - x.max(1, keepdims=True)).sum(1, keepdims=True)   # noqa: E731
    print("%s (logit scale %.3f): max|dlogit|/scale %.2e | worst rel L2 per position %.2e | max|dp| %.2e | max|dvalue| %.2e" % (
        tag, scale, np.abs(lg - logit_ref).max() / scale, rel_l2.max(), np.abs(sm(lg) - sm(logit_ref)).max(), np.abs(v - v_ref).max()), flush=True)


for game, tag in ((S.GAME_C4, "c4"), (S.GAME_TTT, "ttt")):
    for trained in (False, True):
        for seed in range(4):
            net = torch_net.make_net(game, seed=seed, trained_like=trained)
            states = random_states(game, 1500, seed=100 + seed, include_terminal=False)
            enc = np.stack([O.encode(game, s) for s in states])
            _, v_ref, logit_ref = torch_net.forward_probs(net, enc)
            for flags, name in ((0, "tcgen05"), (S.FLAG_EVAL_SIMT, "simt")):
                with S.Engine(game=game, num_games=4, evaluator=S.EVAL_NET, flags=flags) as e:
                    e.load_weights(torch_net.to_safetensors_tch(net))
                    _, v, lg = e.predict(states, want_logits=True)
                report("%s %s seed %d %s" % (tag, "trained-like" if trained else "fresh", seed, name), lg, v, logit_ref, v_ref)

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_chess import export_all, random_games  # noqa: E402

games = [g for g in random_games(6, seed=41, max_plies=100) if g.status() == 0][::5][:64]
st, hist = export_all(games)
enc = np.stack([g.encode() for g in games])
for seed in range(2):
    net = torch_net.make_chess_net(seed=seed)
    _, v_ref, logit_ref = torch_net.chess_forward(net, enc)
    with S.ChessEngine(num_games=64, evaluator=S.EVAL_NET) as e:
        e.load_weights(torch_net.chess_to_safetensors_tch(net))
        _, v, lg = e.predict(st, hist, want_logits=True)
    report("chess 10x256 seed %d tcgen05" % seed, lg, v, logit_ref, v_ref)
