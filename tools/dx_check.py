import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import selfplay_b200 as S
from selfplay_b200.weights_init import random_checkpoint
from helpers import random_states
from oracle import pyoracle as O
blob = random_checkpoint(1, 0)
for n in (1, 2, 5, 9, 12):
    states = random_states(O.GAME_C4, n, seed=3, include_terminal=False)
    outs = []
    for flags in (0, S.FLAG_EVAL_DX):
        with S.Engine(game=S.GAME_C4, num_games=4, evaluator=S.EVAL_NET, flags=flags) as e:
            e.load_weights(blob)
            pol, v, lg = e.predict(states, want_logits=True)
            outs.append((lg.copy(), v.copy()))
    d = np.abs(outs[0][0] - outs[1][0]).max(axis=1)
    print("n", n, "max|dlogit| per board", np.round(d, 4), "dv", np.round(np.abs(outs[0][1] - outs[1][1]), 4), flush=True)
    if n == 1: print(outs[0][0], outs[1][0])
