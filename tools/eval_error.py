"""Error of the evaluator kernels against the torch fp32 network over several random nets (tolerance in tests: 1e-2 of scale)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import selfplay_b200 as S
from oracle import pyoracle as O, torch_net
from helpers import random_states
for game, tag in ((S.GAME_C4, "c4"), (S.GAME_TTT, "ttt")):
    for seed in range(6):
        net = torch_net.make_net(game, seed=seed)
        states = random_states(game, 1500, seed=100 + seed, include_terminal=False)
        enc = np.stack([O.encode(game, s) for s in states])
        probs_ref, v_ref, logit_ref = torch_net.forward_probs(net, enc)
        scale = float(np.abs(logit_ref).max())
        row = []
        for flags, name in ((0, "pair"), (S.FLAG_EVAL_V1, "v1"), (S.FLAG_EVAL_SIMT, "simt")):
            with S.Engine(game=game, num_games=4, evaluator=S.EVAL_NET, flags=flags) as e:
                e.load_weights(torch_net.to_safetensors_tch(net))
                pol, v, lg = e.predict(states, want_logits=True)
            row.append("%s logit %.2e value %.2e" % (name, np.abs(lg - logit_ref).max() / scale, np.abs(v - v_ref).max()))
        print("%s seed %d (logit scale %.3f): " % (tag, seed, scale) + " | ".join(row), flush=True)
