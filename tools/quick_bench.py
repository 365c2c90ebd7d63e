import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from oracle import torch_net

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
net = torch_net.make_net(1, seed=0, randomize_bn=False)
blob = torch_net.to_safetensors_tch(net)
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_DET) as e:
    roots = synthetic_roots_device(e, G)
    for rep in range(3):
        e.reset_games(roots); e.reset_counters()
        e.search(sims)
        ms = e.last_search_timing()[0]
        c = e.counters()
        print("DET fused: %.2f ms  %.1f M sims/s  mean path %.2f  launches %d" % (ms, G * sims / ms / 1e3, c["path_length_sum"] / c["simulations"], c["kernel_launches"]), flush=True)
for flags, tag in [(0, "umma graph"), (S.FLAG_NO_GRAPH, "umma nograph")]:
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags) as e:
        e.load_weights(blob)
        for rep in range(3):
            e.reset_games(roots); e.reset_counters()
            t0 = time.time()
            e.search(sims)
            wall = (time.time() - t0) * 1e3
            ms = e.last_search_timing()[0]
            c = e.counters()
            print("NET %s: %.2f ms (wall %.2f)  %.2f M sims/s  %.1f us/step  evals %d  mean path %.2f" % (tag, ms, wall, G * sims / ms / 1e3, ms * 1e3 / sims, c["evaluations"], c["path_length_sum"] / c["simulations"]), flush=True)
