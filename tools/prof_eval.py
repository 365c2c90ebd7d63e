import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from oracle import torch_net
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 40
net = torch_net.make_net(1, seed=0, randomize_bn=False)
with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=S.FLAG_NO_GRAPH) as e:
    e.load_weights(torch_net.to_safetensors_tch(net))
    roots = synthetic_roots_device(e, G)
    e.reset_games(roots)
    e.search(sims)
    print("search ms", e.last_search_timing()[0], "per step us", e.last_search_timing()[0] * 1e3 / sims)
