"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, tiny sizes.
usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py [tree|eval|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("tree", "all"):
    for game in (S.GAME_TTT, S.GAME_C4):
        for flags, k in ((0, 1), (S.FLAG_FORCE_SPLIT, 1), (S.FLAG_FORCE_SPLIT | S.FLAG_LOCKSTEP | S.FLAG_NO_GRAPH, 1), (0, 4)):
            with S.Engine(game=game, num_games=24, evaluator=S.EVAL_DET, flags=flags, leaves_per_tree=k, max_nodes_per_tree=64) as e:
                roots = synthetic_roots_device(e, 24, max_ply=8 if game == S.GAME_C4 else 3)
                e.reset_games(roots)
                for _ in range(4):                       # search, on-device ply (reroot / finish / restart), pool growth
                    e.search(40)
                    e.selfplay_step(S.MOVE_GREEDY_LAST_MAX, restart_roots=roots)
                e.search(16)
                e.selfplay_step(S.MOVE_TEMPERATURE, temperature=1.25, seed=7)
                a, c, i, n = e.root_children_all()
                live = [g for g in range(24) if n[g] > 0]
                e.advance([int(i[g][0]) for g in live], slots=live)
                e.drain_trajectories()
                e.get_state(0, 0), e.node_stats(0, 0), e.counters()
            print("tree ok", game, flags, k, flush=True)
if what in ("eval", "all"):
    for game in (S.GAME_TTT, S.GAME_C4):
        for flags in (0, S.FLAG_LOCKSTEP | S.FLAG_NO_GRAPH):
            with S.Engine(game=game, num_games=20, evaluator=S.EVAL_NET, flags=flags) as e:
                e.load_weights(random_checkpoint(1 if game == S.GAME_C4 else 0, 0))
                roots = synthetic_roots_device(e, 20, max_ply=8 if game == S.GAME_C4 else 3)
                e.predict(roots, want_logits=True)
                e.reset_games(roots)
                e.search(12)
                e.load_weights(random_checkpoint(1 if game == S.GAME_C4 else 0, 1))
                e.search(5)
            print("eval ok", game, flags, flush=True)
print("sanitize target done")
