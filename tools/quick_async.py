"""A/B of the search pipelines at the bench workload: asynchronous (default) vs lock-step (SPB_FLAG_LOCKSTEP).
usage: python tools/quick_async.py [games] [sims] [reps] [async-only]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selfplay_b200 as S
from selfplay_b200.synth import synthetic_roots_device
from selfplay_b200.weights_init import random_checkpoint

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
SIMS = int(sys.argv[2]) if len(sys.argv) > 2 else 800
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ONLY_ASYNC = len(sys.argv) > 4
blob = random_checkpoint(1, 0)
digests = {}
for name, flags in (("async", 0), ("lockstep", S.FLAG_LOCKSTEP)):
    if ONLY_ASYNC and name != "async":
        continue
    with S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, flags=flags, max_nodes_per_tree=8192) as e:
        e.load_weights(blob)
        roots = synthetic_roots_device(e, G)
        ms = []
        for r in range(REPS):
            e.reset_games(roots)
            e.reset_counters()
            e.search(SIMS)
            ms.append(e.last_search_timing()[0])
        a, c, i, n = e.root_children_all()
        ctr = e.counters()
        digests[name] = hashlib.sha256(a.tobytes() + c.tobytes() + i.tobytes() + n.tobytes()).hexdigest()
        best = min(ms)
        fl = 26630268.0 * ctr["evaluations"]
        print("%-9s ms/search %s  best %.2f ms = %.2f M sims/s, %.0f TFLOP/s over the search, evals %d, launches %d, digest %s" % (
            name, ["%.2f" % m for m in ms], best, G * SIMS / best / 1e3, fl / (best * 1e-3) / 1e12, ctr["evaluations"],
            ctr["kernel_launches"], digests[name][:16]), flush=True)
        if name == "async":
            st = e.async_stats()
            nb, nc, nv = max(1, st["batches"]), max(1, st["eval_ctas"]), max(1, st["tree_visits"])
            print("    boards/batch %.2f, claim wait/CTA %.1f ms (of which ticket wait %.1f ms), claims that found the ring empty %.1f%%, "
                  "mean avail at claim %.1f" % (st["boards"] / nb, st["claim_wait_ns"] / nc / 1e6, st["ticket_wait_ns"] / nc / 1e6,
                                                100.0 * st["claims_found_empty"] / nb, st["avail_sum"] / nb))
            print("    tree warps %d busy %.1f%%, %.2f us per tree visit, ready backlog per pop %.2f, pops that found nothing waiting %.1f%%" % (
                st["tree_warps"], 100.0 * st["tree_busy_ns"] / max(1, st["tree_warps"]) / (best * 1e6), st["tree_busy_ns"] / nv / 1e3,
                st["ready_backlog_sum"] / nv, 100.0 * st["ready_pops_starved"] / nv))
            print("   ", st)
if len(digests) == 2:
    print("identical results:", digests["async"] == digests["lockstep"])
