#!/usr/bin/env python
"""bench.py — MCTS simulations/sec for Connect4 at 800 sims/move (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W          # this framework, one rank per GPU
    python bench.py --impl reference ...                   # the reference's algorithm on the host CPU cores

One "step" = one `search` call (ref: Mcts::search, src/mcts.rs:196) of 800 simulations over 4,096 concurrent
Connect4 games per GPU, every tree fresh at a seeded synthetic root, evaluator = the
reference's 4x64 conv ResNet with random-init weights (bf16 on tcgen05, fp32 accumulate).  Work per GPU is
fixed (weak scaling); `value` is the whole-job aggregate.

  value     sims/s with the roots already resident in HBM; device time (CUDA events on the engine's stream),
            max over ranks.
  e2e       the same metric through the C ABI with HOST buffers: spb_reset_games (H2D) + spb_search +
            spb_root_children_all (D2H) per step, wall clock between synchronisations, max over ranks.
  roofline  the dominant kernel (the fused tcgen05 evaluator): FLOPs per launch / average launch duration
            measured live with CUDA events, against the measured bf16 peak of MEASURED_PEAKS.json.
  cpu_baseline  the C++ restatement of the reference (oracle/) with the torch CPU fp32 net, all host cores,
            on a bounded sample of the same workload (rank 0, N = 1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GAMES_PER_GPU = 4096
SIMS = 800
REF_GAMES_PER_THREAD = 100        # learner_concurrent.rs:56 num_batched_self_play_games
METRIC = "MCTS simulations/sec (whole box) Connect4 @800 sims/move"
WORKLOAD = "connect4_6x7_batched_selfplay_4096_games_x_800_sims_per_gpu"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.proc, self.lines = device_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (C++ restatement of mcts.rs + connect_four.rs) with the torch CPU net
# -------------------------------------------------------------------------------------------------
def cpu_reference_run(num_searches: int, threads: int | None = None, repeats: int = 1, warmup: int = 0, blob: bytes | None = None):
    """T worker threads, each an independent `search` over its own 100 games (the shape of main.rs:169's
    SelfPlayWorkers), evaluator = torch CPU fp32 forward of the same weights, 1 intra-op thread per worker.
    Returns list of (sims, seconds) per repeat."""
    import numpy as np
    import torch

    from oracle import pyoracle as O
    from oracle import torch_net
    from selfplay_b200.weights_init import random_checkpoint
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import synthetic_roots

    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(1)
    net = torch_net.load_tch_safetensors(blob or random_checkpoint(1, 0), 1)
    roots = synthetic_roots(O.GAME_C4, threads * REF_GAMES_PER_THREAD)

    def fn(enc):
        p, v, _ = torch_net.forward_probs(net, np.array(enc, copy=True))
        return p, v

    cb = O.make_eval_callback(O.GAME_C4, fn)
    out = []
    for r in range(warmup + repeats):
        sims, sec = O.baseline_run(O.GAME_C4, roots, threads, REF_GAMES_PER_THREAD, num_searches, evaluator=O.EVAL_NET, callback=cb)
        if r >= warmup:
            out.append((sims, sec))
    return out, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0                                          # rank 0 alone runs the CPU arm
    sims_per_step = args.ref_sims
    runs, threads = cpu_reference_run(sims_per_step, repeats=args.steps, warmup=args.warmup)
    total_sims = sum(s for s, _ in runs)
    total_sec = sum(t for _, t in runs)
    value = total_sims / total_sec
    sample = "%d threads x %d games x %d sims per step (same seeded roots, torch CPU fp32 evaluator, 1 intra-op thread per worker)" % (
        threads, REF_GAMES_PER_THREAD, sims_per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sims/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_sec / max(1, len(runs)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": "oracle port of src/mcts.rs + src/game/connect_four.rs (the Rust crate cannot be built here: no rustc/cargo), evaluator = torch CPU fp32",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
# this framework
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import selfplay_b200 as S
    from selfplay_b200.synth import synthetic_roots_device
    from selfplay_b200.weights_init import random_checkpoint

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")

    G, sims = args.games, args.sims
    blob = random_checkpoint(1, 0)
    # node pools sized for the full-game leg (a tree can carry nearly all of its nodes through a re-root and adds at most
    # 7 per simulation), so that no pool growth (spb_search widens the pools on demand) lands inside a timed region
    eng = S.Engine(game=S.GAME_C4, num_games=G, evaluator=S.EVAL_NET, device=local_rank,
                   game_id_base=rank * G, game_id_stride=world * G,
                   max_nodes_per_tree=1024 * ((7 * sims * (args.full_game_moves + 1) + 1024) // 1024))
    eng.load_weights(blob)
    roots = synthetic_roots_device(eng, G, start=rank * G)       # games are sharded by rank: rank r owns ids [r*G, (r+1)*G)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        eng.reset_games(roots)
        eng.search(sims)
        eng.root_children_all()

    # ---- timed region 1: device-resident (`value`) ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.reset_counters()
    barrier()
    dev_ms, t0 = 0.0, time.perf_counter()
    for _ in range(args.steps):
        eng.reset_games(roots)            # untimed part of `value`: fresh trees; roots then live in HBM
        eng.search(sims)                  # synchronous; its device time is measured with CUDA events inside
        dev_ms += eng.last_search_timing()[0]
    barrier()
    wall_s = time.perf_counter() - t0
    ctr = eng.counters()
    launches = ctr["kernel_launches"]
    dev_ms = rank_max(dev_ms)
    wall_s = rank_max(wall_s)

    # ---- timed region 2: end to end through the C ABI with host buffers (`e2e`) ----------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.reset_games(roots)            # H2D of the step's inputs
        eng.search(sims)
        acts, counts, ids, ncs = eng.root_children_all()   # D2H of the step's result
    barrier()
    e2e_s = rank_max(time.perf_counter() - t0)
    clocks = sampler.stop()
    h2d = G * 16
    d2h = G * (2 * 4 * S.MAX_ACTIONS + 4 + S.MAX_ACTIONS)
    assert int(counts[0].sum()) == sims - 1

    # ---- roofline of the dominant kernel (evaluator), measured live -----------------------------------
    eval_ms, n_pos, flops_pos = eng.time_evaluator(iters=30)
    peak_tf, peak_hbm, peak_src = _peaks()
    achieved_tf = flops_pos * n_pos / (eval_ms * 1e-3) / 1e12
    total_sims = world * G * sims * args.steps
    value = total_sims / (dev_ms * 1e-3)
    D = ctr["path_length_sum"] / max(1, ctr["simulations"])
    bbar = ctr["children_created"] / max(1, ctr["evaluations"])
    tree_bytes_per_sim = 16 * bbar * D + 36 * bbar + 16 * D + 96
    evals_per_step = ctr["evaluations"] / args.steps / sims
    step_us = dev_ms * 1e3 / args.steps / sims
    tree_us = max(1e-9, step_us - eval_ms * 1e3 * (evals_per_step / max(1, n_pos)))

    line = {
        "metric": METRIC, "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if (G, sims) == (GAMES_PER_GPU, SIMS) else "connect4_%d_games_x_%d_sims_per_gpu" % (G, sims),
                   "games_per_gpu": G, "sims_per_move": sims, "evaluator": "connect4 4x64 conv ResNet, random init (numpy seed 0), BN folded",
                   "parallelism": "games sharded by rank, no collective on the search path",
                   "cache": "inputs larger than L2: per-GPU node pools touched per step ~%d MB" % (ctr["nodes_live"] * 20 // (1 << 20)),
                   "timing": "CUDA events on the engine stream around each search, max over ranks; wall clock %.3f s" % wall_s},
        "e2e": {"value": total_sims / e2e_s, "unit": "sims/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "traffic": 849152, "kernel": "umma_v2::k_eval_umma<Connect4>", "peak_source": peak_src,
                     "positions_per_launch": n_pos, "flops_per_position": flops_pos, "avg_launch_ms": eval_ms},
        "roofline_tree": {"bound": "hbm", "unit": "GB/s", "peak": peak_hbm, "bytes_per_sim": tree_bytes_per_sim, "mean_path_length": D,
                          "mean_branching": bbar, "tree_us_per_step": tree_us,
                          "achieved": tree_bytes_per_sim * G / (tree_us * 1e-6) / 1e9,
                          "frac": tree_bytes_per_sim * G / (tree_us * 1e-6) / 1e9 / peak_hbm},
        "counters": {k: ctr[k] for k in ("simulations", "evaluations", "terminal_leaves")},
    }
    # ---- full-game leg: greedy last-max moves, subtree reuse, finished slots refilled from the same roots ------
    if args.full_game_moves > 0:
        eng.reset_games(roots)
        eng.drain_trajectories()
        barrier()
        t0 = time.perf_counter()
        finished = 0
        for _ in range(args.full_game_moves):
            eng.search(sims)                                   # carry + 800 more simulations per move (mcts.rs:161-192)
            finished += eng.selfplay_step(S.MOVE_GREEDY_LAST_MAX, restart_roots=roots)
        barrier()
        fg_s = rank_max(time.perf_counter() - t0)
        pos, _ = eng.drain_trajectories()
        line["full_game"] = {"moves": args.full_game_moves, "value": world * G * sims * args.full_game_moves / fg_s, "unit": "sims/s",
                             "finished_games_rank0": int(finished), "trajectory_positions_rank0": int(len(pos)),
                             "note": "search + on-device move selection + use_subtree per move, wall clock"}
    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) -----------------------------------
    if world == 1 and not args.no_cpu_baseline:
        eng.close()
        runs, threads = cpu_reference_run(args.ref_sims, repeats=1, warmup=0, blob=blob)
        s, t = runs[0]
        line["cpu_baseline"] = {"value": s / t, "unit": "sims/s", "cores": threads, "kind": "port",
                                "sample": "%d threads x %d games x %d sims, same seeded roots and weights, torch CPU fp32 evaluator" % (threads, REF_GAMES_PER_THREAD, args.ref_sims)}
    else:
        line["cpu_baseline"] = None
    # ---- multi-GPU: the one exchange of the path — trajectories to the learner rank -------------------
    if world > 1:
        from selfplay_b200.distributed import gather_records, gather_trajectories
        eng.selfplay_step(S.MOVE_GREEDY_LAST_MAX)
        gather_records(np.zeros(0, S.POSITION_DTYPE), np.zeros(0, np.uint64), dst=0)   # NCCL sets up its gather connections on first use
        barrier()
        t0 = time.perf_counter()
        pos, gids = gather_trajectories(eng, dst=0)
        line["trajectory_gather"] = {"positions": int(len(pos)) if rank == 0 else None, "seconds": time.perf_counter() - t0}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--sims", type=int, default=SIMS)
    ap.add_argument("--ref-sims", type=int, default=100, help="simulations per step of the CPU arm's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-game-moves", type=int, default=4, help="extra leg: self-play moves with subtree reuse (0 = skip)")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # not under torchrun: launch one rank per GPU ourselves
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version there) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
