"""CPU tests of the multi-GPU plumbing (gloo, world_size 2): game sharding and the trajectory gather
produce a result byte-identical to a 1-rank run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from selfplay_b200.distributed import merge_trajectories, shard_range  # noqa: E402
from selfplay_b200.engine import POSITION_DTYPE  # noqa: E402


def fake_game(game_id: int):
    """Deterministic fake trajectory of a game: 3 + game_id % 5 positions."""
    n = 3 + game_id % 5
    pos = np.zeros(n, dtype=POSITION_DTYPE)
    for ply in range(n):
        pos[ply]["stones"] = (game_id * 1000 + ply, game_id * 7 + ply)
        pos[ply]["visit_counts"] = [(game_id + ply + a) % 97 for a in range(9)]
        pos[ply]["current_player"] = ply & 1
        pos[ply]["ply"] = ply
        pos[ply]["outcome"] = (game_id % 3) - 1
    return pos


def rank_records(lo, hi, shuffle_seed):
    """What one rank drains: its games in a rank-local (nondeterministic on a GPU) finishing order."""
    games = list(range(lo, hi))
    np.random.default_rng(shuffle_seed).shuffle(games)
    parts = [fake_game(g) for g in games]
    pos = np.concatenate(parts) if parts else np.zeros(0, POSITION_DTYPE)
    ids = np.concatenate([np.full(len(p), g, np.uint64) for p, g in zip(parts, games)]) if parts else np.zeros(0, np.uint64)
    # the engine already orders by (game id, ply) inside one rank; keep the shuffle to prove the merge sorts globally
    return pos, ids


def test_shard_range_partitions_all_games():
    for total in (1, 7, 100, 4096, 32768):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                seen += list(range(lo, hi))
                assert hi - lo in (total // world, total // world + 1)
            assert seen == list(range(total))


def test_merge_orders_by_game_then_ply():
    a = rank_records(0, 5, 1)
    b = rank_records(5, 9, 2)
    pos, ids = merge_trajectories([b, a])
    want_pos, want_ids = merge_trajectories([rank_records(0, 9, 3)])
    assert pos.tobytes() == want_pos.tobytes() and ids.tobytes() == want_ids.tobytes()
    assert list(ids) == sorted(ids) and all(pos["ply"][i] <= pos["ply"][i + 1] or ids[i] != ids[i + 1] for i in range(len(ids) - 1))


def _worker(rank, world, port, total_games, out_path):
    import torch.distributed as dist
    from selfplay_b200.distributed import gather_records, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total_games, rank, world)
    pos, ids = rank_records(lo, hi, 100 + rank)
    mpos, mids = gather_records(pos, ids, dst=0)
    if rank == 0:
        np.savez(out_path, pos=mpos.view(np.uint8), ids=mids)
    else:
        assert len(mpos) == 0
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_games", [11, 64])
def test_two_rank_gather_equals_single_rank(tmp_path, total_games):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, total_games, out), nprocs=2, join=True)
    got = np.load(out)
    want_pos, want_ids = merge_trajectories([rank_records(0, total_games, 7)])
    assert got["pos"].tobytes() == want_pos.tobytes()
    assert got["ids"].tobytes() == want_ids.tobytes()
