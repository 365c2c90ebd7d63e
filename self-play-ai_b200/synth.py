"""Seeded synthetic Connect4 / tic-tac-toe root positions, generated with the DEVICE game
rules (spb_game_valid_actions / spb_game_next_states), so the benchmark's inputs never touch the oracle.

Game g: a random legal playout of `splitmix64(0x5EED0000 + g) % max_ply` plies from the empty board; at every
ply the generator state advances r = splitmix64(r) and the (r % n_legal)-th legal action is played; if the
playout ends the game, g is re-drawn as g + 2^32.  tests/helpers.py builds the same positions with the oracle.
"""
import numpy as np

from .engine import STATE_DTYPE

M64 = np.uint64((1 << 64) - 1)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x.copy()
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_roots_device(engine, n: int, start: int = 0, max_ply: int = 21) -> np.ndarray:
    out = np.zeros(n, dtype=STATE_DTYPE)
    g = np.arange(start, start + n, dtype=np.uint64)
    todo = np.arange(n)
    while len(todo):
        r = _splitmix64(np.uint64(0x5EED0000) + g[todo])
        plies = (r % np.uint64(max_ply)).astype(np.int64)
        states = np.zeros(len(todo), dtype=STATE_DTYPE)
        alive = np.ones(len(todo), dtype=bool)
        for ply in range(int(plies.max()) if len(plies) else 0):
            act_idx = np.nonzero(alive & (plies > ply))[0]
            if not len(act_idx):
                break
            r[act_idx] = _splitmix64(r[act_idx])
            masks = engine.game_valid_actions(states[act_idx])
            n_legal = np.array([bin(int(m)).count("1") for m in masks], dtype=np.uint64)
            k = (r[act_idx] % n_legal).astype(np.int64)
            actions = np.zeros(len(act_idx), np.uint8)
            for j, (m, kk) in enumerate(zip(masks, k)):
                bits = [a for a in range(9) if int(m) >> a & 1]
                actions[j] = bits[kk]
            nxt, err = engine.game_next_states(states[act_idx], actions)
            assert not err.any()
            states[act_idx] = nxt
            ended = nxt["status"] != 0
            alive[act_idx[ended]] = False
        ok = alive
        out[todo[ok]] = states[ok]
        g[todo[~ok]] += np.uint64(1 << 32)
        todo = todo[~ok]
    return out
