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


def synthetic_chess_roots_device(engine, n: int, start: int = 0, max_ply: int = 41):
    """Seeded synthetic chess roots, generated with the DEVICE rules (spb_chess_legal_moves / spb_chess_next_states): game g
    is a random legal playout of `splitmix64(0xC4E55000 + g) % max_ply` plies from the start position; per ply
    r = splitmix64(r) and the (r % n_legal)-th legal move (this repo's move order) is played; a playout that ends the game
    is re-drawn with g + 2^32.  -> (states[n] CHESS_STATE_DTYPE, history[n, 512] u64).  tests build the same roots with the
    oracle (tests/test_chess_search.py)."""
    from .chess import CHESS_STATE_DTYPE, MAX_HISTORY, start_position
    out = np.zeros(n, dtype=CHESS_STATE_DTYPE)
    out_hist = np.zeros((n, MAX_HISTORY), np.uint64)
    g = np.arange(start, start + n, dtype=np.uint64)
    todo = np.arange(n)
    first = start_position()[0]
    while len(todo):
        r = _splitmix64(np.uint64(0xC4E55000) + g[todo])
        plies = (r % np.uint64(max_ply)).astype(np.int64)
        states = np.repeat(np.array([first], dtype=CHESS_STATE_DTYPE), len(todo))
        hist = np.zeros((len(todo), MAX_HISTORY), np.uint64)
        alive = np.ones(len(todo), dtype=bool)
        for ply in range(int(plies.max()) + 1 if len(plies) else 0):
            moves, counts, _, status, _ = engine.legal_moves(states, hist)
            alive &= status == 0                                       # a position that ended the game is re-drawn
            act_idx = np.nonzero(alive & (plies > ply))[0]
            if not len(act_idx):
                break
            r[act_idx] = _splitmix64(r[act_idx])
            k = (r[act_idx] % counts[act_idx].astype(np.uint64)).astype(np.int64)
            chosen = moves[act_idx, k]
            nxt, h2, err = engine.next_states(states[act_idx], hist[act_idx], chosen)
            assert not err.any()
            states[act_idx] = nxt
            hist[act_idx] = h2
        # final status of the positions reached
        _, _, _, status, _ = engine.legal_moves(states, hist)
        ok = alive & (status == 0)
        out[todo[ok]] = states[ok]
        out_hist[todo[ok]] = hist[ok]
        g[todo[~ok]] += np.uint64(1 << 32)
        todo = todo[~ok]
    return out, out_hist
