"""Shared test helpers: seeded synthetic positions built with the oracle's rules."""
import numpy as np

from oracle import pyoracle as O

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def synthetic_root(game, g, max_ply=21):
    """Game g: seeded random legal playout of splitmix64(0x5EED0000+g) % max_ply plies from the empty board;
    move = (rng %  n_legal)-th legal action with rng advanced by splitmix64 per ply; re-drawn with g += 2^32
    when the playout ends the game."""
    while True:
        r = splitmix64(0x5EED0000 + g)
        plies = r % max_ply
        s = O.State()
        ok = True
        for _ in range(plies):
            va = O.valid_actions(game, s)
            r = splitmix64(r)
            s = O.next_state(game, s, va[r % len(va)])
            if s.status != O.ONGOING:
                ok = False
                break
        if ok:
            return s
        g += 1 << 32


def synthetic_roots(game, n, start=0, max_ply=21):
    return [synthetic_root(game, start + i, max_ply) for i in range(n)]


def random_states(game, n, seed, include_terminal=True):
    """Random reachable positions (uniform random playouts of random length)."""
    rng = np.random.default_rng(seed)
    out = []
    max_len = 42 if game == O.GAME_C4 else 9
    while len(out) < n:
        s = O.State()
        L = int(rng.integers(0, max_len + 1))
        for _ in range(L):
            va = O.valid_actions(game, s)
            if not va:
                break
            s = O.next_state(game, s, int(rng.choice(va)))
        if include_terminal or s.status == O.ONGOING:
            out.append(s)
    return out
