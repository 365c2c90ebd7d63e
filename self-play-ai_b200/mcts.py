"""Host-side mirror of the reference's search interface (ref: src/mcts.rs:8-44,196), same names and
argument meaning, so that learner code written against the reference reads the same here:

    mcts = Mcts(args, engine)                  # Mcts { args, model }            mcts.rs:41-44
    trees = [Tree(engine, slot) ...]           # Tree::with_root_state / default  mcts.rs:67-89
    results = mcts.search(trees)               # Vec<(Policy, Vec<(usize, f32)>)> mcts.rs:196
    tree.use_subtree(child_id)                 # mcts.rs:161

The trees live on the GPU; a `Tree` is a handle on one engine slot.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .engine import Engine, State


@dataclass
class Args:
    """ref: mcts.rs:8-18, defaults mcts.rs:46-59.  Only `c` (via the Tree) and `num_searches` are read on
    the search path; the rest is carried for the learners."""
    c: float = 2.0
    num_searches: int = 600
    temperature: float = 1.25
    num_learn_iters: int = 10
    num_self_play_iters: int = 500
    num_parallel_self_play_games: int = 100
    batch_size: int = 32
    num_epochs: int = 4


class Tree:
    """Handle on one engine slot (ref: `Tree<T>` mcts.rs:32-39)."""

    def __init__(self, engine: Engine, slot: int, root_state: State | None = None):
        self.engine, self.slot = engine, slot
        self.state_history: list = []       # mcts.rs:37
        self.policy_history: list = []      # mcts.rs:38
        engine.reset_games(None if root_state is None else [root_state], slots=[slot])

    @classmethod
    def with_root_state(cls, engine: Engine, slot: int, state: State) -> "Tree":   # mcts.rs:86
        return cls(engine, slot, state)

    def use_subtree(self, new_root_id: int) -> None:                                  # mcts.rs:161
        self.engine.advance([new_root_id], slots=[self.slot])

    def arena_len(self) -> int:
        return self.engine.arena_len(self.slot)

    def node_state(self, node_id: int) -> State:                                      # tree.arena[id].state
        return self.engine.get_state(self.slot, node_id)


class Mcts:
    """ref: `Mcts<T>` mcts.rs:41-44; `search` mcts.rs:196-332."""

    def __init__(self, args: Args, engine: Engine):
        self.args, self.engine = args, engine

    def search(self, trees):
        """Runs args.num_searches lock-step simulations for every live slot of the engine and returns, for each
        tree in `trees`, (policy, [(child_arena_id, visit_count as f32), ...]) exactly like mcts.rs:310-331."""
        self.engine.search(self.args.num_searches)
        acts, counts, ids, ncs = self.engine.root_children_all()
        out = []
        for t in trees:
            s, n = t.slot, int(ncs[t.slot])
            policy = np.zeros(self.engine.A, np.float32)
            for j in range(n):
                policy[acts[s, j]] = np.float32(counts[s, j])
            policy = policy / policy.sum(dtype=np.float32) if n else policy
            out.append((policy, [(int(ids[s, j]), float(counts[s, j])) for j in range(n)]))
        return out
