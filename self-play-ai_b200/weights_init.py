"""Random-init checkpoints of the reference's policy/value nets, written as the safetensors file tch's
VarStore::save would produce (ref: src/learner.rs:192; architecture src/model/connect_four.rs:50-73,
src/model/tictactoe.rs:50-73, src/model/mod.rs:152-184).  numpy only — used by bench.py to make synthetic
weights ("random-init weights of that architecture"); the reference arm loads the same bytes into torch.

Initialisation follows torch/tch defaults closely enough for a benchmark: conv / linear weights and biases
~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)), BatchNorm gamma 1, beta 0, running_mean 0, running_var 1.
"""
import json
import struct

import numpy as np

GEOMETRY = {0: (3, 3, 9), 1: (6, 7, 7)}   # game -> rows, cols, actions


def _uniform(rng, shape, fan_in):
    b = 1.0 / np.sqrt(fan_in)
    return rng.uniform(-b, b, size=shape).astype(np.float32)


def random_checkpoint(game: int, seed: int = 0, blocks: int = 4, hidden: int = 64) -> bytes:
    rows, cols, actions = GEOMETRY[game]
    rng = np.random.default_rng(seed)
    created = []                                   # (base name, array) in tch creation order

    def conv(ic, oc):
        created.append(("bias", _uniform(rng, (oc,), ic * 9)))
        created.append(("weight", _uniform(rng, (oc, ic, 3, 3), ic * 9)))

    def bn(c):
        created.append(("weight", np.ones(c, np.float32)))
        created.append(("bias", np.zeros(c, np.float32)))
        created.append(("running_mean", np.zeros(c, np.float32)))
        created.append(("running_var", np.ones(c, np.float32)))

    def linear(i, o):
        created.append(("bias", _uniform(rng, (o,), i)))
        created.append(("weight", _uniform(rng, (o, i), i)))

    conv(3, hidden); bn(hidden)
    for _ in range(blocks):
        conv(hidden, hidden); bn(hidden)
        conv(hidden, hidden); bn(hidden)
    conv(hidden, 32); bn(32); linear(32 * rows * cols, actions)
    conv(hidden, 3); bn(3); linear(3 * rows * cols, 1)

    header, chunks, off, seen, count = {}, [], 0, set(), 0
    for base, arr in created:
        name = base if base not in seen else "%s__%d" % (base, count)     # tch de-duplicates with the variable count
        seen.add(base)
        count += 1
        b = np.ascontiguousarray(arr, dtype=np.float32).tobytes()
        header[name] = {"dtype": "F32", "shape": list(arr.shape), "data_offsets": [off, off + len(b)]}
        off += len(b)
        chunks.append(b)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    return struct.pack("<Q", len(hj)) + hj + b"".join(chunks)


def random_chess_checkpoint(seed: int = 0, blocks: int = 10, hidden: int = 256) -> bytes:
    """Random-init checkpoint of the chess net (ref: src/model/chess.rs:50-73) with tch's creation-order names:
    torso (conv, bn) x 21, policy head conv1x1 256->256, conv1x1 256->73, value head conv1x1 256->1, Linear 64->256,
    Linear 256->1."""
    rng = np.random.default_rng(seed)
    created = []

    def conv(ic, oc, k):
        created.append(("bias", _uniform(rng, (oc,), ic * k * k)))
        created.append(("weight", _uniform(rng, (oc, ic, k, k), ic * k * k)))

    def bn(c):
        created.append(("weight", np.ones(c, np.float32)))
        created.append(("bias", np.zeros(c, np.float32)))
        created.append(("running_mean", np.zeros(c, np.float32)))
        created.append(("running_var", np.ones(c, np.float32)))

    def linear(i, o):
        created.append(("bias", _uniform(rng, (o,), i)))
        created.append(("weight", _uniform(rng, (o, i), i)))

    conv(19, hidden, 3); bn(hidden)
    for _ in range(blocks):
        conv(hidden, hidden, 3); bn(hidden)
        conv(hidden, hidden, 3); bn(hidden)
    conv(hidden, 256, 1); conv(256, 73, 1)
    conv(hidden, 1, 1); linear(64, 256); linear(256, 1)

    header, chunks, off, seen, count = {}, [], 0, set(), 0
    for base, arr in created:
        name = base if base not in seen else "%s__%d" % (base, count)
        seen.add(base)
        count += 1
        b = np.ascontiguousarray(arr, dtype=np.float32).tobytes()
        header[name] = {"dtype": "F32", "shape": list(arr.shape), "data_offsets": [off, off + len(b)]}
        off += len(b)
        chunks.append(b)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    return struct.pack("<Q", len(hj)) + hj + b"".join(chunks)
