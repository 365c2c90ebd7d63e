"""Second, independent restatement of the reference hot path in pure Python + numpy.float32.

TEST INFRASTRUCTURE.  Written directly from the Rust sources (not from oracle/oracle.cc) with
list-of-lists boards and dict nodes, so that an error of transcription in either restatement
shows up as a disagreement.  Slow: only for small cases.

Follows src/mcts.rs:91-192,214-331; src/game/connect_four.rs:127-279; src/game/tictactoe.rs:135-236.
"""
import numpy as np

F = np.float32
M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def ndarray_sum(xs):
    """ndarray 0.15 unrolled_fold order."""
    xs = [F(x) for x in xs]
    p = [F(0)] * 8
    i = 0
    while len(xs) - i >= 8:
        for k in range(8):
            p[k] = F(p[k] + xs[i + k])
        i += 8
    acc = F(0)
    acc = F(acc + F(p[0] + p[4]))
    acc = F(acc + F(p[1] + p[5]))
    acc = F(acc + F(p[2] + p[6]))
    acc = F(acc + F(p[3] + p[7]))
    for x in xs[i:]:
        acc = F(acc + x)
    return acc


class C4:
    A, ROWS, COLS = 7, 6, 7

    def __init__(self):
        self.board = [[None] * 7 for _ in range(6)]
        self.player = 0
        self.n = 0
        self.status = 0  # 0 ongoing 1 tied 2 won

    def clone(self):
        s = C4()
        s.board = [r[:] for r in self.board]
        s.player, s.n, s.status = self.player, self.n, self.status
        return s

    def winner(self, lr, lc):
        b = self.board
        row = b[lr]
        for i in range(0, 7 - 4 + 1):
            if row[i] is not None and row[i] == row[i + 1] == row[i + 2] == row[i + 3]:
                return row[i]
        for i in range(0, 6 - 4 + 1):
            if b[i][lc] is not None and b[i][lc] == b[i + 1][lc] == b[i + 2][lc] == b[i + 3][lc]:
                return b[i][lc]
        start = max(-4, -min(lc, lr))
        end = min(0, min(7 - (lc + 4), 6 - (lr + 4)))
        for i in range(start, end + 1):
            r, c = lr + i, lc + i
            if b[r][c] is not None and b[r][c] == b[r + 1][c + 1] == b[r + 2][c + 2] == b[r + 3][c + 3]:
                return b[r][c]
        return None

    def next_state(self, a):
        if self.status != 0:
            return None
        row = next((i for i in range(6) if self.board[i][a] is None), None)
        if row is None:
            return None
        s = self.clone()
        s.board[row][a] = self.player
        s.player ^= 1
        s.n += 1
        if s.winner(row, a) is not None:
            s.status = 2
        elif s.n == 42:
            s.status = 1
        return s

    def valid_actions(self):
        if self.status != 0:
            return []
        return [c for c in range(7) if self.board[5][c] is None]

    def stones(self):
        x = o = 0
        for r in range(6):
            for c in range(7):
                if self.board[r][c] == 0:
                    x |= 1 << (c * 7 + r)
                elif self.board[r][c] == 1:
                    o |= 1 << (c * 7 + r)
        return x, o

    def encoding(self):
        e = np.zeros((3, 6, 7), dtype=np.float32)
        for r in range(6):
            for c in range(7):
                p = self.board[r][c]
                e[2 if p is None else (0 if p == self.player else 1), r, c] = 1.0
        return e


class TTT:
    A, ROWS, COLS = 9, 3, 3

    def __init__(self):
        self.board = [[None] * 3 for _ in range(3)]
        self.player = 0
        self.n = 0
        self.status = 0

    def clone(self):
        s = TTT()
        s.board = [r[:] for r in self.board]
        s.player, s.n, s.status = self.player, self.n, self.status
        return s

    def next_state(self, a):
        if self.status != 0:
            return None
        r, c = divmod(a, 3)
        if self.board[r][c] is not None:
            return None
        s = self.clone()
        b = s.board
        b[r][c] = self.player
        s.player ^= 1
        s.n += 1
        row_win = b[r][0] == b[r][1] == b[r][2]
        col_win = b[0][c] == b[1][c] == b[2][c]
        d1 = r == c and b[0][0] == b[1][1] == b[2][2]
        d2 = ((r == 1 and c == 1) or abs(r - c) == 2) and b[0][2] == b[1][1] == b[2][0]
        if row_win or col_win or d1 or d2:
            s.status = 2
        elif s.n == 9:
            s.status = 1
        return s

    def valid_actions(self):
        if self.status != 0:
            return []
        return [r * 3 + c for r in range(3) for c in range(3) if self.board[r][c] is None]

    def stones(self):
        x = o = 0
        for r in range(3):
            for c in range(3):
                if self.board[r][c] == 0:
                    x |= 1 << (r * 3 + c)
                elif self.board[r][c] == 1:
                    o |= 1 << (r * 3 + c)
        return x, o

    def encoding(self):
        e = np.zeros((3, 3, 3), dtype=np.float32)
        for r in range(3):
            for c in range(3):
                p = self.board[r][c]
                e[2 if p is None else (0 if p == self.player else 1), r, c] = 1.0
        return e


def det_eval(state):
    st = state.stones()
    mine, opp = st[state.player], st[state.player ^ 1]
    h = splitmix64(mine ^ splitmix64(opp))
    probs = [F(1 + ((h >> (4 * a)) & 7)) / F(64) for a in range(state.A)]
    v = F(F(((h >> 40) & 0xFF)) - F(128)) / F(128)
    return probs, v


def uniform_eval(state):
    return [F(1.0)] * state.A, F(0.0)


def mask(state, probs):
    va = set(state.valid_actions())
    masked = [F(probs[a]) * (F(1) if a in va else F(0)) for a in range(state.A)]
    s = ndarray_sum(masked)
    return [F(m / s) for m in masked]


class Tree:
    def __init__(self, state, c=2.0):
        self.c = F(c)
        self.arena = [dict(state=state, parent=None, action=None, prior=None, children=[], N=0, W=F(0))]

    def ucb(self, pid, cid):
        p, ch = self.arena[pid], self.arena[cid]
        q = F(0) if ch["N"] == 0 else F(F(F(-ch["W"]) / F(ch["N"]) + F(1)) / F(2))
        u = F(self.c * ch["prior"])
        u = F(u * np.sqrt(F(p["N"])))
        u = F(u / F(F(1) + F(ch["N"])))
        return F(q + u)

    def select(self, pid):
        ch = self.arena[pid]["children"]
        best = ch[0]
        for cand in ch[1:]:
            if not (self.ucb(pid, best) > self.ucb(pid, cand)):
                best = cand
        return best

    def expand(self, pid, policy):
        st = self.arena[pid]["state"]
        for a in st.valid_actions():
            self.arena[pid]["children"].append(len(self.arena))
            self.arena.append(dict(state=st.next_state(a), parent=pid, action=a, prior=F(policy[a]),
                                   children=[], N=0, W=F(0)))

    def backprop(self, nid, v):
        sign = F(1)
        node = self.arena[nid]
        while True:
            node["N"] += 1
            node["W"] = F(node["W"] + F(sign * v))
            sign = F(-sign)
            if node["parent"] is None:
                break
            node = self.arena[node["parent"]]

    def use_subtree(self, rid):
        old = self.arena
        new = []
        root = dict(old[rid])
        root["parent"] = None
        queue = [root]
        while queue:
            node = queue.pop(0)
            nid = len(new)
            for cid in node["children"]:
                ch = dict(old[cid])
                ch["parent"] = nid
                queue.append(ch)
            node["children"] = []
            if node["parent"] is not None:
                new[node["parent"]]["children"].append(nid)
            new.append(node)
        self.arena = new

    def root_counts(self):
        return [self.arena[c]["N"] for c in self.arena[0]["children"]]


def search(trees, num_searches, evaluator):
    for _ in range(num_searches):
        todo = []
        for t in trees:
            nid = 0
            while t.arena[nid]["children"]:
                nid = t.select(nid)
            st = t.arena[nid]["state"]
            if st.status == 2:
                t.backprop(nid, F(-1))
            elif st.status == 1:
                t.backprop(nid, F(0))
            else:
                todo.append((t, nid))
        for t, nid in todo:
            st = t.arena[nid]["state"]
            probs, v = evaluator(st)
            t.expand(nid, mask(st, probs))
            t.backprop(nid, v)
