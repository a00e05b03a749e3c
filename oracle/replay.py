"""TEST INFRASTRUCTURE -- CPU restatement of the reference replay path
(deep_rl/component/replay.py:57-196), with the random draws made injectable so that the CUDA
kernels and this oracle can consume the SAME stream.

Storage is kept the way the reference keeps it (python lists of per-step items, one
``np.array`` stack per sampled transition) so that timing this port is representative of the
reference's own CPU cost.
"""
import random
from collections import namedtuple

import numpy as np

from .sum_tree import SumTree

Transition = namedtuple("Transition", ["state", "action", "reward", "next_state", "mask"])          # replay.py:15
PrioritizedTransition = namedtuple(                                                                    # replay.py:16-17
    "PrioritizedTransition", ["state", "action", "reward", "next_state", "mask", "sampling_prob", "idx"])

KEYS = ("state", "action", "reward", "mask")


class UniformReplay:
    def __init__(self, memory_size, batch_size, n_step=1, discount=1, history_length=1):
        self.memory_size, self.batch_size = int(memory_size), int(batch_size)
        self.n_step, self.discount, self.history_length = int(n_step), discount, int(history_length)
        self.data = {k: [] for k in KEYS}
        self.pos = 0
        self._size = 0

    def size(self):
        return self._size

    def full(self):
        return self._size == self.memory_size

    def feed(self, data):
        """replay.py:75-90.  NOTE the reference quirk kept on purpose: in the overwrite branch the
        item goes to ``storage[self.pos]`` (the position at the START of the call), not to the
        running ``pos`` (replay.py:87)."""
        pos, size = self.pos, self._size
        for k, vs in data.items():
            if k not in self.data:
                raise RuntimeError("Undefined key")
            store = self.data[k]
            pos, size = self.pos, self._size
            for v in vs:
                if pos >= len(store):
                    store.append(v)
                    size += 1
                else:
                    store[self.pos] = v
                pos = (pos + 1) % self.memory_size
        self.pos, self._size = pos, size

    def valid_index(self, i):                                           # replay.py:105-110
        hl, n = self.history_length, self.n_step
        if i - hl + 1 >= 0 and i + n < self.pos:
            return True
        if i - hl + 1 >= self.pos and i + n < self._size:
            return True
        return False

    def construct_transition(self, i):                                  # replay.py:112-140
        if not self.valid_index(i):
            return None
        hl, n = self.history_length, self.n_step
        first = i - hl + 1
        frames = self.data["state"]
        state = [frames[t] for t in range(first, i + 1)]
        nxt = [frames[t] for t in range(first + n, i + n + 1)]
        if hl == 1:
            state, nxt = state[0], nxt[0]
        state, nxt = np.array(state), np.array(nxt)
        rew = self.data["reward"][i:i + n]
        msk = self.data["mask"][i:i + n]
        cum_r, cum_m = 0, 1
        for t in range(n - 1, -1, -1):                                  # replay.py:137-139
            cum_r = rew[t] + msk[t] * self.discount * cum_r
            cum_m = cum_m and msk[t]
        return Transition(state, self.data["action"][i], cum_r, nxt, cum_m)

    @staticmethod
    def _stack(rows, cls):
        return cls(*[np.asarray(col) for col in zip(*rows)])          # replay.py:101-103

    def sample(self, batch_size=None, candidates=None):
        """replay.py:92-103.  ``candidates`` (iterable of ints) replaces ``np.random.randint(0,size)``;
        returns (Transition, accepted_indices, n_candidates_consumed)."""
        bs = self.batch_size if batch_size is None else batch_size
        rows, taken, used = [], [], 0
        it = iter(candidates) if candidates is not None else None
        while len(rows) < bs:
            i = int(next(it)) if it is not None else np.random.randint(0, self._size)
            used += 1
            tr = self.construct_transition(i)
            if tr is not None:
                rows.append(tr)
                taken.append(i)
        return self._stack(rows, Transition), np.asarray(taken, dtype=np.int64), used


class PrioritizedReplay(UniformReplay):
    def __init__(self, memory_size, batch_size, n_step=1, discount=1, history_length=1):
        super().__init__(memory_size, batch_size, n_step, discount, history_length)
        self.tree = SumTree(memory_size)
        self.max_priority = 1

    def feed(self, data):                                               # replay.py:160-162: ONE leaf per call
        super().feed(data)
        self.tree.add(self.max_priority)

    def sample(self, batch_size=None, uniforms=None, fills=None):
        """replay.py:164-191.  ``uniforms[i]`` in [0,1) replaces the draw inside
        ``random.uniform(a, b)`` (CPython computes ``a + (b-a)*random()``); ``fills`` is the stream
        of list positions that replaces ``random.choice`` in the back-fill loop."""
        bs = self.batch_size if batch_size is None else batch_size
        total = self.tree.total()
        segment = total / bs
        rows = []
        for i in range(bs):
            a = segment * i
            b = segment * (i + 1)
            u = random.random() if uniforms is None else float(uniforms[i])
            s = a + (b - a) * u
            idx, p, data_index = self.tree.get(s)
            tr = self.construct_transition(data_index)
            if tr is None:
                continue
            rows.append(PrioritizedTransition(*tr, sampling_prob=p / self.tree.total(), idx=idx))
        k = 0
        while len(rows) < bs:                                           # replay.py:184-186
            if fills is None:
                rows.append(random.choice(rows))
            else:
                rows.append(rows[int(fills[k]) % len(rows)])
                k += 1
        return self._stack(rows, PrioritizedTransition)

    def update_priorities(self, info):                                  # replay.py:193-196
        for idx, priority in info:
            self.max_priority = max(self.max_priority, priority)
            self.tree.update(idx, priority)
