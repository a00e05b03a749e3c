"""TEST INFRASTRUCTURE -- CPU restatement of the reference sum tree (deep_rl/utils/sum_tree.py:6-67).

Array heap of ``2*cap-1`` float64 nodes, leaves at ``[cap-1, 2cap-2]``.  Facts that the CUDA
kernel must reproduce bit-for-bit and that this file encodes:

* ``update`` writes the leaf and then adds the SAME float64 ``change`` to every ancestor, one
  ``+=`` per node, in call order (sum_tree.py:16-20,58-60) -- internal nodes are never
  recomputed from their children, so their bits depend on the order of the additions.
* ``update`` is a no-op unless the leaf is in ``pending_idx`` (set by ``add`` and by ``get``),
  and removes it (sum_tree.py:54-57): the first update of a leaf after a ``get`` wins.
* ``get`` descends with ``s <= tree[left]`` -> left, else ``s -= tree[left]`` and right
  (sum_tree.py:23-33) and marks the leaf pending (sum_tree.py:63-67).
The recursion of the reference is unrolled into loops; the arithmetic is identical.
"""
import numpy as np


class SumTree:
    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.tree = np.zeros(2 * self.capacity - 1, dtype=np.float64)   # sum_tree.py:10
        self.write = 0
        self.n_entries = 0
        self.pending = set()

    def total(self):                                                    # sum_tree.py:35-36
        return self.tree[0]

    def update(self, idx, p):                                           # sum_tree.py:54-60
        idx = int(idx)
        if idx not in self.pending:
            return
        self.pending.discard(idx)
        change = p - self.tree[idx]
        self.tree[idx] = p
        node = idx
        while True:                                                     # sum_tree.py:16-20
            node = (node - 1) // 2
            self.tree[node] += change
            if node == 0:
                break

    def add(self, p):                                                   # sum_tree.py:39-51
        idx = self.write + self.capacity - 1
        self.pending.add(idx)
        self.update(idx, p)
        self.write = (self.write + 1) % self.capacity
        self.n_entries = min(self.n_entries + 1, self.capacity)

    def get(self, s):                                                   # sum_tree.py:23-33,63-67
        idx, n = 0, len(self.tree)
        while True:
            left = 2 * idx + 1
            if left >= n:
                break
            if s <= self.tree[left]:
                idx = left
            else:
                s = s - self.tree[left]
                idx = left + 1
        self.pending.add(idx)
        return idx, self.tree[idx], idx - self.capacity + 1
