"""HBM-resident sum tree with the reference's interface (``deep_rl/utils/sum_tree.py:6-67``).

``tree`` is a float64 device tensor of ``2*capacity-1`` heap nodes; ``pending`` (uint8 per data slot)
replaces the reference's ``pending_idx`` set.  All arithmetic runs in ``csrc/sumtree.cu`` and is
bit-identical to the reference (sequential float64 additions per node, in call order).  The scalar
methods below (``add`` / ``update`` / ``get`` / ``total``) keep the reference's call signatures for
drop-in use; the batched entry points (``add_n`` / ``update_batch`` / ``sample_batch``) are what
``PrioritizedReplay`` uses.
"""
import numpy as np
import torch

from .. import _lib


class SumTree:
    def __init__(self, capacity, device=None, ring_state=None):
        from .config import Config
        self.device = _lib.require_cuda(device if device is not None else Config.DEVICE)
        self.capacity = int(capacity)
        if self.capacity < 2:
            raise ValueError("SumTree capacity must be >= 2 (the reference recurses forever at 1, sum_tree.py:16-20)")
        self.tree = torch.zeros(2 * self.capacity - 1, dtype=torch.float64, device=self.device)
        self.pending = torch.zeros(self.capacity, dtype=torch.uint8, device=self.device)
        # ring_state[3] = write cursor; shared with the replay ring when owned by a PrioritizedReplay
        self.ring_state = ring_state if ring_state is not None else torch.zeros(8, dtype=torch.int64, device=self.device)
        if ring_state is None:
            self.ring_state[2] = self.capacity
        self.n_entries = 0
        self._scratch = torch.empty(2 * 1024, dtype=torch.float64, device=self.device)
        self._one = torch.ones(1, dtype=torch.float64, device=self.device)
        self._status = torch.zeros(2, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ batched device entry points
    def add_n(self, n, priority):
        """``n`` x ``add(priority)``; ``priority`` is a float64 device scalar tensor (replay.py:162)."""
        _lib.call("b2rl_sumtree_add", _lib.ptr(self.tree), _lib.ptr(self.pending), self.capacity,
                  _lib.ptr(self.ring_state), _lib.ptr(priority), int(n), _lib.ptr(self._scratch), _lib.stream())
        self.n_entries = min(self.n_entries + int(n), self.capacity)

    def update_batch(self, tree_idx, priority, max_priority):
        """``update(idx_i, p_i)`` for the rows in order, and ``max_priority = max(.., p_i)`` (replay.py:193-196).
        ``tree_idx`` int64 [B], ``priority`` float32 [B], ``max_priority`` float64 [1], all on the device."""
        B = tree_idx.numel()
        _lib.call("b2rl_sumtree_update", _lib.ptr(self.tree), _lib.ptr(self.pending), self.capacity,
                  _lib.ptr(tree_idx), _lib.ptr(priority), B, _lib.ptr(max_priority), _lib.ptr(self._scratch),
                  _lib.stream())

    def sample_batch(self, B, history, n_step, tree_idx_out, data_idx_out, prob_out, status_out, uniforms=None,
                     fills=None, seed=0):
        _lib.call("b2rl_sumtree_sample", _lib.ptr(self.tree), _lib.ptr(self.pending), self.capacity,
                  _lib.ptr(self.ring_state), _lib.ptr(uniforms), _lib.ptr(fills), int(seed), int(history), int(n_step),
                  int(B), _lib.ptr(tree_idx_out), _lib.ptr(data_idx_out), _lib.ptr(prob_out), _lib.ptr(status_out),
                  _lib.stream())

    # ------------------------------------------------------------------ reference-shaped scalar API
    @property
    def write(self):
        return int(self.ring_state[3].item())

    def total(self):
        return float(self.tree[0].item())

    def add(self, p, data=None):
        self.add_n(1, torch.tensor([float(p)], dtype=torch.float64, device=self.device))

    def update(self, idx, p):
        ti = torch.tensor([int(idx)], dtype=torch.int64, device=self.device)
        pr = torch.tensor([np.float32(p)], dtype=torch.float32, device=self.device)
        if float(np.float32(p)) != float(p):
            raise ValueError("SumTree.update takes float32-representable priorities (to_np of an fp32 tensor)")
        scratch_max = torch.full((1,), float("inf"), dtype=torch.float64, device=self.device)
        self.update_batch(ti, pr, scratch_max)

    def get(self, s):
        """``get(s)`` (sum_tree.py:63-67) -> (tree_idx, priority, data_idx); one device descent."""
        pre = torch.tensor([float(s)], dtype=torch.float64, device=self.device)
        ti = torch.empty(1, dtype=torch.int64, device=self.device)
        pr = torch.empty(1, dtype=torch.float64, device=self.device)
        _lib.call("b2rl_sumtree_get", _lib.ptr(self.tree), _lib.ptr(self.pending), self.capacity, _lib.ptr(pre), 1,
                  _lib.ptr(ti), _lib.ptr(pr), _lib.stream())
        idx = int(ti.item())
        return idx, float(pr.item()), idx - self.capacity + 1
