"""Size-independent properties of the replay index algebra and the sum tree, checked on the CPU restatement (oracle/) with
hypothesis (the reference has no tests; beyond the golden vectors generated from the live reference, these
properties are what the GPU tests assert at BASELINE's full sizes -- tests/test_gpu_parity.py -- where the oracle cannot run).
"""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle.replay import UniformReplay
from oracle.sum_tree import SumTree


def _ring(cap, fed, hl, n):
    rp = UniformReplay(cap, 4, n_step=n, discount=0.9, history_length=hl)
    for t in range(fed):
        rp.feed(dict(state=[np.full((2,), t, dtype=np.int64)], action=[t], reward=[float(t % 3 - 1)], mask=[int(t % 7 != 0)]))
    return rp


@settings(max_examples=120, deadline=None)
@given(cap=st.integers(6, 40), extra=st.integers(0, 90), hl=st.integers(1, 4), n=st.integers(1, 3))
def test_valid_transitions_never_straddle_the_write_cursor(cap, extra, hl, n):
    """replay.py:105-140: a valid index yields hl + n consecutive frames in TIME order -- the window never crosses the
    ring seam at `pos` (frames there are the oldest next to the newest) -- and the n-step reward / mask recursion."""
    fed = min(cap, hl + n + 1) + extra
    rp = _ring(cap, fed, hl, n)
    assert rp.size() == min(fed, cap)
    n_valid = 0
    for i in range(rp.size()):
        tr = rp.construct_transition(i)
        assert (tr is not None) == rp.valid_index(i)
        if tr is None:
            continue
        n_valid += 1
        s = np.asarray(tr.state).reshape(hl, -1)[:, 0]
        s2 = np.asarray(tr.next_state).reshape(hl, -1)[:, 0]
        assert np.all(np.diff(s) == 1) and np.all(s2 == s + n)            # consecutive time stamps
        t0 = int(s[-1])                                                   # time stamp of the frame at index i
        assert int(tr.action) == t0
        r, m = 0.0, 1
        for k in range(n - 1, -1, -1):
            mk = int((t0 + k) % 7 != 0)
            r = float((t0 + k) % 3 - 1) + mk * 0.9 * r
            m = m and mk
        assert tr.reward == r and int(tr.mask) == int(m)
    # every slot is valid except hl-1 at the oldest end, n at the newest end (and the same again around the seam)
    assert n_valid >= rp.size() - 2 * (hl - 1 + n)


@settings(max_examples=80, deadline=None)
@given(cap=st.integers(2, 33), ops=st.lists(st.tuples(st.sampled_from("agu"), st.floats(0.015625, 64.0, width=32),
                                                      st.floats(0.0, 1.0)), min_size=1, max_size=120))
def test_sum_tree_invariants(cap, ops):
    """sum_tree.py: after any add / get / update sequence every internal node equals the sum of its children up to the
    rounding of the `+= change` propagation, `get(s)` lands on the leaf whose prefix-sum bracket contains s, an update of
    a leaf that is not pending is ignored and the first update after a get wins."""
    t = SumTree(cap)
    for kind, p, u in ops:
        if kind == "a":
            t.add(float(p))
        elif kind == "g" and t.total() > 0:
            s = u * t.total()
            idx, prio, data = t.get(s)
            assert cap - 1 <= idx < 2 * cap - 1 and data == idx - cap + 1 and prio == t.tree[idx]
            order = _leaf_order(cap)                                      # leaves left-to-right (two depths when cap is not 2^k)
            before = sum(t.tree[j] for j in order[:order.index(idx)])
            assert before - 1e-9 * t.total() <= s <= before + prio + 1e-9 * t.total()
        elif kind == "u":
            idx = cap - 1 + int(u * cap) % cap
            pending = idx in t.pending
            old = t.tree[idx]
            t.update(idx, float(p))
            assert t.tree[idx] == (float(p) if pending else old)
            t.update(idx, float(p) + 1.0)                                 # second update: no longer pending -> ignored
            assert t.tree[idx] == (float(p) if pending else old)
        leaves = t.tree[cap - 1:]
        for node in range(cap - 1):
            assert abs(t.tree[node] - (t.tree[2 * node + 1] + t.tree[2 * node + 2])) <= 1e-9 * max(1.0, leaves.sum())


def _leaf_order(cap):
    """Leaves of the array heap in left-to-right (in-order) sequence."""
    out, stack = [], [0]
    while stack:
        node = stack.pop()
        left = 2 * node + 1
        if left >= 2 * cap - 1:
            out.append(node)
        else:
            stack.append(left + 1)
            stack.append(left)
    return out
