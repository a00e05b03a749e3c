"""The CUDA-graph learner (deeprl_b200/learner.py, the path bench.py times) against the same update run eagerly, and the
async-replay prefetch branch (the graph form of ReplayWrapper(async_=True), reference replay.py:214-262).

Tolerance: the captured graph replays exactly the kernels of the eager update; the only run-to-run freedom is the order
of fp32 atomic adds (bias gradients, narrow-head weight gradients), so losses agree to 2e-3 relative over a few updates."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import bench
    import deeprl_b200 as rl
    rl.select_device(0)
    rl.Config.COMPUTE_DTYPE = torch.bfloat16
    bench.CAP = 30_000
    return bench, rl


def make(env, workload, prefetch):
    bench, rl = env
    return bench.build_learner(rl, workload, torch.device("cuda", 0), 0, 1, prefetch=prefetch)


def eager_losses(learner, n):
    out = []
    for _ in range(n):
        learner._main(), learner._opt()
        out.append(float(learner.loss))
    return out


def graph_losses(learner, n):
    out = []
    for _ in range(n):
        learner.update()
        out.append(float(learner.loss))
    return out


@pytest.mark.parametrize("workload", ["dqn", "per", "c51", "qr"])
def test_graph_replay_equals_eager_updates(env, workload):
    a, b = make(env, workload, False), make(env, workload, False)
    ref = eager_losses(a, 8)                               # 3 warm-up + 5 compared
    b.capture(warmup=3)
    got = graph_losses(b, 5)
    assert all(np.isfinite(ref)) and all(np.isfinite(got))
    np.testing.assert_allclose(got, ref[3:], rtol=2e-3)
    # the ring cursor advanced by feeds_per_update per update in both
    sa, sb = a.replay.ring_state.cpu().tolist(), b.replay.ring_state.cpu().tolist()
    assert sa[:3] == sb[:3]


@pytest.mark.parametrize("workload", ["dqn", "per"])
def test_prefetch_trains_on_the_previous_sample(env, workload):
    """async replay: update k uses the batch sampled during update k-1.  The very first update samples its own batch
    (nothing was prefetched), so it must equal the synchronous learner's first update; every later one must equal an eager
    replay of the same schedule."""
    sync, pre, pre2 = make(env, workload, False), make(env, workload, True), make(env, workload, True)
    l_sync = eager_losses(sync, 1)
    l_pre = eager_losses(pre, 8)
    np.testing.assert_allclose(l_pre[0], l_sync[0], rtol=2e-3)
    # 8 prefetch updates fed 9 x feeds rows (the first one also primed its own batch); the sync learner fed 1 x feeds
    assert pre.replay.ring_state.cpu().tolist()[0] == sync.replay.ring_state.cpu().tolist()[0] + 7 * pre.feeds + pre.feeds
    pre2.capture(warmup=3)
    got = graph_losses(pre2, 5)
    np.testing.assert_allclose(got, l_pre[3:], rtol=2e-3)
    # the batch buffers alternate: the two parities are different tensors
    assert pre2._batch[0].state.data_ptr() != pre2._batch[1].state.data_ptr()


def test_update_from_host_feeds_the_ring(env):
    bench, rl = env
    lr = make(env, "dqn", True)
    lr.capture(warmup=2, with_h2d=True)
    pos0 = int(lr.replay.ring_state[0])
    rng = np.random.RandomState(0)
    frames = rng.randint(0, 256, (4, 84 * 84)).astype(np.uint8)
    loss = lr.update_from_host(frames, np.arange(4, dtype=np.int32), np.ones(4), np.ones(4, dtype=np.int32))
    assert np.isfinite(loss)
    cap = lr.replay.memory_size if hasattr(lr.replay, "memory_size") else bench.CAP
    rows = [(pos0 + i) % cap for i in range(4)]
    got = lr.replay.frames[rows].cpu().numpy().reshape(4, -1)
    assert np.array_equal(got, frames)
    assert lr.replay.action[rows].cpu().tolist() == [0, 1, 2, 3]
