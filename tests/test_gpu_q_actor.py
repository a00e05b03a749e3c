"""The captured actor forward of the DQN family (component/actor.py GraphedQActor; reference DQN_agent.py:24-45): action values
of the graph replay against (i) the same network called eagerly on the same bf16 space-to-depth input (identical kernels) and (ii) the oracle's fp32 restatement of ``network(ImageNormalizer(stack))`` (oracle/nets.py) within bf16
tolerance; then through ``DQNAgent.step()`` with ``config.cuda_graph``: the graph is re-captured when the learner takes over
the packed operands, and keeps tracking the fp32 master weights after optimizer steps."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.fixture(scope="module")
def rl():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import deeprl_b200 as rl
    rl.select_device(0)
    rl.Config.COMPUTE_DTYPE = torch.bfloat16
    return rl


def oracle_q(net, kind, stacks, atoms=None):
    from oracle import nets
    sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
    x = torch.from_numpy(np.asarray(stacks, dtype=np.float64) / 255.0).float()      # ImageNormalizer + tensor() (normalizer.py:64-66)
    phi = nets.nature_body(sd, x)
    if kind == "vanilla":
        return nets.vanilla_q(sd, phi)
    if kind == "dueling":
        return nets.dueling_q(sd, phi)
    if kind == "categorical":
        prob, _ = nets.categorical(sd, phi, 4, 51)
        return (prob * torch.from_numpy(atoms).float()).sum(-1)
    return nets.quantile(sd, phi, 4, 50).mean(-1)


@pytest.mark.parametrize("kind", ["vanilla", "dueling", "categorical", "quantile"])
@pytest.mark.parametrize("n_env", [1, 3])
def test_graphed_q_values(rl, kind, n_env):
    from deeprl_b200.component.actor import GraphedQActor
    from deeprl_b200.network.fused import frame_scale
    torch.manual_seed(5)
    body = rl.NatureConvBody(in_channels=4)
    atoms = np.linspace(-10, 10, 51)
    atoms_t = rl.tensor(atoms)
    net, q_fn = {
        "vanilla": (lambda: rl.VanillaNet(4, body), lambda p: p["q"]),
        "dueling": (lambda: rl.DuelingNet(4, body), lambda p: p["q"]),
        "categorical": (lambda: rl.CategoricalNet(4, 51, body), lambda p: (p["prob"] * atoms_t).sum(-1)),
        "quantile": (lambda: rl.QuantileNet(4, 50, body), lambda p: p["quantile"].mean(-1)),
    }[kind]
    net = net()
    ga = GraphedQActor(net, q_fn, n_env, 4, (84, 84), 1.0 / 255)
    rng = np.random.default_rng(3)
    for trial in range(3):                                # replays of one capture on fresh inputs
        stacks = rng.integers(0, 256, (n_env, 4, 84, 84), dtype=np.uint8)
        q = ga.q_values(list(stacks))
        assert q.shape == (n_env, 4) and q.dtype == np.float32
        # (i) the same kernels, eagerly, on the layout the graph builds
        x = torch.from_numpy(stacks).cuda().view(n_env, 4, 21, 4, 21, 4).permute(0, 2, 4, 1, 3, 5).reshape(n_env, 21, 21, 64)
        assert torch.equal(ga.x.float(), x.float())      # exact integers, channel = f*16 + dy*4 + dx
        with torch.no_grad(), frame_scale(1.0 / 255):
            q_eager = q_fn(net(x.to(torch.bfloat16).permute(0, 3, 1, 2))).float().cpu().numpy()
        # (fc4's split-K partials meet through fp32 atomics: the order, hence the last bf16 bit of a feature, can differ)
        np.testing.assert_allclose(q, q_eager, rtol=0, atol=5e-3 * max(1.0, float(np.abs(q_eager).max())))
        # (ii) the reference's arithmetic in fp32
        q_ref = oracle_q(net, kind, stacks, atoms).numpy()
        np.testing.assert_allclose(q, q_ref, rtol=0, atol=3e-2 * max(1.0, float(np.abs(q_ref).max())))
    assert ga.replays == 3


@pytest.mark.parametrize("agent_kind", ["dqn", "c51"])
def test_agent_step_uses_the_graphed_actor(rl, agent_kind):
    c = rl.Config()
    c.merge(dict(tag=None))
    c.task_fn = lambda: rl.Task("SyntheticAtari-v0", seed=2)
    c.eval_env = rl.Task("SyntheticAtari-v0", seed=2)
    c.random_action_prob = rl.LinearSchedule(1.0, 0.01, 1e6)
    c.batch_size = 32
    c.replay_fn = lambda: rl.ReplayWrapper(rl.UniformReplay, dict(memory_size=2000, batch_size=32, n_step=1, discount=0.99,
                                                                  history_length=4), async_=False)
    c.state_normalizer, c.reward_normalizer = rl.ImageNormalizer(), rl.SignNormalizer()
    c.discount, c.history_length, c.double_q, c.n_step = 0.99, 4, False, 1
    c.target_network_update_freq, c.exploration_steps, c.sgd_update_frequency, c.gradient_clip = 20, 200, 4, 5
    c.async_actor = False
    c.cuda_graph = True
    if agent_kind == "dqn":
        c.optimizer_fn = lambda p: torch.optim.RMSprop(p, lr=0.00025, alpha=0.95, eps=0.01, centered=True)
        c.network_fn = lambda: rl.VanillaNet(c.action_dim, rl.NatureConvBody(in_channels=4))
        ag = rl.DQNAgent(c)
    else:
        c.optimizer_fn = lambda p: torch.optim.Adam(p, lr=0.00025, eps=0.01 / 32)
        c.categorical_v_min, c.categorical_v_max, c.categorical_n_atoms = -10, 10, 51
        c.network_fn = lambda: rl.CategoricalNet(c.action_dim, c.categorical_n_atoms, rl.NatureConvBody(in_channels=4))
        ag = rl.CategoricalDQNAgent(c)
    np.random.seed(0)
    for _ in range(40):
        ag.step()
    ga = ag.actor._graph_actor
    assert ga and ga.replays == 160
    sig0 = ga._sig
    for _ in range(60):                                   # exploration ends at step 50: the learner is created and captured
        ag.step()
    torch.cuda.synchronize()
    assert getattr(ag, "_learner", None) is not None and ag._learner.updates > 30
    assert ga._sig != sig0 and ga._sig[0] is False        # re-captured: the optimizer kernel now maintains the packed operands
    assert ga.replays == 400
    # the captured forward tracks the fp32 master weights after the updates
    stacks = np.stack([np.asarray(s) for s in ag.actor._state])
    q = ga.q_values(list(stacks))
    kind = "vanilla" if agent_kind == "dqn" else "categorical"
    q_ref = oracle_q(ag.network, kind, stacks, np.linspace(-10, 10, 51)).numpy()
    np.testing.assert_allclose(q, q_ref, rtol=0, atol=3e-2 * max(1.0, float(np.abs(q_ref).max())))
    ag.close()


@pytest.mark.parametrize("replay_cls", ["UniformReplay", "PrioritizedReplay"])
def test_feed_many_equals_sequential_feeds(rl, replay_cls):
    """``feed_many`` (one staging upload per agent step) leaves ring, cursors and sum tree exactly as ``feed`` per env step
    (replay.py:75-90, 160-162) -- including the wrap-around of a full ring and multi-item calls (the reference's quirk)."""
    rng = np.random.default_rng(1)
    cls = getattr(rl, replay_cls)
    a = cls(memory_size=37, batch_size=8, n_step=1, discount=0.99, history_length=4)
    b = cls(memory_size=37, batch_size=8, n_step=1, discount=0.99, history_length=4)
    for step in range(30):
        group = []
        for _ in range(int(rng.integers(1, 6))):
            n = int(rng.integers(1, 4))
            group.append(dict(state=rng.integers(0, 256, (n, 84, 84), dtype=np.uint8), action=rng.integers(0, 4, n),
                              reward=rng.normal(size=n), mask=rng.integers(0, 2, n)))
        for d in group:
            a.feed(d)
        b.feed_many(group)
        assert (a.pos, a._size) == (b.pos, b._size)
    torch.cuda.synchronize()
    for name in ("frames", "action", "reward", "mask", "ring_state"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    if replay_cls == "PrioritizedReplay":
        assert torch.equal(a.tree.tree, b.tree.tree) and a.tree.n_entries == b.tree.n_entries
