"""Device-side actor step (csrc/actor.cu, component/actor.py) against the host path it replaces: the host
``MeanStdNormalizer`` (pinned by tests/test_oracle_golden.py against the RunningMeanStd restatement) + the torch
``GaussianActorCriticNet`` forward (network_heads.py:173-214) on the same observations, with the same normals."""
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
    rl.Config.COMPUTE_DTYPE = torch.float32
    torch.backends.cuda.matmul.allow_tf32 = False
    return rl


def _net(rl, D=17, A=6):
    torch.manual_seed(0)
    net = rl.GaussianActorCriticNet(D, A, actor_body=rl.FCBody(D, gate=torch.tanh), critic_body=rl.FCBody(D, gate=torch.tanh))
    with torch.no_grad():
        net.std.copy_(torch.linspace(-0.5, 0.7, A))
        for p in net.parameters():
            if p.dim() == 1 and p is not net.std:
                p.uniform_(-0.1, 0.1)                      # non-zero biases (layer_init zeroes them)
        net.fc_action.weight.mul_(300.0), net.fc_critic.weight.mul_(300.0)     # layer_init(1e-3) heads: make them matter
    return net


@pytest.mark.parametrize("N", [16, 1, 64])
def test_actor_step_matches_host_normalizer_and_torch_forward(rl, N):
    from deeprl_b200.component.actor import DeviceGaussianActor, supported
    from oracle.running_mean_std import RunningMeanStd
    D, A = 17, 6
    net = _net(rl, D, A)
    host_norm, dev_norm = rl.MeanStdNormalizer(), rl.MeanStdNormalizer()
    assert supported(net, dev_norm)
    actor = DeviceGaussianActor(net, dev_norm, N)
    orc = RunningMeanStd(shape=(1, D))
    rng = np.random.RandomState(N)
    for it in range(6):
        raw = (rng.randn(N, D) * (1 + it) + 0.3 * it).astype(np.float32)
        z = torch.from_numpy(rng.randn(N, A).astype(np.float32)).cuda()
        update = it != 3                                    # one read-only step in the middle
        if update:
            orc.update(raw)
        host_norm.read_only = not update
        x = host_norm(raw)
        with torch.no_grad():
            mean = torch.tanh(net.fc_action(net.actor_body(rl.tensor(x))))
            action = mean + torch.nn.functional.softplus(net.std) * z
            ref = net(x, action)
        out = actor.step(raw, z=z, update=update)
        torch.cuda.synchronize()
        np.testing.assert_allclose(out["state"].cpu().numpy(), np.asarray(x, dtype=np.float32), rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(out["mean"].cpu().numpy(), ref["mean"].cpu().numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(out["action"].cpu().numpy(), action.cpu().numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(out["v"].cpu().numpy(), ref["v"].cpu().numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(out["entropy"].cpu().numpy(), ref["entropy"].cpu().numpy(), rtol=1e-5, atol=1e-5)
        # log-prob of the device's own action under the torch distribution
        with torch.no_grad():
            lp = net(x, out["action"])["log_pi_a"]
        np.testing.assert_allclose(out["log_pi_a"].cpu().numpy(), lp.cpu().numpy(), rtol=1e-4, atol=1e-4)
        # running moments: device (float64 Chan merge) vs the host normaliser vs the RunningMeanStd restatement
        np.testing.assert_allclose(actor.rm_mean.cpu().numpy(), host_norm.rms.mean.reshape(-1), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(actor.rm_var.cpu().numpy(), host_norm.rms.var.reshape(-1), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(actor.rm_var.cpu().numpy(), orc.var.reshape(-1), rtol=1e-6, atol=1e-7)
        assert abs(float(actor.rm_count) - orc.count) < 1e-9
    actor.pull_stats()
    np.testing.assert_allclose(dev_norm.rms.mean.reshape(-1), host_norm.rms.mean.reshape(-1), rtol=1e-6, atol=1e-7)


def test_actor_philox_sampling_statistics(rl):
    from deeprl_b200.component.actor import DeviceGaussianActor
    D, A, N = 17, 6, 64
    net = _net(rl, D, A)
    actor = DeviceGaussianActor(net, rl.MeanStdNormalizer(), N, seed=5)
    raw = np.random.RandomState(0).randn(N, D).astype(np.float32)
    zs = []
    sd = torch.nn.functional.softplus(net.std.detach())
    for _ in range(200):
        out = actor.step(raw, update=False)
        zs.append(((out["action"] - out["mean"]) / sd).cpu().numpy())
    z = np.concatenate(zs).ravel()
    assert int(actor.counter) == 200 * N * A
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs(np.mean(z ** 3)) < 0.05 and abs(np.mean(z ** 4) - 3.0) < 0.15
    assert len(np.unique(np.round(z, 6))) > 0.97 * z.size      # (birthday collisions of 76 800 draws at 1e-6 resolution)


def test_ppo_agent_uses_the_device_actor(rl):
    torch.manual_seed(0), np.random.seed(0)
    c = rl.Config()
    c.merge(dict(tag=None))
    c.num_workers = 8
    c.task_fn = lambda: rl.Task("SyntheticCheetah-v0", num_envs=8, seed=0)
    c.eval_env = rl.Task("SyntheticCheetah-v0", seed=0)
    c.network_fn = lambda: rl.GaussianActorCriticNet(c.state_dim, c.action_dim, actor_body=rl.FCBody(c.state_dim, gate=torch.tanh),
                                                     critic_body=rl.FCBody(c.state_dim, gate=torch.tanh))
    c.actor_opt_fn = lambda p: torch.optim.Adam(p, 3e-4)
    c.critic_opt_fn = lambda p: torch.optim.Adam(p, 1e-3)
    c.discount, c.use_gae, c.gae_tau, c.gradient_clip = 0.99, True, 0.95, 0.5
    c.rollout_length, c.optimization_epochs, c.mini_batch_size, c.ppo_ratio_clip, c.target_kl = 64, 2, 64, 0.2, 0.01
    c.state_normalizer = rl.MeanStdNormalizer()
    ag = rl.PPOAgent(c)
    for _ in range(2):
        ag.step()
    torch.cuda.synchronize()
    assert ag._device_actor() is not None
    # every observation batch was counted once: reset + 2 x 64 steps
    assert abs(c.state_normalizer.rms.count - (1e-4 + 8 * (1 + 2 * 64))) < 1e-6
    assert all(torch.isfinite(p).all() for p in ag.network.parameters())
    ag.close()


def test_arena_rollout_equals_the_per_step_rollout(rl):
    """``PPOAgent._rollout_device`` writing the actor's outputs into rollout-sized arenas (and uploading rewards / masks once)
    returns exactly the entries of the per-step form (same launches, same Philox stream, same env seeds)."""
    def agent(arena):
        torch.manual_seed(0), np.random.seed(0)
        c = rl.Config()
        c.merge(dict(tag=None))
        c.num_workers = 8
        c.task_fn = lambda: rl.Task("SyntheticCheetah-v0", num_envs=8, seed=3)
        c.eval_env = rl.Task("SyntheticCheetah-v0", seed=3)
        c.network_fn = lambda: rl.GaussianActorCriticNet(c.state_dim, c.action_dim,
                                                         actor_body=rl.FCBody(c.state_dim, gate=torch.tanh),
                                                         critic_body=rl.FCBody(c.state_dim, gate=torch.tanh))
        c.actor_opt_fn = lambda p: torch.optim.Adam(p, 3e-4)
        c.critic_opt_fn = lambda p: torch.optim.Adam(p, 1e-3)
        c.discount, c.use_gae, c.gae_tau, c.gradient_clip = 0.99, True, 0.95, 0.5
        c.rollout_length, c.optimization_epochs, c.mini_batch_size, c.ppo_ratio_clip, c.target_kl = 48, 2, 64, 0.2, 0.01
        c.state_normalizer = rl.MeanStdNormalizer()
        c.device_actor_arena = arena
        return rl.PPOAgent(c)

    a, b = agent(False), agent(True)
    for it in range(3):                                    # (the second / third rollout start from a state seen before)
        ea, eb = a._rollout(), b._rollout()
        torch.cuda.synchronize()
        for name, x, y in zip(ea._fields, ea, eb):
            assert torch.equal(x, y), (it, name)
        assert a.total_steps == b.total_steps
        ra, rb = a.config.state_normalizer.rms, b.config.state_normalizer.rms
        assert np.array_equal(ra.mean, rb.mean) and np.array_equal(ra.var, rb.var) and ra.count == rb.count
    assert a._device_actor() is not None and b._device_actor() is not None
    a.close(), b.close()
