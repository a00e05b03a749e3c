"""n-step Q-learning (reference ``agent/NStepDQN_agent.py:26-70``) against the golden trajectory recorded from the
UNMODIFIED reference (tests/golden/make_golden_nstep.py -> nstep.npz: 16 steps of 5 workers x rollout 5 on the synthetic
CartPole task, target sync every 12 env steps, 5 terminals):

* the oracle's restatement (oracle/agents.py nstep_dqn_update) reproduces returns and parameters   -- CPU
* the product's NStepDQNAgent on the CPU device reproduces the parameters                          -- CPU
* the product's CUDA path (K7 scan kernel for the returns, fused DQN loss kernel)                  -- GPU
"""
import numpy as np
import pytest
import torch

import deeprl_b200 as rl
from oracle import agents

T, N, FREQ = 5, 5, 12


def _t(x, dtype=torch.float32):
    return torch.from_numpy(np.asarray(x)).to(dtype)


def test_oracle_nstep_trajectory(golden):
    g = golden("nstep")
    keys = [str(k) for k in g["keys"]]
    sd = agents.leafify({k: _t(g["init." + k]) for k in keys})
    tgt = {k: v.detach().clone() for k, v in sd.items()}
    params = [sd[k] for k in keys]
    opt = torch.optim.RMSprop(params, 0.001)
    state = _t(g["state0"])
    env_steps = 0
    for it in range(g["params"].shape[0]):
        sl = slice(it * T, (it + 1) * T)
        states = torch.cat([state[None], _t(g["next_states"][sl])])
        for _ in range(T):                               # NStepDQN_agent.py:47-49: sync inside the rollout, before the update
            env_steps += 1
            if env_steps % FREQ == 0:
                tgt = {k: v.detach().clone() for k, v in sd.items()}
        ret, _ = agents.nstep_dqn_update(sd, tgt, params, opt, states, _t(g["actions"][sl], torch.long),
                                         _t(g["rewards"][sl]).unsqueeze(-1), _t(1 - g["dones"][sl].astype(np.int64)).unsqueeze(-1),
                                         0.99, 5)
        np.testing.assert_allclose(ret.detach().numpy(), g["cap_ret"][it], rtol=1e-6, atol=1e-6)
        flat = np.concatenate([p.detach().numpy().ravel() for p in params])
        np.testing.assert_allclose(flat, g["params"][it], rtol=0, atol=2e-6)
        flat_t = np.concatenate([tgt[k].numpy().ravel() for k in keys])
        np.testing.assert_allclose(flat_t, g["target_params"][it], rtol=0, atol=2e-6)
        state = states[-1]


def _replayed_agent(g, monkeypatch):
    keys = [str(k) for k in g["keys"]]

    class Replay:                                        # Task stand-in that replays the recorded env stream
        def __init__(self):
            self.k = 0
            self.state_dim, self.action_dim, self.name = 4, 2, "replayed"

        def reset(self):
            return list(g["state0"])

        def step(self, actions):
            k = self.k
            self.k += 1
            return list(g["next_states"][k]), g["rewards"][k], g["dones"][k], tuple({"episodic_return": None} for _ in range(N))

        def close(self):
            pass

    c = rl.Config()
    c.merge(dict(tag=None))
    c.num_workers = N
    c.task_fn = Replay
    c.optimizer_fn = lambda p: torch.optim.RMSprop(p, 0.001)
    c.network_fn = lambda: rl.VanillaNet(2, rl.FCBody(4))
    c.random_action_prob = rl.LinearSchedule(0.6, 0.1, 200)
    c.discount, c.target_network_update_freq, c.rollout_length, c.gradient_clip = 0.99, FREQ, T, 5
    ag = rl.NStepDQNAgent(c)
    ag.network.load_state_dict({k: torch.from_numpy(g["init." + k]) for k in keys})
    ag.target_network.load_state_dict(ag.network.state_dict())
    # the recorded actions (epsilon-greedy draws of the reference run) are part of the record
    from deeprl_b200.agent import NStepDQN_agent as mod
    monkeypatch.setattr(mod, "epsilon_greedy", lambda eps, q: g["actions"][ag.task.k])
    return ag


def test_product_cpu_matches_reference_trajectory(golden, monkeypatch):
    rl.select_device(-1)
    g = golden("nstep")
    ag = _replayed_agent(g, monkeypatch)
    for it in range(g["params"].shape[0]):
        ag.step()
        flat = np.concatenate([p.detach().numpy().ravel() for p in ag.network.parameters()])
        np.testing.assert_allclose(flat, g["params"][it], rtol=0, atol=2e-6)
        flat_t = np.concatenate([p.detach().numpy().ravel() for p in ag.target_network.parameters()])
        np.testing.assert_allclose(flat_t, g["target_params"][it], rtol=0, atol=2e-6)
    assert ag.total_steps == g["params"].shape[0] * T * N


def test_epsilon_greedy_draws_match_reference_record(golden):
    """Without replaying the actions: same numpy seed + same network -> the first rollout's actions of the record (the draw
    order of torch_utils.py:51-58 is randint, then rand).  The synthetic task draws from its own generator."""
    rl.select_device(-1)
    g = golden("nstep")
    keys = [str(k) for k in g["keys"]]
    np.random.seed(21), torch.manual_seed(21)
    c = rl.Config()
    c.merge(dict(tag=None))
    c.num_workers = N
    c.task_fn = lambda: rl.Task("CartPole-v0", num_envs=N, seed=8)
    c.optimizer_fn = lambda p: torch.optim.RMSprop(p, 0.001)
    c.network_fn = lambda: rl.VanillaNet(2, rl.FCBody(4))
    c.random_action_prob = rl.LinearSchedule(0.6, 0.1, 200)
    c.discount, c.target_network_update_freq, c.rollout_length, c.gradient_clip = 0.99, FREQ, T, 5
    ag = rl.NStepDQNAgent(c)
    np.testing.assert_array_equal(np.asarray(ag.states, np.float32), g["state0"])
    ag.network.load_state_dict({k: torch.from_numpy(g["init." + k]) for k in keys})
    ag.target_network.load_state_dict(ag.network.state_dict())
    seen = []
    real = ag.task.step
    ag.task.step = lambda a: (seen.append(np.asarray(a).copy()), real(a))[1]
    ag.step()
    np.testing.assert_array_equal(np.stack(seen), g["actions"][:T])
    flat = np.concatenate([p.detach().numpy().ravel() for p in ag.network.parameters()])
    np.testing.assert_allclose(flat, g["params"][0], rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_product_cuda_matches_reference_trajectory(golden, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    rl.select_device(0)
    rl.Config.COMPUTE_DTYPE = torch.float32
    g = golden("nstep")
    ag = _replayed_agent(g, monkeypatch)
    assert next(ag.network.parameters()).is_cuda
    for it in range(g["params"].shape[0]):
        ag.step()
        flat = np.concatenate([p.detach().cpu().numpy().ravel() for p in ag.network.parameters()])
        # fp32 GEMMs on the device accumulate in another order than MKL: 1e-4 on parameters of O(1) after up to 16 RMSprop steps
        np.testing.assert_allclose(flat, g["params"][it], rtol=0, atol=1e-4)
    assert torch.isfinite(ag.last_loss)


def test_n_step_dqn_feature_launcher_cpu():
    """examples.py:408-424 wiring through ``run_steps`` on the CPU device."""
    import examples
    rl.select_device(-1)
    rl.random_seed(0)
    examples.n_step_dqn_feature(game="CartPole-v0", max_steps=5 * 5 * 30, log_interval=0, tag=None)
