"""Pin ``oracle/`` (the CPU restatement) against the golden vectors that tests/golden/make_golden.py
produced by running the REFERENCE ITSELF.  CPU only.  Integer / index / fp64-tree results are
compared bit-for-bit; fp32 losses and parameters at the tolerance written beside each check."""
import random

import numpy as np
import pytest
import torch

from oracle import agents, losses, nets
from oracle.replay import PrioritizedReplay, UniformReplay
from oracle.running_mean_std import RunningMeanStd
from oracle.sum_tree import SumTree

torch.set_num_threads(1)


# ------------------------------------------------------------------------------------------ sum tree
@pytest.mark.parametrize("cap", [2, 5, 8, 1000])
def test_sumtree_trace_bit_exact(golden, cap):
    g = golden("sumtree")
    t = SumTree(cap)
    ops, a0, a1 = g["cap%d_ops" % cap], g["cap%d_a0" % cap], g["cap%d_a1" % cap]
    snaps = g["cap%d_snaps" % cap] if cap <= 8 else None
    for k, op in enumerate(ops):
        if op == 0:
            t.add(a0[k])
        elif op == 1:
            idx, p, di = t.get(a0[k])
            assert (idx, di) == (g["cap%d_res_idx" % cap][k], g["cap%d_res_data" % cap][k])
            assert p == g["cap%d_res_p" % cap][k]
        else:
            t.update(int(a0[k]), a1[k])
        if snaps is not None:
            assert np.array_equal(t.tree, snaps[k])
    assert np.array_equal(t.tree, g["cap%d_tree" % cap])
    assert sorted(t.pending) == list(g["cap%d_pending" % cap])


def test_sumtree_batched_rounds_bit_exact(golden):
    g = golden("sumtree")
    cap, B = 1000, g["batch_u"].shape[1]
    t = SumTree(cap)
    for _ in range(cap):
        t.add(1.0)
    for it in range(g["batch_u"].shape[0]):
        seg = t.total() / B
        u = g["batch_u"][it]
        got = [t.get(seg * i + (seg * (i + 1) - seg * i) * u[i]) for i in range(B)]
        assert np.array_equal([x[1] for x in got], g["batch_p"][it])
        for i, p in zip(g["batch_idx"][it], g["batch_prio"][it]):
            t.update(int(i), p)
        assert np.array_equal(t.tree, g["batch_trees"][it])


# ------------------------------------------------------------------------------------------ uniform replay
def _feed_all(rp, g, pre, upto=None):
    fr, ac, rw, mk = g[pre + "frames"], g[pre + "actions"], g[pre + "rewards"], g[pre + "masks"]
    for i in range(len(fr) if upto is None else upto):
        rp.feed(dict(state=[fr[i]], action=[ac[i]], reward=[rw[i]], mask=[mk[i]]))


def test_uniform_replay_matches_reference(golden):
    g = golden("replay_uniform")
    for c in range(int(g["n_cases"])):
        pre = "u%d_" % c
        M, hl, n, feeds, B = g[pre + "cfg"]
        rp = UniformReplay(M, B, n, float(g[pre + "discount"]), hl)
        _feed_all(rp, g, pre)
        assert (rp.pos, rp.size()) == (g[pre + "pos"], g[pre + "size"])
        valid = np.asarray([rp.valid_index(i) for i in range(rp.size())])
        assert np.array_equal(valid, g[pre + "valid"])
        for k, i in enumerate(g[pre + "valid_idx"]):
            tr = rp.construct_transition(int(i))
            assert np.array_equal(tr.state, g[pre + "tr_state"][k]) and np.array_equal(tr.next_state, g[pre + "tr_next"][k])
            assert tr.action == g[pre + "tr_action"][k] and tr.reward == g[pre + "tr_reward"][k]
            assert int(tr.mask) == g[pre + "tr_mask"][k]
        smp, taken, used = rp.sample(candidates=g[pre + "cand"])
        assert np.array_equal(smp.state, g[pre + "s_state"]) and np.array_equal(smp.next_state, g[pre + "s_next"])
        assert np.array_equal(smp.action, g[pre + "s_action"]) and np.array_equal(smp.reward, g[pre + "s_reward"])
        assert np.array_equal(np.asarray(smp.mask, np.int64), g[pre + "s_mask"])


def test_uniform_feed_quirk(golden):
    g = golden("replay_uniform")
    rp = UniformReplay(4, 1)
    rp.feed(dict(state=[0, 1, 2, 3], action=[0, 1, 2, 3], reward=[0, 1, 2, 3], mask=[1, 1, 1, 1]))
    rp.feed(dict(state=[10, 11], action=[10, 11], reward=[10, 11], mask=[1, 1]))
    assert np.array_equal(np.asarray(rp.data["state"]), g["quirk_state"])
    assert (rp.pos, rp.size()) == (g["quirk_pos"], g["quirk_size"])


# ------------------------------------------------------------------------------------------ prioritized replay
def test_prioritized_replay_matches_reference(golden):
    g = golden("replay_per")
    for c in range(int(g["n_cases"])):
        pre = "p%d_" % c
        M, hl, n, feeds, B, rounds = g[pre + "cfg"]
        fr, ac, rw, mk = g[pre + "frames"], g[pre + "actions"], g[pre + "rewards"], g[pre + "masks"]
        rp = PrioritizedReplay(M, B, n, float(g[pre + "discount"]), hl)
        _feed_all(rp, g, pre)
        for rd in range(rounds):
            assert rp.tree.total() == g[pre + "total"][rd]
            smp = rp.sample(uniforms=g[pre + "u"][rd], fills=g[pre + "fills"][rd])
            for k in ("state", "next_state", "action", "reward", "mask", "sampling_prob", "idx"):
                assert np.array_equal(np.asarray(getattr(smp, k)), g[pre + "s_" + k][rd]), (c, rd, k)
            idx = np.asarray(smp.idx, np.float32).astype(np.int64)
            rp.update_priorities(zip(idx, g[pre + "prio"][rd]))
            base = feeds + rd * 2
            for j in range(2):
                k2 = (base + j) % feeds
                rp.feed(dict(state=[fr[k2]], action=[ac[k2]], reward=[rw[k2]], mask=[mk[k2]]))
            assert np.array_equal(rp.tree.tree, g[pre + "tree"][rd])
            assert rp.max_priority == g[pre + "max_priority"][rd]


# ------------------------------------------------------------------------------------------ loss boundary
def _t(x, dtype=torch.float32):
    return torch.from_numpy(np.asarray(x)).to(dtype)


def test_loss_boundary_matches_reference(golden):
    g = golden("losses")
    a, r, m = _t(g["action"], torch.long), _t(g["reward"]), _t(g["mask"])
    q, qt, qo = _t(g["dqn_q"]), _t(g["dqn_qn_t"]), _t(g["dqn_qn_o"])
    for double in (0, 1):
        for n in (1, 3):
            d = losses.dqn_delta(q, qt, qo if double else None, a, r, m, 0.99 ** n)
            np.testing.assert_allclose(d.numpy(), g["dqn_d%d_n%d_delta" % (double, n)], rtol=0, atol=0)
            np.testing.assert_allclose(losses.dqn_reduce(d).numpy(), g["dqn_d%d_n%d_loss" % (double, n)], rtol=1e-7)
    lp, pt, po, atoms = _t(g["c51_logp"]), _t(g["c51_pn_t"]), _t(g["c51_pn_o"]), _t(g["c51_atoms"])
    for double in (0, 1):
        kl = losses.c51_kl(lp, pt, po if double else None, a, r, m, atoms, -10, 10, 0.99)
        np.testing.assert_allclose(kl.numpy(), g["c51_d%d_kl" % double], rtol=1e-6, atol=1e-6)
    v = losses.qr_loss(_t(g["qr_quant"]), _t(g["qr_qn"]), a, r, m, 0.99)
    np.testing.assert_allclose(v.numpy(), g["qr_vec"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(v.mean().numpy(), g["qr_loss"], rtol=1e-6)


# ------------------------------------------------------------------------------------------ DQN-family update steps
class _Batch:
    def __init__(self, g, pre, i, fields):
        for f in fields:
            setattr(self, f, g[pre + "b_" + f][i])


@pytest.mark.parametrize("name", ["dqn_per", "dqn_uni", "c51", "qr"])
def test_dqn_family_update_trajectory(golden, name):
    """oracle.agents.DQNFamilyOracle.update, fed the batches the reference sampled, reproduces the
    reference's delta vector and parameter trajectory (DQN_agent.py:115-138).  Tolerance 2e-6
    absolute on parameters after up to ~25 optimizer steps (same torch ops, same order)."""
    g = golden("agent_steps")
    pre = name + "_"
    keys = [str(k) for k in g[pre + "keys"]]
    sd = {k: _t(g[pre + "init." + k]) for k in keys}
    head = dict(dqn_per="dueling", dqn_uni="vanilla", c51="categorical", qr="quantile")[name]
    if name.startswith("dqn"):
        opt_fn = lambda p: torch.optim.RMSprop(p, lr=0.00025, alpha=0.95, eps=0.01, centered=True)
    elif name == "c51":
        opt_fn = lambda p: torch.optim.Adam(p, lr=0.00025, eps=0.01 / 32)
    else:
        opt_fn = lambda p: torch.optim.Adam(p, lr=0.00005, eps=0.01 / 32)
    from deeprl_b200.utils.schedule import LinearSchedule
    orc = agents.DQNFamilyOracle(
        sd, head, "fc", 2, opt_fn, 0.99, 1, double_q=(name == "dqn_per"), gradient_clip=5,
        atoms=np.linspace(-100, 100, 50) if name == "c51" else None, v_min=-100, v_max=100,
        num_quantiles=20 if name == "qr" else None, replay_beta=LinearSchedule(0.4, 1.0, 200))
    fields = ["state", "action", "reward", "next_state", "mask"] + (["sampling_prob", "idx"] if name == "dqn_per" else [])
    nb = g[pre + "delta"].shape[0]
    for i in range(nb):
        b = _Batch(g, pre, i, fields)
        d = orc.compute_loss(b)
        np.testing.assert_allclose(d.detach().numpy(), g[pre + "delta"][i], rtol=1e-5, atol=2e-6)
        orc.update(b)
        flat = np.concatenate([orc.sd[k].detach().numpy().ravel() for k in _param_order(keys, head)])
        np.testing.assert_allclose(flat, g[pre + "params"][i], rtol=0, atol=2e-6)
        tflat = np.concatenate([orc.target_sd[k].numpy().ravel() for k in _param_order(keys, head)])
        if not np.allclose(tflat, g[pre + "target"][i], atol=2e-6):
            orc.sync_target()                                       # DQN_agent.py:136-138 fired on this step
            tflat = np.concatenate([orc.target_sd[k].numpy().ravel() for k in _param_order(keys, head)])
            np.testing.assert_allclose(tflat, g[pre + "target"][i], rtol=0, atol=2e-6)


def _param_order(keys, head):
    """nn.Module.parameters() order of the reference nets: head linear(s) are registered BEFORE the
    body (network_heads.py:14-15,27-29,43-46,92-95)."""
    return keys


# ------------------------------------------------------------------------------------------ on-policy
def test_gae_matches_reference(golden):
    g = golden("onpolicy")
    for (T, N) in ((128, 8), (2048, 16)):
        k = "gae_%d_%d_" % (T, N)
        adv, ret = losses.gae(_t(g[k + "reward"]), _t(g[k + "mask"]), _t(g[k + "v"]), 0.99, 0.95)
        assert np.array_equal(adv.numpy(), g[k + "adv"]) and np.array_equal(ret.numpy(), g[k + "ret"])
    for it in range(g["a2c_gae_reward"].shape[0]):                  # captured inside A2CAgent.step, (5, 8)
        adv, ret = losses.gae(_t(g["a2c_gae_reward"][it]), _t(g["a2c_gae_mask"][it]), _t(g["a2c_gae_v"][it]), 0.99, 0.95)
        assert np.array_equal(adv.numpy(), g["a2c_gae_advantage"][it]) and np.array_equal(ret.numpy(), g["a2c_gae_ret"][it])


def test_a2c_step_trajectory(golden):
    """A2C_agent.py:22-64 on CartPole (8 workers, rollout 5): parameters after each of 6 steps, atol 2e-6."""
    g = golden("onpolicy")
    keys = [str(k) for k in g["a2c_keys"]]
    sd = agents.leafify({k: _t(g["a2c_init." + k]) for k in keys})
    params = [sd[k] for k in keys]
    opt = torch.optim.RMSprop(params, 0.001)
    T, iters = 5, g["a2c_params"].shape[0]
    state = _t(g["a2c_state0"])
    for it in range(iters):
        sl = slice(it * T, (it + 1) * T)
        states = torch.cat([state[None], _t(g["a2c_next_states"][sl])])
        actions = _t(g["a2c_actions"][sl], torch.long)
        rewards = _t(g["a2c_rewards"][sl]).unsqueeze(-1)
        masks = _t(1 - g["a2c_dones"][sl].astype(np.int64)).unsqueeze(-1)
        adv, ret = agents.a2c_update(sd, params, opt, states, actions, rewards, masks, 0.99, 0.95, 0.01, 1.0, 0.5)
        np.testing.assert_allclose(adv.numpy(), g["a2c_gae_advantage"][it], rtol=1e-5, atol=1e-6)
        flat = np.concatenate([p.detach().numpy().ravel() for p in params])
        np.testing.assert_allclose(flat, g["a2c_params"][it], rtol=0, atol=2e-6)
        state = states[-1]


def test_ppo_step_trajectory(golden):
    """PPO_agent.py:63-99 (non-shared): same minibatch permutations -> same parameters, atol 2e-6."""
    g = golden("onpolicy")
    pre = "ppo_small_"
    T, N, mb, epochs, its = g[pre + "cfg"]
    keys = [str(k) for k in g[pre + "keys"]]
    names = [str(k) for k in g[pre + "param_names"]]
    sd = agents.leafify({k: _t(g[pre + "init." + k]) for k in keys})
    actor = [sd[k] for k in names if k.startswith(("actor_body", "fc_action"))] + [sd["std"]]
    critic = [sd[k] for k in names if k.startswith(("critic_body", "fc_critic"))]
    a_opt, c_opt = torch.optim.Adam(actor, 3e-4), torch.optim.Adam(critic, 1e-3)
    real = np.random.permutation
    try:
        for it in range(its):
            perms = list(g[pre + "perms"][it])
            np.random.permutation = lambda x, perms=perms: perms.pop(0)
            f = lambda k: _t(g[pre + k][it]).reshape(T * N, -1)
            adv, ret = losses.gae(_t(g[pre + "reward"][it]), _t(g[pre + "mask"][it]), _t(g[pre + "v"][it]), 0.99, 0.95)
            assert np.array_equal(adv.numpy(), g[pre + "advantage"][it])
            agents.ppo_update(sd, None, None, a_opt, c_opt, f("state"), f("action"), f("log_pi_a"),
                              ret.reshape(-1, 1), adv.reshape(-1, 1), epochs, mb, 0.2, 0, 0.01)
            flat = np.concatenate([sd[k].detach().numpy().ravel() for k in names])
            np.testing.assert_allclose(flat, g[pre + "params"][it], rtol=0, atol=2e-6)
    finally:
        np.random.permutation = real


def test_running_mean_std_closed_form():
    """baselines RunningMeanStd restatement vs closed-form batch statistics (known-answer test)."""
    rng = np.random.RandomState(0)
    xs = [rng.randn(n, 3) * 2 + 1 for n in (7, 1, 64, 13)]
    r = RunningMeanStd(shape=(1, 3), epsilon=1e-4)
    for x in xs:
        r.update(x)
    allx = np.concatenate(xs)
    n = len(allx)
    # the initial pseudo-count (mean 0, var 1, weight 1e-4) is part of the definition
    mean = allx.sum(0) / (n + 1e-4)
    ex2 = ((allx ** 2).sum(0) + 1e-4 * 1.0) / (n + 1e-4)
    np.testing.assert_allclose(r.mean[0], mean, rtol=1e-12)
    np.testing.assert_allclose(r.var[0], ex2 - mean ** 2, rtol=1e-10)
    assert abs(r.count - (n + 1e-4)) < 1e-9
