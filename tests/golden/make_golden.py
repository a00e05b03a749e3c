"""Generate the golden fixtures in tests/golden/*.npz FROM THE REFERENCE ITSELF.

Run by hand in the build container (``python tests/golden/make_golden.py``): it imports the
unmodified reference from /root/reference through ``oracle/ref_shim.py`` (in-memory, no source
copy), drives its own classes on seeded synthetic inputs and stores inputs + outputs.  The
reference ships no tests / golden vectors, so these files are what pins
``oracle/`` (tests/test_oracle_golden.py) and, through it, the CUDA kernels.  /root/reference
does not exist on the GPU box: only the .npz files travel.

Reference entry points exercised (file:line under /root/reference/deep_rl):
  utils/sum_tree.py:6-67              SumTree.add / get / update / total
  component/replay.py:57-149          UniformReplay.feed / valid_index / construct_transition / sample
  component/replay.py:152-196         PrioritizedReplay.feed / sample / update_priorities
  agent/DQN_agent.py:78-138           DQNAgent.compute_loss / reduce_loss / step (PER block, clip, opt)
  agent/CategoricalDQN_agent.py:60-89 CategoricalDQNAgent.compute_loss / reduce_loss
  agent/QuantileRegressionDQN_agent.py:55-77
  agent/A2C_agent.py:22-64            A2CAgent.step
  agent/PPO_agent.py:29-99            PPOAgent.step
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shim import import_reference  # noqa: E402

ref = import_reference()
from deeprl_b200.component.envs import Task as SynthTask  # noqa: E402  (host env, duck-typed for the reference)

torch.set_num_threads(1)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-24s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


class NullLogger:
    def info(self, *a, **k): pass
    debug = warning = add_scalar = add_histogram = info


# --------------------------------------------------------------------------------------- sum tree
def gen_sumtree():
    out = {}
    for cap in (2, 5, 8, 1000):
        rng = np.random.RandomState(cap)
        t = ref.SumTree(cap)
        ops, a0, a1, res = [], [], [], []
        snaps = []
        n_ops = 60 if cap <= 8 else 6000
        for k in range(n_ops):
            r = rng.rand()
            if k < cap or r < 0.3:                                   # add (fills the ring first)
                p = float(rng.rand() * 3 + 0.01)
                t.add(p, None)
                ops.append(0), a0.append(p), a1.append(0.0), res.append((0, 0.0, 0))
            elif r < 0.65 and cap > 1:                               # get
                s = float(rng.rand() * t.total())
                idx, p, di = t.get(s)
                ops.append(1), a0.append(s), a1.append(0.0), res.append((idx, float(p), di))
            elif cap > 1:                                            # update (maybe not pending, maybe duplicate)
                idx = int(rng.randint(cap - 1, 2 * cap - 1))
                p = float(np.float32(rng.rand() * 2 + 0.01))
                t.update(idx, p)
                ops.append(2), a0.append(float(idx)), a1.append(p), res.append((0, 0.0, 0))
            else:
                continue
            if cap <= 8:
                snaps.append(t.tree.copy())
        out["cap%d_ops" % cap] = np.asarray(ops, np.int64)
        out["cap%d_a0" % cap] = np.asarray(a0, np.float64)
        out["cap%d_a1" % cap] = np.asarray(a1, np.float64)
        out["cap%d_res_idx" % cap] = np.asarray([r[0] for r in res], np.int64)
        out["cap%d_res_p" % cap] = np.asarray([r[1] for r in res], np.float64)
        out["cap%d_res_data" % cap] = np.asarray([r[2] for r in res], np.int64)
        out["cap%d_tree" % cap] = t.tree.copy()
        out["cap%d_pending" % cap] = np.asarray(sorted(t.pending_idx), np.int64)
        if snaps:
            out["cap%d_snaps" % cap] = np.stack(snaps)
    # batch sample/update trace in the shape the CUDA kernel sees it: B stratified gets then B updates
    cap, B = 1000, 64
    rng = np.random.RandomState(7)
    t = ref.SumTree(cap)
    for _ in range(cap):
        t.add(1.0, None)
    trees, idxs, ps, us, prios = [], [], [], [], []
    for it in range(12):
        total = t.total()
        seg = total / B
        u = rng.rand(B)
        got = [t.get(seg * i + (seg * (i + 1) - seg * i) * u[i]) for i in range(B)]
        idx = np.asarray([g[0] for g in got], np.int64)
        if it % 3 == 2:
            idx[B // 2:] = idx[:B // 2]                              # force duplicates inside one batch
        pr = (np.abs(rng.randn(B)).astype(np.float32) + np.float32(0.01)) ** np.float32(0.5)
        for i, p in zip(idx, pr):
            t.update(int(i), p)                                      # p is np.float32, as from to_np(priorities)
        us.append(u), idxs.append(idx), ps.append([g[1] for g in got]), prios.append(pr), trees.append(t.tree.copy())
    out.update(batch_u=np.stack(us), batch_idx=np.stack(idxs), batch_p=np.asarray(ps, np.float64),
               batch_prio=np.stack(prios), batch_trees=np.stack(trees))
    save("sumtree", **out)


# --------------------------------------------------------------------------------------- uniform replay
def _feed_stream(rng, n, frame_shape):
    frames = rng.randint(0, 256, size=(n,) + frame_shape).astype(np.uint8)
    actions = rng.randint(0, 6, size=n)
    rewards = rng.choice([-1.0, 0.0, 1.0], size=n, p=[0.2, 0.6, 0.2])
    masks = (rng.rand(n) > 0.1).astype(np.int32)
    return frames, actions, rewards, masks


def gen_uniform():
    out = {}
    cases = [  # (memory_size, history, n_step, discount, n_feeds, batch)
        (16, 1, 1, 0.99, 10, 4),      # not full
        (16, 4, 1, 0.99, 16, 8),      # exactly full, pos = 0
        (16, 4, 1, 0.99, 23, 8),      # wrapped, pos mid
        (32, 4, 3, 0.9, 75, 8),       # n-step 3, wrapped twice
        (32, 2, 2, 1.0, 31, 8),       # pos = size-1 edge
        (64, 4, 1, 0.99, 100, 16),
    ]
    for c, (M, hl, n, disc, feeds, B) in enumerate(cases):
        rng = np.random.RandomState(100 + c)
        fr, ac, rw, mk = _feed_stream(rng, feeds, (3, 3))
        rp = ref.UniformReplay(memory_size=M, batch_size=B, n_step=n, discount=disc, history_length=hl)
        for i in range(feeds):
            rp.feed(dict(state=[fr[i]], action=[ac[i]], reward=[rw[i]], mask=[mk[i]]))
        size = rp.size()
        valid = np.asarray([rp.valid_index(i) for i in range(size)])
        cvi = rp.compute_valid_indices()
        vi = np.nonzero(valid)[0]
        trs = [rp.construct_transition(int(i)) for i in vi]
        np.random.seed(1000 + c)
        smp = rp.sample()
        np.random.seed(1000 + c)
        cand = np.asarray([np.random.randint(0, size) for _ in range(64 * B)], np.int64)
        pre = "u%d_" % c
        out.update({
            pre + "cfg": np.asarray([M, hl, n, feeds, B], np.int64), pre + "discount": np.float64(disc),
            pre + "frames": fr, pre + "actions": ac, pre + "rewards": rw, pre + "masks": mk,
            pre + "pos": np.int64(rp.pos), pre + "size": np.int64(size), pre + "valid": valid,
            pre + "compute_valid_indices": cvi, pre + "valid_idx": vi,
            pre + "tr_state": np.stack([t.state for t in trs]), pre + "tr_next": np.stack([t.next_state for t in trs]),
            pre + "tr_action": np.asarray([t.action for t in trs]), pre + "tr_reward": np.asarray([t.reward for t in trs], np.float64),
            pre + "tr_mask": np.asarray([t.mask for t in trs], np.int64),
            pre + "cand": cand, pre + "s_state": smp.state, pre + "s_next": smp.next_state,
            pre + "s_action": smp.action, pre + "s_reward": np.asarray(smp.reward, np.float64),
            pre + "s_mask": np.asarray(smp.mask, np.int64),
        })
    # the multi-element-feed quirk (replay.py:87): storage[self.pos] instead of storage[pos] once full
    rp = ref.UniformReplay(memory_size=4, batch_size=1)
    rp.feed(dict(state=[0, 1, 2, 3], action=[0, 1, 2, 3], reward=[0, 1, 2, 3], mask=[1, 1, 1, 1]))
    rp.feed(dict(state=[10, 11], action=[10, 11], reward=[10, 11], mask=[1, 1]))
    out.update(quirk_state=np.asarray(rp.state), quirk_pos=np.int64(rp.pos), quirk_size=np.int64(rp.size()))
    out["n_cases"] = np.int64(len(cases))
    save("replay_uniform", **out)


# --------------------------------------------------------------------------------------- prioritized replay
def gen_per():
    out = {}
    cases = [(32, 4, 1, 0.99, 32, 8), (32, 4, 1, 0.99, 50, 8), (64, 1, 1, 0.99, 40, 16), (64, 4, 3, 0.95, 150, 16)]
    real_choice = random.choice
    for c, (M, hl, n, disc, feeds, B) in enumerate(cases):
        rng = np.random.RandomState(200 + c)
        fr, ac, rw, mk = _feed_stream(rng, feeds, (3, 3))
        rp = ref.PrioritizedReplay(memory_size=M, batch_size=B, n_step=n, discount=disc, history_length=hl)
        for i in range(feeds):
            rp.feed(dict(state=[fr[i]], action=[ac[i]], reward=[rw[i]], mask=[mk[i]]))
        pre = "p%d_" % c
        rounds = 6
        U, F, NF = [], [], []
        S = {k: [] for k in ("state", "next_state", "action", "reward", "mask", "sampling_prob", "idx")}
        PR, TR, MP, TOT = [], [], [], []
        for rd in range(rounds):
            random.seed(300 + 10 * c + rd)
            u = np.asarray([random.random() for _ in range(B)])          # the stream random.uniform consumes
            random.seed(300 + 10 * c + rd)
            fills = []

            def choice(seq, fills=fills):                                # instrumented stdlib call (replay.py:186)
                i = random.randrange(len(seq))
                fills.append(i)
                return seq[i]

            random.choice = choice
            TOT.append(rp.tree.total())
            smp = rp.sample()
            random.choice = real_choice
            U.append(u), NF.append(len(fills)), F.append(np.asarray(fills + [0] * (B - len(fills)), np.int64))
            for k in S:
                S[k].append(np.asarray(getattr(smp, k)))
            prio = ((np.abs(rng.randn(B)) + 0.01) ** 0.5).astype(np.float32)
            rp.update_priorities(zip(np.asarray(smp.idx, np.float32).astype(np.int64), prio))  # tensor(idx).long() round trip
            # interleave a few feeds so the ring moves between rounds
            base = feeds + rd * 2
            for j in range(2):
                k2 = (base + j) % feeds
                rp.feed(dict(state=[fr[k2]], action=[ac[k2]], reward=[rw[k2]], mask=[mk[k2]]))
            PR.append(prio), TR.append(rp.tree.tree.copy()), MP.append(rp.max_priority)
        out.update({
            pre + "cfg": np.asarray([M, hl, n, feeds, B, rounds], np.int64), pre + "discount": np.float64(disc),
            pre + "frames": fr, pre + "actions": ac, pre + "rewards": rw, pre + "masks": mk,
            pre + "u": np.stack(U), pre + "fills": np.stack(F), pre + "n_fills": np.asarray(NF, np.int64),
            pre + "total": np.asarray(TOT, np.float64), pre + "prio": np.stack(PR), pre + "tree": np.stack(TR),
            pre + "max_priority": np.asarray(MP, np.float64),
        })
        for k in S:
            out[pre + "s_" + k] = np.stack(S[k])
    out["n_cases"] = np.int64(len(cases))
    save("replay_per", **out)


# --------------------------------------------------------------------------------------- loss boundary
class FakeNet:
    """Stands in for a network at the loss boundary: returns the prepared head outputs keyed by
    which input tensor it is called on (states vs next_states)."""

    def __init__(self, table):
        self.table = table

    def __call__(self, x):
        return self.table[int(x[0, 0].item())]


def _bare(cls, config):
    a = object.__new__(cls)
    a.config = config
    return a


def gen_losses():
    out = {}
    B, A = 32, 6
    g = torch.Generator().manual_seed(11)
    states = torch.zeros(B, 1)                    # key 0
    next_states = torch.ones(B, 1)                # key 1
    action = torch.randint(0, A, (B,), generator=g).numpy()
    reward = torch.randint(-1, 2, (B,), generator=g).double().numpy()
    mask = (torch.rand(B, generator=g) > 0.2).int().numpy()
    Tr = ref.Transition

    def cfg(**kw):
        c = ref.Config()
        c.state_normalizer = ref.RescaleNormalizer()
        c.discount, c.n_step, c.batch_size = 0.99, 1, B
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    tr = Tr(state=states, action=action, reward=reward, next_state=next_states, mask=mask)
    out.update(action=action, reward=reward, mask=mask)
    # DQN (plain / double, n_step 1 / 3)
    q, qn_t, qn_o = (torch.randn(B, A, generator=g) for _ in range(3))
    qn_t[3, :] = 0.25                              # exact tie -> argmax / max tie-breaking
    for double in (False, True):
        for n in (1, 3):
            ag = _bare(ref.DQNAgent, cfg(double_q=double, n_step=n))
            ag.network = FakeNet({0: dict(q=q), 1: dict(q=qn_o)})
            ag.target_network = FakeNet({1: dict(q=qn_t)})
            d = ag.compute_loss(tr)
            out["dqn_d%d_n%d_delta" % (double, n)] = d.numpy()
            out["dqn_d%d_n%d_loss" % (double, n)] = ag.reduce_loss(d).numpy()
    out.update(dqn_q=q.numpy(), dqn_qn_t=qn_t.numpy(), dqn_qn_o=qn_o.numpy())
    # C51
    N = 51
    logits, ln_t, ln_o = (torch.randn(B, A, N, generator=g) * 2 for _ in range(3))
    prob, logp = torch.softmax(logits, -1), torch.log_softmax(logits, -1)
    pn_t, pn_o = torch.softmax(ln_t, -1), torch.softmax(ln_o, -1)
    for double in (False, True):
        c = cfg(double_q=double, categorical_v_min=-10, categorical_v_max=10, categorical_n_atoms=N)
        ag = _bare(ref.CategoricalDQNAgent, c)
        ag.network = FakeNet({0: dict(prob=prob, log_prob=logp), 1: dict(prob=pn_o)})
        ag.target_network = FakeNet({1: dict(prob=pn_t)})
        c.atoms = np.linspace(c.categorical_v_min, c.categorical_v_max, c.categorical_n_atoms)
        ag.batch_indices = ref.range_tensor(B)
        ag.atoms = ref.tensor(c.atoms)
        ag.delta_atom = (c.categorical_v_max - c.categorical_v_min) / float(c.categorical_n_atoms - 1)
        kl = ag.compute_loss(tr)
        out["c51_d%d_kl" % double] = kl.numpy()
        out["c51_d%d_loss" % double] = ag.reduce_loss(kl).numpy()
    out.update(c51_logp=logp.numpy(), c51_pn_t=pn_t.numpy(), c51_pn_o=pn_o.numpy(), c51_atoms=ag.atoms.numpy())
    # QR-DQN
    NQ = 200
    quant, qn = torch.randn(B, A, NQ, generator=g), torch.randn(B, A, NQ, generator=g)
    c = cfg(num_quantiles=NQ)
    ag = _bare(ref.QuantileRegressionDQNAgent, c)
    ag.network = FakeNet({0: dict(quantile=quant)})
    ag.target_network = FakeNet({1: dict(quantile=qn)})
    ag.batch_indices = ref.range_tensor(B)
    ag.quantile_weight = 1.0 / NQ
    ag.cumulative_density = ref.tensor((2 * np.arange(NQ) + 1) / (2.0 * NQ)).view(1, -1)
    lv = ag.compute_loss(tr)
    out.update(qr_quant=quant.numpy(), qr_qn=qn.numpy(), qr_vec=lv.numpy(), qr_loss=ag.reduce_loss(lv).numpy())
    save("losses", **out)


# --------------------------------------------------------------------------------------- DQN-family agent steps
def _sd_np(net):
    return {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}


def gen_agent_steps():
    out = {}
    specs = [
        ("dqn_per", ref.DQNAgent, dict(double_q=True, replay_cls=ref.PrioritizedReplay)),
        ("dqn_uni", ref.DQNAgent, dict(double_q=False, replay_cls=ref.UniformReplay)),
        ("c51", ref.CategoricalDQNAgent, dict(double_q=False, replay_cls=ref.UniformReplay)),
        ("qr", ref.QuantileRegressionDQNAgent, dict(double_q=False, replay_cls=ref.UniformReplay)),
    ]
    for name, cls, kw in specs:
        np.random.seed(5), torch.manual_seed(5), random.seed(5)
        c = ref.Config()
        c.merge(dict(tag=None, n_step=1))
        c.task_fn = lambda: SynthTask("CartPole-v0", seed=3)
        c.eval_env = c.task_fn()
        c.history_length, c.batch_size, c.discount = 1, 16, 0.99
        c.double_q = kw["double_q"]
        if name.startswith("dqn"):
            c.optimizer_fn = lambda p: torch.optim.RMSprop(p, lr=0.00025, alpha=0.95, eps=0.01, centered=True)
            if name == "dqn_per":
                c.network_fn = lambda: ref.DuelingNet(c.action_dim, ref.FCBody(c.state_dim, hidden_units=(32, 32)))
            else:
                c.network_fn = lambda: ref.VanillaNet(c.action_dim, ref.FCBody(c.state_dim, hidden_units=(32, 32)))
        elif name == "c51":
            c.optimizer_fn = lambda p: torch.optim.Adam(p, lr=0.00025, eps=0.01 / 32)
            c.categorical_v_min, c.categorical_v_max, c.categorical_n_atoms = -100, 100, 50   # examples.py:189-191
            c.network_fn = lambda: ref.CategoricalNet(c.action_dim, c.categorical_n_atoms,
                                                      ref.FCBody(c.state_dim, hidden_units=(32, 32)))
        else:
            c.optimizer_fn = lambda p: torch.optim.Adam(p, lr=0.00005, eps=0.01 / 32)
            c.num_quantiles = 20
            c.network_fn = lambda: ref.QuantileNet(c.action_dim, c.num_quantiles,
                                                   ref.FCBody(c.state_dim, hidden_units=(32, 32)))
        rk = dict(memory_size=256, batch_size=c.batch_size, n_step=1, discount=c.discount, history_length=1)
        c.replay_fn = lambda: ref.ReplayWrapper(kw["replay_cls"], rk, False)
        c.replay_eps, c.replay_alpha = 0.01, 0.5
        c.replay_beta = ref.LinearSchedule(0.4, 1.0, 200)
        c.random_action_prob = ref.LinearSchedule(1.0, 0.1, 100)
        c.target_network_update_freq, c.exploration_steps = 5, 40
        c.sgd_update_frequency, c.gradient_clip, c.async_actor = 4, 5, False
        ag = cls(c)
        ag.logger = NullLogger()
        init_sd = _sd_np(ag.network)
        rec = dict(batches=[], deltas=[])
        orig_sample, orig_loss = ag.replay.sample, ag.compute_loss

        def sample(rec=rec, f=orig_sample):
            t = f()
            rec["batches"].append(t)
            return t

        def compute_loss(t, rec=rec, f=orig_loss):
            d = f(t)
            rec["deltas"].append(d.detach().numpy().copy())
            return d

        ag.replay.sample, ag.compute_loss = sample, compute_loss
        params_after, target_after = [], []
        for it in range(30):
            ag.step()
            if ag.total_steps > c.exploration_steps:
                params_after.append(np.concatenate([p.detach().numpy().ravel() for p in ag.network.parameters()]))
                target_after.append(np.concatenate([p.detach().numpy().ravel() for p in ag.target_network.parameters()]))
        nb = len(rec["batches"])
        assert nb == len(params_after) and nb > 10
        pre = name + "_"
        for k, v in init_sd.items():
            out[pre + "init." + k] = v
        out[pre + "keys"] = np.asarray(list(init_sd.keys()))
        for f in rec["batches"][0]._fields:
            out[pre + "b_" + f] = np.stack([np.asarray(getattr(b, f)) for b in rec["batches"]])
        out[pre + "delta"] = np.stack(rec["deltas"])
        out[pre + "params"] = np.stack(params_after)
        out[pre + "target"] = np.stack(target_after)
        if name == "dqn_per":
            out[pre + "tree"] = ag.replay.replay.tree.tree.copy()
            out[pre + "max_priority"] = np.float64(ag.replay.replay.max_priority)
        # (reference bug: ReplayWrapper.close() needs .pipe, absent in sync mode -- replay.py:276-278)
    save("agent_steps", **out)


# --------------------------------------------------------------------------------------- A2C / PPO / GAE
def gen_onpolicy():
    out = {}
    captured = {}
    real_extract = ref.Storage.extract

    def extract(self, keys):
        T = self.memory_size
        captured["reward"] = torch.stack(self.reward[:T]).detach().numpy().copy()
        captured["mask"] = torch.stack(self.mask[:T]).detach().numpy().copy()
        captured["v"] = torch.stack(self.v[:T + 1]).detach().numpy().copy()
        captured["advantage"] = torch.stack(self.advantage[:T]).detach().numpy().copy()
        captured["ret"] = torch.stack(self.ret[:T]).detach().numpy().copy()
        for k in ("state", "action", "log_pi_a"):
            if len(getattr(self, k)) and getattr(self, k)[0] is not None:
                captured[k] = torch.stack(list(getattr(self, k)[:T])).detach().numpy().copy()
        return real_extract(self, keys)

    ref.Storage.extract = extract
    real_perm = np.random.permutation

    # ---- A2C, CartPole, (T, N) = (5, 8) -- BASELINE configs[0]
    np.random.seed(9), torch.manual_seed(9)
    c = ref.Config()
    c.merge(dict(tag=None))
    c.num_workers = 8
    c.task_fn = lambda: SynthTask("CartPole-v0", num_envs=c.num_workers, seed=4)
    c.eval_env = SynthTask("CartPole-v0", seed=4)
    c.optimizer_fn = lambda p: torch.optim.RMSprop(p, 0.001)
    c.network_fn = lambda: ref.CategoricalActorCriticNet(c.state_dim, c.action_dim, ref.FCBody(c.state_dim, gate=torch.tanh))
    c.discount, c.use_gae, c.gae_tau, c.entropy_weight, c.rollout_length, c.gradient_clip = 0.99, True, 0.95, 0.01, 5, 0.5
    ag = ref.A2CAgent(c)
    ag.logger = NullLogger()
    states_seen = []
    real_task_step = ag.task.step
    out["a2c_state0"] = np.asarray(ag.states, np.float32)

    def task_step(a):
        r = real_task_step(a)
        states_seen.append((np.asarray(a).copy(), np.asarray(r[0], np.float32), np.asarray(r[1], np.float64), np.asarray(r[2])))
        return r

    ag.task.step = task_step
    init = _sd_np(ag.network)
    P = []
    G = {k: [] for k in ("reward", "mask", "v", "advantage", "ret")}
    for it in range(6):
        ag.step()
        P.append(np.concatenate([p.detach().numpy().ravel() for p in ag.network.parameters()]))
        for k in G:
            G[k].append(captured[k])
    for k, v in init.items():
        out["a2c_init." + k] = v
    out["a2c_keys"] = np.asarray(list(init.keys()))
    out["a2c_actions"] = np.stack([s[0] for s in states_seen])
    out["a2c_next_states"] = np.stack([s[1] for s in states_seen])
    out["a2c_rewards"] = np.stack([s[2] for s in states_seen])
    out["a2c_dones"] = np.stack([s[3] for s in states_seen])
    out["a2c_params"] = np.stack(P)
    for k in G:
        out["a2c_gae_" + k] = np.stack(G[k])

    # ---- PPO continuous (non-shared): small (T, N) = (16, 4), 3 epochs x mb 8 (ragged tail: 64 % 24 -> mb 24)
    for tag, (T, N, mb, epochs) in dict(small=(16, 4, 24, 3)).items():
        np.random.seed(10), torch.manual_seed(10)
        c = ref.Config()
        c.merge(dict(tag=None))
        c.num_workers = N
        c.task_fn = lambda: SynthTask("SyntheticCheetah-v0", num_envs=N, seed=6)
        c.eval_env = SynthTask("SyntheticCheetah-v0", seed=6)
        c.network_fn = lambda: ref.GaussianActorCriticNet(
            c.state_dim, c.action_dim, actor_body=ref.FCBody(c.state_dim, gate=torch.tanh),
            critic_body=ref.FCBody(c.state_dim, gate=torch.tanh))
        c.actor_opt_fn = lambda p: torch.optim.Adam(p, 3e-4)
        c.critic_opt_fn = lambda p: torch.optim.Adam(p, 1e-3)
        c.discount, c.use_gae, c.gae_tau, c.gradient_clip = 0.99, True, 0.95, 0.5
        c.rollout_length, c.optimization_epochs, c.mini_batch_size = T, epochs, mb
        c.ppo_ratio_clip, c.target_kl = 0.2, 0.01
        c.state_normalizer = ref.MeanStdNormalizer()
        ag = ref.PPOAgent(c)
        ag.logger = NullLogger()
        init = _sd_np(ag.network)
        its = 3
        perms_all, P, C = [], [], {k: [] for k in ("reward", "mask", "v", "advantage", "ret", "state", "action", "log_pi_a")}
        for it in range(its):
            perms = []

            def perm(x, perms=perms):
                p = real_perm(x)
                perms.append(np.asarray(p).copy())
                return p

            np.random.permutation = perm
            ag.step()
            np.random.permutation = real_perm
            perms_all.append(np.stack(perms))
            P.append(np.concatenate([p.detach().numpy().ravel() for p in ag.network.parameters()]))
            for k in C:
                C[k].append(captured[k])
        pre = "ppo_%s_" % tag
        for k, v in init.items():
            out[pre + "init." + k] = v
        out[pre + "keys"] = np.asarray(list(init.keys()))
        out[pre + "param_names"] = np.asarray([n for n, _ in ag.network.named_parameters()])
        out[pre + "cfg"] = np.asarray([T, N, mb, epochs, its], np.int64)
        out[pre + "perms"] = np.stack(perms_all)
        out[pre + "params"] = np.stack(P)
        for k in C:
            out[pre + k] = np.stack(C[k])
        out[pre + "rms_mean"], out[pre + "rms_var"] = c.state_normalizer.rms.mean, c.state_normalizer.rms.var
    ref.Storage.extract = real_extract

    # ---- GAE alone at the BASELINE shapes, through the reference loop re-run on stored tensors
    # (the loop lives inline in PPO_agent.py:51-61; the capture above already pins it at (5,8) and (16,4);
    #  here the same reference code path is driven at (128,8) and (2048,16) via a PPOAgent whose rollout is synthetic)
    for (T, N) in ((128, 8), (2048, 16)):
        g = torch.Generator().manual_seed(T)
        reward = torch.randn(T, N, 1, generator=g)
        mask = (torch.rand(T, N, 1, generator=g) > 0.02).float()
        v = torch.randn(T + 1, N, 1, generator=g)
        # reference arithmetic, statement by statement (PPO_agent.py:51-61) on the reference's Storage object
        st = ref.Storage(T)
        st.reward, st.mask, st.v = list(reward), list(mask), list(v)
        st.placeholder()
        cfgx = ref.Config()
        cfgx.discount, cfgx.gae_tau, cfgx.use_gae, cfgx.num_workers = 0.99, 0.95, True, N
        advantages = ref.tensor(np.zeros((N, 1)))
        returns = v[T].detach()
        for i in reversed(range(T)):
            returns = st.reward[i] + cfgx.discount * st.mask[i] * returns
            td_error = st.reward[i] + cfgx.discount * st.mask[i] * st.v[i + 1] - st.v[i]
            advantages = advantages * cfgx.gae_tau * cfgx.discount * st.mask[i] + td_error
            st.advantage[i] = advantages.detach()
            st.ret[i] = returns.detach()
        out["gae_%d_%d_reward" % (T, N)] = reward.numpy()
        out["gae_%d_%d_mask" % (T, N)] = mask.numpy()
        out["gae_%d_%d_v" % (T, N)] = v.numpy()
        out["gae_%d_%d_adv" % (T, N)] = torch.stack(st.advantage).numpy()
        out["gae_%d_%d_ret" % (T, N)] = torch.stack(st.ret).numpy()
    save("onpolicy", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["sumtree", "uniform", "per", "losses", "agent_steps", "onpolicy"]
    for w in which:
        globals()["gen_" + w]()
