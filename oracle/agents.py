"""TEST INFRASTRUCTURE -- torch-CPU restatements of the reference agents' ``step`` bodies, used
(i) to pin the oracle end-to-end against goldens produced by the real reference and (ii) as the
timed CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``; kind = "port").

Networks are functional (oracle/nets.py) over a dict of leaf tensors named like the reference's
``state_dict`` so parameters can be exchanged with the product and with the real reference.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import losses, nets
from .replay import PrioritizedTransition


def f32(x):                                                             # utils/torch_utils.py:20-25
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(np.asarray(x, dtype=np.float32))


def leafify(sd):
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}


def clip_grad_norm(params, max_norm):                                   # torch.nn.utils.clip_grad_norm_ semantics
    torch.nn.utils.clip_grad_norm_(params, max_norm)


class DQNFamilyOracle:
    """DQN_agent.py:48-138, CategoricalDQN_agent.py:27-89, QuantileRegressionDQN_agent.py:23-77
    with a synchronous actor/replay.  ``head`` in {"vanilla","dueling","categorical","quantile"};
    ``body`` in {"nature","fc"}."""

    def __init__(self, sd, head, body, action_dim, optimizer_fn, discount, n_step=1, double_q=False,
                 gradient_clip=5.0, state_coef=1.0, atoms=None, v_min=None, v_max=None, num_quantiles=None,
                 replay_eps=0.01, replay_alpha=0.5, replay_beta=None, gate=F.relu):
        self.sd = leafify(sd)
        self.target_sd = {k: v.detach().clone() for k, v in sd.items()}
        self.head, self.body, self.A = head, body, action_dim
        self.params = list(self.sd.values())
        self.opt = optimizer_fn(self.params)
        self.discount_n = discount ** n_step
        self.double_q, self.clip, self.coef = double_q, gradient_clip, state_coef
        self.atoms = None if atoms is None else f32(atoms)
        self.v_min, self.v_max, self.N = v_min, v_max, num_quantiles
        self.eps, self.alpha, self.beta = replay_eps, replay_alpha, replay_beta
        self.gate = gate

    def forward(self, sd, x):
        phi = nets.nature_body(sd, x) if self.body == "nature" else nets.fc_body(sd, x, "body.", self.gate)
        if self.head == "vanilla":
            return dict(q=nets.vanilla_q(sd, phi))
        if self.head == "dueling":
            return dict(q=nets.dueling_q(sd, phi))
        if self.head == "categorical":
            p, lp = nets.categorical(sd, phi, self.A, self.atoms.numel())
            return dict(prob=p, log_prob=lp)
        return dict(quantile=nets.quantile(sd, phi, self.A, self.N))

    def normalize(self, x):                                             # normalizer.py:58-61 (float64 product)
        if not isinstance(x, torch.Tensor):
            x = np.asarray(x)
        return f32(self.coef * x)

    def compute_loss(self, tr):
        s, s2 = self.normalize(tr.state), self.normalize(tr.next_state)
        r, m, a = f32(tr.reward), f32(tr.mask), f32(tr.action).long()
        if self.head in ("vanilla", "dueling"):
            with torch.no_grad():
                qn_t = self.forward(self.target_sd, s2)["q"]
                qn_o = self.forward(self.sd, s2)["q"] if self.double_q else None
            q = self.forward(self.sd, s)["q"]
            return losses.dqn_delta(q, qn_t, qn_o, a, r, m, self.discount_n)
        if self.head == "categorical":
            with torch.no_grad():
                pn_t = self.forward(self.target_sd, s2)["prob"]
                pn_o = self.forward(self.sd, s2)["prob"] if self.double_q else None
            lp = self.forward(self.sd, s)["log_prob"]
            return losses.c51_kl(lp, pn_t, pn_o, a, r, m, self.atoms, self.v_min, self.v_max, self.discount_n)
        qn = self.forward(self.target_sd, s2)["quantile"].detach()
        q = self.forward(self.sd, s)["quantile"]
        return losses.qr_loss(q, qn, a, r, m, self.discount_n)

    def reduce_loss(self, loss):
        return losses.dqn_reduce(loss) if self.head in ("vanilla", "dueling") else loss.mean()

    def update(self, tr, replay=None):
        """DQN_agent.py:115-134 for one sampled batch; returns the scalar loss."""
        loss = self.compute_loss(tr)
        if isinstance(tr, PrioritizedTransition) or hasattr(tr, "sampling_prob"):
            prio, _, loss = losses.per_block(loss, f32(tr.sampling_prob), self.beta(), self.eps, self.alpha)
            if replay is not None:
                replay.update_priorities(zip(f32(tr.idx).long().numpy(), prio.detach().numpy()))
        loss = self.reduce_loss(loss)
        self.opt.zero_grad()
        loss.backward()
        clip_grad_norm(self.params, self.clip)
        self.opt.step()
        return loss.detach()

    def sync_target(self):                                              # DQN_agent.py:136-138
        for k, v in self.sd.items():
            self.target_sd[k].copy_(v.detach())


def random_sample(indices, batch_size):                                 # utils/misc.py:55-62
    indices = np.asarray(np.random.permutation(indices))
    full = len(indices) // batch_size * batch_size
    for row in indices[:full].reshape(-1, batch_size):
        yield row
    if len(indices) % batch_size:
        yield indices[full:]


def ppo_update(sd, actor_keys, critic_keys, actor_opt, critic_opt, states, actions, log_pi_old, ret, adv,
               epochs, mini_batch, clip, entropy_weight, target_kl, gate=torch.tanh):
    """PPO_agent.py:63-99, non-shared-representation branch (examples.py:496-522)."""
    adv = losses.normalize_advantage(adv)
    for _ in range(epochs):
        for idx in random_sample(np.arange(states.size(0)), mini_batch):
            idx = torch.from_numpy(np.asarray(idx, dtype=np.float32)).long()   # tensor(batch_indices).long()
            out = nets.gaussian_actor_critic(sd, states[idx], actions[idx], gate)
            pl, vl, kl = losses.ppo_losses(out["log_pi_a"], out["entropy"], out["v"], log_pi_old[idx],
                                           adv[idx], ret[idx], clip, entropy_weight)
            if kl <= 1.5 * target_kl:
                actor_opt.zero_grad()
                pl.backward()
                actor_opt.step()
            critic_opt.zero_grad()
            vl.backward()
            critic_opt.step()
    return adv


def a2c_update(sd, params, opt, states, actions, rewards, masks, discount, tau, entropy_weight,
               value_loss_weight, gradient_clip, gate=torch.tanh):
    """A2C_agent.py:22-64 for one rollout whose env interaction is given: ``states`` (T+1,N,obs),
    ``actions`` (T,N), ``rewards``/``masks`` (T,N,1).  Forward passes keep their graphs, GAE runs on
    detached values, one backward + clip + optimizer step.  Returns (adv, ret) for inspection."""
    T = actions.shape[0]
    preds = [nets.categorical_actor_critic(sd, states[t], actions[t], gate) for t in range(T)]
    last = nets.categorical_actor_critic(sd, states[T], actions[T - 1], gate)
    v = torch.stack([p["v"] for p in preds] + [last["v"]])
    adv, ret = losses.gae(rewards, masks, v.detach(), discount, tau)
    cat = lambda k: torch.cat([p[k] for p in preds], dim=0)
    loss = losses.a2c_loss(cat("log_pi_a"), cat("v"), ret.reshape(-1, 1), adv.reshape(-1, 1), cat("entropy"),
                           entropy_weight, value_loss_weight)
    opt.zero_grad()
    loss.backward()
    clip_grad_norm(params, gradient_clip)
    opt.step()
    return adv, ret


def nstep_dqn_update(sd, target_sd, params, opt, states, actions, rewards, masks, discount, gradient_clip, gate=F.relu):
    """NStepDQN_agent.py:26-70 for one rollout whose env interaction is given: ``states`` (T+1,N,obs), ``actions`` (T,N),
    ``rewards``/``masks`` (T,N,1).  The T forward passes keep their graphs (``storage.feed({'q': q})`` :40); the bootstrap is
    ``max_a target(s_T)`` (:56-57); ``ret_t = r_t + discount * mask_t * ret_{t+1}`` (:58-60);
    ``loss = 0.5 * mean((q[a] - ret)^2)`` (:63); clip + optimizer step (:64-67).  The target sync inside the rollout (:48-49)
    is the caller's.  Returns (ret (T,N,1), loss)."""
    T = actions.shape[0]
    q = [nets.vanilla_q(sd, nets.fc_body(sd, states[t], "body.", gate)) for t in range(T)]
    with torch.no_grad():
        ret = nets.vanilla_q(target_sd, nets.fc_body(target_sd, states[T], "body.", gate)).max(dim=1, keepdim=True)[0]
    rets = [None] * T
    for i in reversed(range(T)):
        ret = rewards[i] + discount * masks[i] * ret
        rets[i] = ret
    qa = torch.cat(q, dim=0).gather(1, actions.reshape(-1, 1).long())
    loss = 0.5 * (qa - torch.cat(rets, dim=0)).pow(2).mean()
    opt.zero_grad()
    loss.backward()
    clip_grad_norm(params, gradient_clip)
    opt.step()
    return torch.stack(rets), loss.detach()
