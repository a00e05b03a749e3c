"""Synchronous advantage actor-critic behind the reference's interface (``deep_rl/agent/A2C_agent.py:13-64``).

The rollout keeps the autograd graph of its T forward passes exactly like the reference (``storage.feed(prediction)``
stores tensors with grad, A2C_agent.py:31) and does ONE backward.  On a CUDA device the GAE recurrence and the
objective (+ its gradient with respect to log-prob / entropy / value) are the kernels of ``csrc/onpolicy.cu``; on
``select_device(-1)`` -- BASELINE configs[0], "a2c_feature CartPole, 8 workers, CPU only, plumbing" -- the same
statements run as torch expressions, which is the reference's own path, not a fallback for a missing library.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..component import Storage
from ..utils import tensor, to_np
from .BaseAgent import BaseAgent


def gae_torch(reward, mask, value, discount, tau, use_gae):
    """A2C_agent.py:43-53 == PPO_agent.py:51-61 (torch statements; CPU device only)."""
    T = len(reward)
    adv = torch.zeros_like(value[0])
    ret = value[T].detach()
    advs, rets = [None] * T, [None] * T
    for i in reversed(range(T)):
        ret = reward[i] + discount * mask[i] * ret
        if not use_gae:
            adv = ret - value[i].detach()
        else:
            td = reward[i] + discount * mask[i] * value[i + 1].detach() - value[i].detach()
            adv = adv * tau * discount * mask[i] + td
        advs[i], rets[i] = adv.detach(), ret.detach()
    return advs, rets


def compute_advantages(storage, config, last_v, exact=True):
    """Fill ``storage.advantage`` / ``storage.ret`` (lists of (N,1) tensors) from reward / mask / v."""
    T = config.rollout_length
    v = [x.detach() for x in storage.v[:T]] + [last_v.detach()]
    if v[0].is_cuda:
        adv, ret = ops.gae(torch.stack(storage.reward[:T]), torch.stack(storage.mask[:T]), torch.stack(v),
                           config.discount, config.gae_tau, config.use_gae, exact=exact)
        storage.advantage, storage.ret = list(adv.unbind(0)), list(ret.unbind(0))
    else:
        storage.advantage, storage.ret = gae_torch(storage.reward[:T], storage.mask[:T], v, config.discount,
                                                   config.gae_tau, config.use_gae)


class A2CAgent(BaseAgent):
    def __init__(self, config):
        BaseAgent.__init__(self, config)
        self.config = config
        self.task = config.task_fn()
        self.network = config.network_fn()
        self.optimizer = config.optimizer_fn(self.network.parameters())
        self.total_steps = 0
        self.states = self.task.reset()
        self.last_loss = None

    def eval_step(self, state):
        with torch.no_grad():
            prediction = self.network(self.config.state_normalizer(np.asarray([np.asarray(s) for s in state])))
        return to_np(prediction["action"])

    def step(self):
        config = self.config
        storage = Storage(config.rollout_length)
        states = self.states
        for _ in range(config.rollout_length):
            prediction = self.network(config.state_normalizer(np.asarray([np.asarray(s) for s in states])))
            next_states, rewards, terminals, info = self.task.step(to_np(prediction["action"]))
            self.record_online_return(info)
            rewards = config.reward_normalizer(rewards)
            storage.feed(prediction)
            storage.feed({"reward": tensor(rewards).unsqueeze(-1), "mask": tensor(1 - terminals).unsqueeze(-1)})
            states = next_states
            self.total_steps += config.num_workers

        self.states = states
        prediction = self.network(config.state_normalizer(np.asarray([np.asarray(s) for s in states])))
        storage.feed(prediction)
        storage.placeholder()
        compute_advantages(storage, config, prediction["v"])

        entries = storage.extract(["log_pi_a", "v", "ret", "advantage", "entropy"])
        self.optimizer.zero_grad()
        if entries.v.is_cuda:
            r = ops.a2c_loss_fused(entries.log_pi_a.detach(), entries.entropy.detach(), entries.v.detach(),
                                   entries.advantage, entries.ret, config.entropy_weight, config.value_loss_weight)
            shape = entries.v.shape
            torch.autograd.backward([entries.log_pi_a, entries.entropy, entries.v],
                                    [r["dlogp"].view(shape), r["dent"].view(shape), r["dv"].view(shape)])
            self.last_loss = r["out"][0]
        else:
            policy_loss = -(entries.log_pi_a * entries.advantage).mean()
            value_loss = 0.5 * (entries.ret - entries.v).pow(2).mean()
            entropy_loss = entries.entropy.mean()
            loss = policy_loss - config.entropy_weight * entropy_loss + config.value_loss_weight * value_loss
            loss.backward()
            self.last_loss = loss.detach()
        nn.utils.clip_grad_norm_(self.network.parameters(), config.gradient_clip)
        self.optimizer.step()
