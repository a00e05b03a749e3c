"""n-step Q-learning behind the reference's interface (``deep_rl/agent/NStepDQN_agent.py:13-70``).

``step()``: a ``rollout_length`` rollout with epsilon-greedy actions whose q-values keep their autograd graphs
(NStepDQN_agent.py:33-46), hard target sync on the reference's schedule inside the rollout (:48-49), bootstrap
``max_a target(s_T)`` (:56-57), backward return scan ``ret_t = r_t + discount * mask_t * ret_{t+1}`` (:58-60), loss
``0.5 * mean((q[a] - ret)^2)`` (:63), clip, optimizer step (:64-67).

On a CUDA device the return scan is the K7 scan kernel (``ops.gae``: its ``ret`` output is this recurrence; csrc/onpolicy.cu) and
the loss and its gradient are the fused DQN kernel (``ops.dqn_loss_fused`` with the n-step return as the target: reward = ret, mask = 0;
csrc/losses.cu).  On ``select_device(-1)`` the same statements run as torch expressions -- the reference's own CPU path.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..component import Storage
from ..utils import epsilon_greedy, tensor, to_np
from .BaseAgent import BaseAgent


class NStepDQNAgent(BaseAgent):
    def __init__(self, config):
        BaseAgent.__init__(self, config)
        self.config = config
        self.task = config.task_fn()
        self.network = config.network_fn()
        self.target_network = config.network_fn()
        self.optimizer = config.optimizer_fn(self.network.parameters())
        self.target_network.load_state_dict(self.network.state_dict())
        self.total_steps = 0
        self.states = self.task.reset()
        self.last_loss = None

    def eval_step(self, state):
        """(the reference defines none for this agent; greedy action, as DQNAgent.eval_step DQN_agent.py:69-75)"""
        with torch.no_grad():
            q = self.network(self.config.state_normalizer(np.asarray([np.asarray(s) for s in state])))["q"]
        return to_np(q.argmax(dim=-1))

    def _obs(self, states):
        return self.config.state_normalizer(np.asarray([np.asarray(s) for s in states]))

    def step(self):
        config = self.config
        T = config.rollout_length
        storage = Storage(T)
        states = self.states
        for _ in range(T):
            q = self.network(self._obs(states))["q"]
            epsilon = config.random_action_prob(config.num_workers)
            actions = epsilon_greedy(epsilon, to_np(q))
            next_states, rewards, terminals, info = self.task.step(actions)
            self.record_online_return(info)
            rewards = config.reward_normalizer(rewards)
            storage.feed({"q": q, "action": tensor(actions).unsqueeze(-1).long(), "reward": tensor(rewards).unsqueeze(-1),
                          "mask": tensor(1 - np.asarray(terminals)).unsqueeze(-1)})
            states = next_states
            self.total_steps += config.num_workers
            if self.total_steps // config.num_workers % config.target_network_update_freq == 0:
                self.target_network.load_state_dict(self.network.state_dict())
        self.states = states
        storage.placeholder()

        with torch.no_grad():
            boot = self.target_network(self._obs(states))["q"].max(dim=1, keepdim=True)[0]
        self.optimizer.zero_grad()
        if boot.is_cuda:
            reward, mask = torch.stack(storage.reward[:T]), torch.stack(storage.mask[:T])
            value = torch.zeros((T + 1,) + tuple(boot.shape), device=boot.device, dtype=torch.float32)
            value[T] = boot
            _, ret = ops.gae(reward, mask, value, config.discount, 1.0, use_gae=False)
            q = torch.cat(storage.q[:T], dim=0)
            rows = q.shape[0]
            # one launch: delta, loss = 0.5 * mean(delta^2) and dloss/dq, with y = ret + 1.0 * q_next * 0
            r = ops.dqn_loss_fused(q.detach(), torch.zeros_like(q), None, torch.cat(storage.action[:T], dim=0).view(-1),
                                   ret.reshape(rows), torch.zeros(rows, device=q.device), 1.0)
            torch.autograd.backward([q], [r["dq"]])
            loss = r["loss"][0]
        else:
            ret = boot
            for i in reversed(range(T)):
                ret = storage.reward[i] + config.discount * storage.mask[i] * ret
                storage.ret[i] = ret
            entries = storage.extract(["q", "action", "ret"])
            loss = 0.5 * (entries.q.gather(1, entries.action) - entries.ret).pow(2).mean()
            loss.backward()
        self.last_loss = loss.detach()
        nn.utils.clip_grad_norm_(self.network.parameters(), config.gradient_clip)
        self.optimizer.step()
