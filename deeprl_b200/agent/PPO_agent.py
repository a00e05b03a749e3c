"""PPO behind the reference's interface (``deep_rl/agent/PPO_agent.py:13-99``): rollout, GAE, advantage
normalisation, ``optimization_epochs`` x minibatches of the clipped surrogate, KL-gated actor step (non-shared
representation) or one clipped optimizer step (shared representation).

CUDA device: GAE = ``b2rl_gae``, advantage normalisation = ``b2rl_normalize_advantage``, each minibatch's surrogate /
value loss / approx-KL and their gradients = one ``b2rl_ppo_loss`` launch.  ``select_device(-1)``: torch statements
(the reference's own CPU path; see A2C_agent.py docstring).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..component import Storage
from ..utils import random_sample, tensor, to_np
from .A2C_agent import compute_advantages
from .BaseAgent import BaseAgent


class PPOAgent(BaseAgent):
    def __init__(self, config):
        BaseAgent.__init__(self, config)
        self.config = config
        self.task = config.task_fn()
        self.network = config.network_fn()
        if config.shared_repr:
            self.opt = config.optimizer_fn(self.network.parameters())
        else:
            self.actor_opt = config.actor_opt_fn(self.network.actor_params)
            self.critic_opt = config.critic_opt_fn(self.network.critic_params)
        self.total_steps = 0
        self.states = self.task.reset()
        self._raw_states, self._raw_seen = self.states, True      # (device actor: un-normalised copy, already counted below)
        self.states = config.state_normalizer(self.states)
        if config.shared_repr:
            self.lr_scheduler = torch.optim.lr_scheduler.LambdaLR(self.opt, lambda step: 1 - step / config.max_steps)
        self.gae_exact = True
        self.last_stats = None

    def eval_step(self, state):
        with torch.no_grad():
            prediction = self.network(self.config.state_normalizer(state))
        return to_np(prediction["action"])

    # ------------------------------------------------------------------ device-side actor (component/actor.py)
    def _device_actor(self):
        """One launch per env step for normaliser + network forward + sampling when the network / normaliser are the reference's
        continuous-control configuration (examples.py:496-522) on a CUDA device (``config.device_actor = False`` turns it off)."""
        if getattr(self, "_actor", None) is None:
            from ..component import actor as dev_actor
            cfg = self.config
            ok = (getattr(cfg, "device_actor", True) and dev_actor.supported(self.network, cfg.state_normalizer)
                  and cfg.num_workers <= 64 and type(cfg.reward_normalizer).__name__ == "RescaleNormalizer")
            self._actor = dev_actor.DeviceGaussianActor(self.network, cfg.state_normalizer, cfg.num_workers) if ok else False
        return self._actor or None

    def _rollout_device(self, actor):
        """``_rollout`` with the per-step normaliser / forward / sampling on the device: the raw observations go up through a
        pinned double buffer, the actions come back for the host envs; everything else stays in HBM.  As in the reference, every
        observation batch updates the running moments exactly once -- when it is first seen (the constructor counted the reset
        batch on the host; the batch that ends a rollout is counted by its value forward and re-used, read-only, by the first
        step of the next rollout)."""
        config = self.config
        T = config.rollout_length
        storage = Storage(T)
        raw, seen = self._raw_states, self._raw_seen
        keys = ("action", "log_pi_a", "entropy", "mean", "v")
        if not getattr(config, "device_actor_arena", True):        # one set of output tensors and two small uploads per env step
            for _ in range(T):
                pred = actor.step(raw, update=not seen)
                raw, rewards, terminals, info = self.task.step(to_np(pred["action"]))
                seen = False
                self.record_online_return(info)
                rewards = config.reward_normalizer(rewards)
                storage.feed({k: pred[k] for k in keys})
                storage.feed({"reward": tensor(rewards).unsqueeze(-1), "mask": tensor(1 - terminals).unsqueeze(-1),
                              "state": pred["state"]})
                self.total_steps += config.num_workers
            pred = actor.step(raw, update=True)             # value of the last state (PPO_agent.py:46-47)
            self._raw_states, self._raw_seen = raw, True
            storage.feed({k: pred[k] for k in keys})
            last_v = pred["v"]
        else:
            # the actor's outputs go straight into rollout-sized arenas; rewards / masks stay on the host until the rollout ends
            N = config.num_workers
            actor.begin_rollout(T)
            rew, msk = np.empty((T, N), dtype=np.float32), np.empty((T, N), dtype=np.float32)
            for t in range(T):
                action = actor.step_into(t, raw, update=not seen)
                raw, rewards, terminals, info = self.task.step(action)
                seen = False
                self.record_online_return(info)
                rew[t] = np.asarray(config.reward_normalizer(rewards), dtype=np.float32)     # tensor(): float32 (torch_utils.py:20-25)
                msk[t] = np.asarray(1 - terminals, dtype=np.float32)
                self.total_steps += N
            actor.step_into(T, raw, update=True)            # value of the last state (PPO_agent.py:46-47)
            self._raw_states, self._raw_seen = raw, True
            roll, dev = actor.roll, actor.dev
            for k in keys:
                setattr(storage, k, list(roll[k].unbind(0)))                                  # T + 1 entries, like feed()
            storage.state = list(roll["state"][:T].unbind(0))
            storage.reward = list(torch.from_numpy(rew).to(dev).unsqueeze(-1).unbind(0))
            storage.mask = list(torch.from_numpy(msk).to(dev).unsqueeze(-1).unbind(0))
            last_v = roll["v"][T]
        storage.placeholder()
        compute_advantages(storage, config, last_v, exact=self.gae_exact)
        actor.pull_stats()                                  # the host normaliser object stays current (eval_step, save)
        entries = storage.extract(["state", "action", "log_pi_a", "ret", "advantage"])
        return type(entries)(*[x.detach().contiguous() for x in entries])

    def _rollout(self):
        config = self.config
        actor = self._device_actor()
        if actor is not None:
            return self._rollout_device(actor)
        storage = Storage(config.rollout_length)
        states = self.states
        for _ in range(config.rollout_length):
            with torch.no_grad():
                prediction = self.network(states)
            next_states, rewards, terminals, info = self.task.step(to_np(prediction["action"]))
            self.record_online_return(info)
            rewards = config.reward_normalizer(rewards)
            next_states = config.state_normalizer(next_states)
            storage.feed(prediction)
            storage.feed({"reward": tensor(rewards).unsqueeze(-1), "mask": tensor(1 - terminals).unsqueeze(-1),
                          "state": tensor(states)})
            states = next_states
            self.total_steps += config.num_workers
        self.states = states
        with torch.no_grad():
            prediction = self.network(states)
        storage.feed(prediction)
        storage.placeholder()
        compute_advantages(storage, config, prediction["v"], exact=self.gae_exact)
        entries = storage.extract(["state", "action", "log_pi_a", "ret", "advantage"])
        return type(entries)(*[x.detach().contiguous() for x in entries])

    def _normalize(self, entries):
        adv = entries.advantage
        if adv.is_cuda:
            ops.normalize_advantage_(adv)
        else:
            adv.copy_((adv - adv.mean()) / adv.std())

    def _minibatch(self, entries, batch_indices):
        config = self.config
        batch_indices = tensor(batch_indices).long()
        entry = type(entries)(*[x[batch_indices] for x in entries])
        prediction = self.network(entry.state, entry.action)
        if entry.state.is_cuda:
            r = ops.ppo_loss_fused(prediction["log_pi_a"].detach(), prediction["entropy"].detach(), prediction["v"].detach(),
                                   entry.log_pi_a, entry.advantage, entry.ret, config.ppo_ratio_clip, config.entropy_weight)
            shape = prediction["v"].shape
            g_lp, g_en, g_v = r["dlogp"].view(shape), r["dent"].view(shape), r["dv"].view(shape)
            self.last_stats = r["out"]
            if config.shared_repr:
                self.opt.zero_grad()
                torch.autograd.backward([prediction["log_pi_a"], prediction["entropy"], prediction["v"]], [g_lp, g_en, g_v])
                nn.utils.clip_grad_norm_(self.network.parameters(), config.gradient_clip)
                self.opt.step()
            else:
                approx_kl = r["out"][2]
                if approx_kl <= 1.5 * config.target_kl:          # host decision, as PPO_agent.py:94
                    self.actor_opt.zero_grad()
                    torch.autograd.backward([prediction["log_pi_a"], prediction["entropy"]], [g_lp, g_en])
                    self.actor_opt.step()
                self.critic_opt.zero_grad()
                prediction["v"].backward(g_v)
                self.critic_opt.step()
            return
        ratio = (prediction["log_pi_a"] - entry.log_pi_a).exp()
        obj = ratio * entry.advantage
        obj_clipped = ratio.clamp(1.0 - config.ppo_ratio_clip, 1.0 + config.ppo_ratio_clip) * entry.advantage
        policy_loss = -torch.min(obj, obj_clipped).mean() - config.entropy_weight * prediction["entropy"].mean()
        value_loss = 0.5 * (entry.ret - prediction["v"]).pow(2).mean()
        approx_kl = (entry.log_pi_a - prediction["log_pi_a"]).mean()
        if config.shared_repr:
            self.opt.zero_grad()
            (policy_loss + value_loss).backward()
            nn.utils.clip_grad_norm_(self.network.parameters(), config.gradient_clip)
            self.opt.step()
        else:
            if approx_kl <= 1.5 * config.target_kl:
                self.actor_opt.zero_grad()
                policy_loss.backward()
                self.actor_opt.step()
            self.critic_opt.zero_grad()
            value_loss.backward()
            self.critic_opt.step()

    def step(self):
        config = self.config
        entries = self._rollout()
        self._normalize(entries)
        if config.shared_repr:
            self.lr_scheduler.step(self.total_steps)
        rows = entries.state.size(0)
        # the graphed path keeps the actor's and the critic's parameters in two separate arenas: a shared phi_body (whose
        # parameters the reference gives to BOTH optimizers, network_heads.py:186-189) or discrete (1-D) actions are served
        # by the eager loop below
        shared_phi = len(getattr(self.network, "phi_params", [])) > 0
        if (getattr(config, "graph_minibatch", False) and entries.state.is_cuda and not config.shared_repr
                and rows % config.mini_batch_size == 0 and not shared_phi and entries.action.dim() == 2):
            self._graphed_epochs(entries)                  # same updates, one CUDA-graph replay each (learner.py)
            return
        for _ in range(config.optimization_epochs):
            for batch_indices in random_sample(np.arange(rows), config.mini_batch_size):
                self._minibatch(entries, batch_indices)

    def _graphed_epochs(self, entries):
        """``config.graph_minibatch = True``: the minibatch loop of PPO_agent.py:68-99 through ``GraphedPPOLearner``.  The
        torch optimizers built by ``actor_opt_fn`` / ``critic_opt_fn`` are replaced by flat-arena Adam with the same
        hyper-parameters on first use; the permutations come from ``np.random.permutation`` exactly as ``random_sample``."""
        from ..learner import GraphedPPOLearner, PersistentPPOLearner
        config = self.config
        rows, mb = entries.state.size(0), config.mini_batch_size
        if getattr(self, "_graph", None) is None:
            a = ops.FlatOptimizer.from_torch(self.actor_opt, self.network.actor_params)
            c = ops.FlatOptimizer.from_torch(self.critic_opt, self.network.critic_params)
            # one persistent kernel for the whole loop where the network is the examples' Gaussian MLP pair
            # (config.persistent_minibatch = False keeps the one-graph-replay-per-minibatch form)
            persistent = (getattr(config, "persistent_minibatch", True) and a.kind == "adam" and c.kind == "adam"
                          and PersistentPPOLearner.supported(self.network, mb))
            cls = PersistentPPOLearner if persistent else GraphedPPOLearner
            self._graph = cls(self.network, a, c, rows, entries.state.shape[1], entries.action.shape[1], mb,
                              config.ppo_ratio_clip, config.entropy_weight, config.target_kl,
                              config.optimization_epochs * (rows // mb))
            self._graph.load(entries)
            self._graph.capture()
        g = self._graph
        g.load(entries)
        batches = [b for _ in range(config.optimization_epochs) for b in random_sample(np.arange(rows), mb)]
        g.run(g.set_batches(batches))
        self.last_stats = g.stats
