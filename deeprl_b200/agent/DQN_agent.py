"""DQN / Double-DQN / prioritized / n-step agent behind the reference's interface
(``deep_rl/agent/DQN_agent.py``: ``DQNActor``:14, ``DQNAgent``:48).

One ``step()`` = ``sgd_update_frequency`` actor transitions fed to the HBM replay ring, then (after the
exploration phase) one gradient update: sample -> networks -> ONE fused target/loss/priority/gradient kernel
(``csrc/losses.cu``) -> backward -> ONE fused global-norm clip + optimizer launch pair (``csrc/optim.cu``) ->
sum-tree priority update on the device (``csrc/sumtree.cu``).  Nothing of the update touches the host.

``compute_loss`` / ``reduce_loss`` keep the reference's contract (per-sample tensor, then reduction,
DQN_agent.py:78-99) and differentiate through the same kernel; a subclass that overrides either is served by
the generic autograd path in ``_generic_update``.
"""
import threading

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..component import LazyFrames, PrioritizedTransition
from ..network.fused import frame_scale
from ..utils import Config, RescaleNormalizer, close_obj, epsilon_greedy, tensor, to_np
from .BaseAgent import BaseActor, BaseAgent


class DQNActor(BaseActor):
    def __init__(self, config):
        BaseActor.__init__(self, config)
        self.config = config
        self.start()

    def _q_tensor(self, prediction):
        """Action values [num_envs, A] on the device (what ``compute_q`` converts to numpy)."""
        return prediction["q"]

    def compute_q(self, prediction):
        return to_np(self._q_tensor(prediction))

    def _graphed(self):
        """``config.cuda_graph``: the forward pass below as one captured launch sequence (component/actor.py GraphedQActor).
        Not for subclasses that redefine ``compute_q`` itself (they get the statements of the reference)."""
        ga = getattr(self, "_graph_actor", None)
        if ga is None:
            from ..component import actor as device_actor
            ok = (type(self).compute_q is DQNActor.compute_q and device_actor.q_actor_supported(self.config, self._network)
                  and all(np.asarray(s).dtype == np.uint8 and np.asarray(s).shape == (4, 84, 84) for s in self._state))
            ga = self._graph_actor = device_actor.GraphedQActor(
                self._network, self._q_tensor, len(self._state), 4, (84, 84), self.config.state_normalizer.coef) if ok else False
        return ga or None

    def _transition(self):
        """DQN_agent.py:24-45: epsilon-greedy on a forward pass of the shared network, one env step."""
        if self._state is None:
            self._state = self._task.reset()
        config = self.config
        if config.noisy_linear:
            self._network.reset_noise()
        ga = self._graphed() if getattr(config, "cuda_graph", False) else None
        if ga is not None:
            with config.lock:
                q_values = ga.q_values(self._state)
        else:
            with config.lock, torch.no_grad():
                prediction = self._network(config.state_normalizer(np.asarray([np.asarray(s) for s in self._state])))
            q_values = self.compute_q(prediction)
        if config.noisy_linear:
            epsilon = 0
        elif self._total_steps < config.exploration_steps:
            epsilon = 1
        else:
            epsilon = config.random_action_prob()
        action = epsilon_greedy(epsilon, q_values)
        next_state, reward, done, info = self._task.step(action)
        entry = [self._state, action, reward, next_state, done, info]
        self._total_steps += 1
        self._state = next_state
        return entry


def _pre_normalized(config, replay):
    """True when the state normalizer is a pure rescale of a uint8 frame ring, so that the fused
    gather->normalize kernel can produce the network input directly (ImageNormalizer, normalizer.py:64-66)."""
    norm = config.state_normalizer
    inner = getattr(replay, "replay", replay)
    return (isinstance(norm, RescaleNormalizer) and hasattr(inner, "sample_normalized")
            and getattr(inner, "item_dtype", None) == np.uint8 and not getattr(replay, "async_", False))


class DQNAgent(BaseAgent):
    def __init__(self, config):
        BaseAgent.__init__(self, config)
        self.config = config
        config.lock = threading.Lock()
        self._build(DQNActor)

    def _build(self, actor_cls):
        config = self.config
        self.replay = config.replay_fn()
        self.actor = actor_cls(config)
        self.network = config.network_fn()
        self.target_network = config.network_fn()
        self.target_network.load_state_dict(self.network.state_dict())
        self.optimizer = config.optimizer_fn(self.network.parameters())
        self._flat = None
        if next(self.network.parameters()).is_cuda:
            try:
                self._flat = ops.FlatOptimizer.from_torch(self.optimizer, list(self.network.parameters()))
            except NotImplementedError:
                self._flat = None             # exotic optimizer: torch.optim + clip_grad_norm_ (library path)
        self.actor.set_network(self.network)
        self.total_steps = 0
        self.last_loss = None

    def close(self):
        close_obj(self.replay)
        close_obj(self.actor)

    def eval_step(self, state):
        self.config.state_normalizer.set_read_only()
        state = self.config.state_normalizer(np.asarray([np.asarray(s) for s in state]))
        with torch.no_grad():
            q = self.network(state)["q"]
        action = to_np(q.argmax(-1))
        self.config.state_normalizer.unset_read_only()
        return action

    # ------------------------------------------------------------------ reference-shaped loss API
    def reduce_loss(self, loss):
        return loss.pow(2).mul(0.5).mean()

    def _inputs(self, transitions):
        if getattr(transitions, "_normalized", False) or getattr(self, "_batch_is_normalized", False):
            return transitions.state, transitions.next_state
        norm = self.config.state_normalizer
        return norm(transitions.state), norm(transitions.next_state)

    def compute_loss(self, transitions):
        """DQN_agent.py:81-99 -> per-sample ``q_target - q`` (autograd through q)."""
        config = self.config
        states, next_states = self._inputs(transitions)
        with torch.no_grad():
            q_next_t = self.target_network(next_states)["q"]
            q_next_o = self.network(next_states)["q"] if config.double_q else None
        q = self.network(states)["q"]
        return ops.dqn_delta(q, q_next_t, q_next_o, tensor(transitions.action), tensor(transitions.reward),
                             tensor(transitions.mask), config.discount ** config.n_step)

    # ------------------------------------------------------------------ update
    def _fused_forward(self, transitions, per):
        """Network passes + the fused loss kernel.  Returns (output tensor to back-propagate, gradient, result)."""
        config = self.config
        states, next_states = self._inputs(transitions)
        with torch.no_grad():
            q_next_t = self.target_network(next_states)["q"]
            q_next_o = self.network(next_states)["q"] if config.double_q else None
        q = self.network(states)["q"]
        r = ops.dqn_loss_fused(q.detach(), q_next_t, q_next_o, transitions.action, transitions.reward, transitions.mask,
                               config.discount ** config.n_step, **per)
        return q, r["dq"], r

    def _per_args(self, transitions):
        if not isinstance(transitions, PrioritizedTransition):
            return {}
        c = self.config
        return dict(is_prob=tensor(transitions.sampling_prob), beta=c.replay_beta(), eps=c.replay_eps, alpha=c.replay_alpha)

    def _apply_gradients(self):
        config = self.config
        if self._flat is not None:
            with config.lock:
                self._flat.step(max_norm=config.gradient_clip or 0.0)
        else:
            if config.gradient_clip:
                nn.utils.clip_grad_norm_(self.network.parameters(), config.gradient_clip)
            with config.lock:
                self.optimizer.step()

    def _zero_grad(self):
        if self._flat is not None:
            self._flat.zero_grad()
        else:
            self.optimizer.zero_grad()

    def _fused_update(self, transitions):
        per = self._per_args(transitions)
        out, grad, r = self._fused_forward(transitions, per)
        if per:
            self.replay.update_priorities((transitions.idx, r["priority"]))      # DQN_agent.py:120-123, on the device
        self._zero_grad()
        out.backward(grad)
        self._apply_gradients()
        return r["loss"]

    def _generic_update(self, transitions):
        """DQN_agent.py:119-134 verbatim in torch, for subclasses that override compute_loss / reduce_loss."""
        config = self.config
        loss = self.compute_loss(transitions)
        if isinstance(transitions, PrioritizedTransition):
            priorities = loss.detach().abs().add(config.replay_eps).pow(config.replay_alpha)
            self.replay.update_priorities((tensor(transitions.idx).long(), priorities.float().contiguous()))
            sampling_probs = tensor(transitions.sampling_prob)
            weights = sampling_probs.mul(sampling_probs.size(0)).add(1e-6).pow(-config.replay_beta())
            weights = weights / weights.max()
            loss = loss.mul(weights)
        loss = self.reduce_loss(loss)
        self._zero_grad()
        loss.backward()
        self._apply_gradients()
        return loss.detach()

    _fused_methods = ("compute_loss", "reduce_loss")

    def _uses_reference_hooks(self):
        cls = type(self)
        owner = self._fused_owner()
        return any(getattr(cls, m) is not getattr(owner, m) for m in self._fused_methods)

    def _fused_owner(self):
        return DQNAgent

    def _sample(self):
        config = self.config
        inner = getattr(self.replay, "replay", self.replay)
        self._batch_scale = 1.0
        if _pre_normalized(config, self.replay) and len(inner.item_shape) == 2:
            dtype, coef = Config.COMPUTE_DTYPE, config.state_normalizer.coef
            self._batch_is_normalized = True
            if dtype == torch.bfloat16 and inner.item_shape[0] % 4 == 0 and inner.item_shape[1] % 4 == 0:
                # throughput mode: exact integer frames in space-to-depth layout, 1/255 folded into conv1
                self._batch_scale = coef
                return inner.sample_normalized(out_dtype=dtype, scale=None, layout="s2d")
            return inner.sample_normalized(out_dtype=dtype, scale=coef, layout="nchw")     # parity mode: exact LUT
        self._batch_is_normalized = False
        return self.replay.sample()

    def step(self):
        config = self.config
        transitions = self.actor.step()
        feeds = []
        for states, actions, rewards, next_states, dones, info in transitions:
            self.record_online_return(info)
            self.total_steps += 1
            feeds.append(dict(
                state=np.array([s[-1] if isinstance(s, LazyFrames) else s for s in states]),
                action=actions,
                reward=[config.reward_normalizer(r) for r in rewards],
                mask=1 - np.asarray(dones, dtype=np.int32)))
        feed_many = getattr(self.replay, "feed_many", None)
        if feed_many is not None:                          # one staging upload for the env steps of this agent step;
            feed_many(feeds)                               # ring / tree state as after feed(d) for d in feeds (replay.py:75-90)
        else:
            for d in feeds:
                self.replay.feed(d)

        if self.total_steps > config.exploration_steps and self._graph_ok():
            self._graph_update()                           # config.cuda_graph: the whole update is one graph replay
        elif self.total_steps > config.exploration_steps:
            transitions = self._sample()
            if config.noisy_linear:
                self.target_network.reset_noise()
                self.network.reset_noise()
            with frame_scale(self._batch_scale):
                if self._uses_reference_hooks():
                    self.last_loss = self._generic_update(transitions)
                else:
                    self.last_loss = self._fused_update(transitions)

        if self.total_steps / config.sgd_update_frequency % config.target_network_update_freq == 0:
            if getattr(self, "_learner", None) is not None:
                self._learner.sync_target()                # load_state_dict + re-pack of the target's bf16 operands
            else:
                self.target_network.load_state_dict(self.network.state_dict())

    # ------------------------------------------------------------------ config.cuda_graph (opt-in)
    _graph_kind = "dqn"

    def _graph_ok(self):
        """The captured update (learner.GraphedDQNLearner) serves the stock agents on a CUDA device with a synchronous uint8
        image replay and the fused optimizer; everything else keeps the eager path."""
        config = self.config
        if not getattr(config, "cuda_graph", False) or self._flat is None or self._uses_reference_hooks():
            return False
        if config.noisy_linear or not _pre_normalized(config, self.replay):
            return False
        inner = getattr(self.replay, "replay", self.replay)
        return len(getattr(inner, "item_shape", ())) == 2 and inner.size() >= inner.batch_size + 64

    def _graph_update(self):
        from ..learner import GraphedDQNLearner
        config = self.config
        lr = getattr(self, "_learner", None)
        if lr is None:
            inner = getattr(self.replay, "replay", self.replay)
            lr = self._learner = GraphedDQNLearner(
                self.network, self.target_network, self._flat, inner, kind=self._graph_kind, discount=config.discount,
                n_step=config.n_step, double_q=bool(config.double_q), gradient_clip=config.gradient_clip or 0.0,
                feeds_per_update=0, compute_dtype=Config.COMPUTE_DTYPE, state_scale=config.state_normalizer.coef,
                replay_eps=getattr(config, "replay_eps", 0.01), replay_alpha=getattr(config, "replay_alpha", 0.5),
                categorical=(getattr(config, "categorical_v_min", -10.0), getattr(config, "categorical_v_max", 10.0)),
                target_sync_every=0, prefetch=False)
            with config.lock:
                lr.capture(warmup=1)                       # NOTE: the warm-up is one real (extra) gradient update
        if lr.per:
            lr.d_beta.fill_(float(config.replay_beta()))
        with config.lock:                                  # the actor reads the shared parameters (DQN_agent.py:30,133)
            lr.update()
            if lr.tail() is None:                          # (the fused tail writes the packed operands itself)
                lr._repack(self.network, lr.scale if lr.dtype == torch.bfloat16 else 1.0)   # the actor's next forward sees theta_k
        self.last_loss = lr.loss
