"""Device-side actor step for the on-policy agents: the work ``PPOAgent`` / ``A2CAgent`` do per env step between
two ``task.step()`` calls (PPO_agent.py:45-50) -- ``MeanStdNormalizer`` (normalizer.py:36-51), the ``GaussianActorCriticNet``
forward (network_heads.py:173-214) and the Normal sample / log-prob / entropy -- as ONE launch of ``b2rl_gaussian_actor_step``
(csrc/actor.cu) on a pinned, double-buffered observation upload.  The envs stay on the host (north_star)."""
import numpy as np
import torch

from .. import _lib
from ..network.network_bodies import DummyBody, FCBody
from ..utils.normalizer import MeanStdNormalizer, RunningMoments

_f32, _f64 = torch.float32, torch.float64


def supported(network, normalizer):
    """GaussianActorCriticNet on a CUDA device with DummyBody phi, two-layer tanh FCBody actor / critic bodies and a
    MeanStdNormalizer (or no normaliser state at all)."""
    try:
        from ..network.network_heads import GaussianActorCriticNet
        ok = (isinstance(network, GaussianActorCriticNet) and isinstance(network.phi_body, DummyBody)
              and all(isinstance(b, FCBody) and len(b.layers) == 2 and b.gate is torch.tanh and not b.noisy_linear
                      for b in (network.actor_body, network.critic_body))
              and network.fc_action.weight.is_cuda and network.fc_action.out_features <= 32
              and max(network.actor_body.layers[0].out_features, network.actor_body.layers[1].out_features) <= 128
              and network.actor_body.layers[0].in_features <= 128
              and isinstance(normalizer, MeanStdNormalizer))
        return bool(ok)
    except Exception:                                    # noqa: BLE001
        return False


class DeviceGaussianActor:
    def __init__(self, network, normalizer, num_envs, seed=0):
        self.net, self.norm = network, normalizer
        dev = network.fc_action.weight.device
        self.dev = dev
        self.N, self.D = int(num_envs), network.actor_body.layers[0].in_features
        self.A = network.fc_action.out_features
        if self.N > 64:
            raise _lib.B2RLError("the device actor serves at most 64 workers per launch")
        self.h_obs = [torch.zeros((self.N, self.D), dtype=_f32, pin_memory=True) for _ in range(2)]
        self.d_obs = [torch.zeros((self.N, self.D), dtype=_f32, device=dev) for _ in range(2)]
        self.slot = 0
        self.rm_mean = torch.zeros(self.D, dtype=_f64, device=dev)
        self.rm_var = torch.ones(self.D, dtype=_f64, device=dev)
        self.rm_count = torch.full((1,), 1e-4, dtype=_f64, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.seed = int(seed)
        self.push_stats()

    # ---- the host normaliser object stays the checkpointed / evaluated one: moments move between it and the device
    def push_stats(self):
        rms = self.norm.rms
        if rms is not None:
            self.rm_mean.copy_(torch.from_numpy(np.asarray(rms.mean, dtype=np.float64).reshape(-1)))
            self.rm_var.copy_(torch.from_numpy(np.asarray(rms.var, dtype=np.float64).reshape(-1)))
            self.rm_count.fill_(float(rms.count))

    def pull_stats(self):
        if self.norm.rms is None:
            self.norm.rms = RunningMoments(shape=(1, self.D))
        rms = self.norm.rms
        rms.mean = self.rm_mean.cpu().numpy().reshape(np.shape(rms.mean))
        rms.var = self.rm_var.cpu().numpy().reshape(np.shape(rms.var))
        rms.count = float(self.rm_count.item())

    def step(self, raw_obs, z=None, given_action=None, update=None):
        """raw (un-normalised) observations [N, D] from the envs -> dict(state (normalised), action, log_pi_a, entropy, mean,
        v): device tensors shaped like GaussianActorCriticNet.forward's.  ``z``: supplied standard normals (parity mode)."""
        n = self.net
        k = self.slot
        self.slot = 1 - k
        self.h_obs[k].numpy()[...] = np.asarray(raw_obs, dtype=np.float32).reshape(self.N, self.D)
        self.d_obs[k].copy_(self.h_obs[k], non_blocking=True)
        out = dict(state=torch.empty((self.N, self.D), dtype=_f32, device=self.dev),
                   action=torch.empty((self.N, self.A), dtype=_f32, device=self.dev),
                   log_pi_a=torch.empty((self.N, 1), dtype=_f32, device=self.dev),
                   entropy=torch.empty((self.N, 1), dtype=_f32, device=self.dev),
                   mean=torch.empty((self.N, self.A), dtype=_f32, device=self.dev),
                   v=torch.empty((self.N, 1), dtype=_f32, device=self.dev))
        ab, cb = n.actor_body.layers, n.critic_body.layers
        w = lambda m: _lib.ptr(m.weight.detach())
        b = lambda m: _lib.ptr(m.bias.detach())
        ro = int(bool(update)) if update is not None else (0 if self.norm.read_only else 1)
        _lib.call("b2rl_gaussian_actor_step", _lib.ptr(self.d_obs[k]), _lib.ptr(self.rm_mean), _lib.ptr(self.rm_var),
                  _lib.ptr(self.rm_count), ro, float(self.norm.clip), float(self.norm.epsilon),
                  w(ab[0]), b(ab[0]), w(ab[1]), b(ab[1]), w(n.fc_action), b(n.fc_action),
                  w(cb[0]), b(cb[0]), w(cb[1]), b(cb[1]), w(n.fc_critic), b(n.fc_critic), _lib.ptr(n.std.detach()),
                  self.N, self.D, ab[0].out_features, ab[1].out_features, self.A,
                  _lib.ptr(None if z is None else z.contiguous()), self.seed, _lib.ptr(self.counter),
                  _lib.ptr(None if given_action is None else given_action.contiguous()),
                  _lib.ptr(out["state"]), _lib.ptr(out["action"]), _lib.ptr(out["log_pi_a"]), _lib.ptr(out["entropy"]),
                  _lib.ptr(out["mean"]), _lib.ptr(out["v"]), _lib.stream())
        return out
