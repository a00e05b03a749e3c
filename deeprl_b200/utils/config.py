"""Attribute-bag run configuration with the reference's names and defaults.

Mirrors ``deep_rl/utils/config.py:11-89`` (reference): same attribute names, same defaults, the
class-level ``DEVICE`` / ``NOISY_LAYER_STD`` globals, ``eval_env`` setter that derives
``state_dim / action_dim / task_name`` (config.py:74-79) and ``merge`` (config.py:84-89).
The defaults live in one table instead of forty assignments; behaviour is identical.
"""
import argparse

import torch

from .normalizer import RescaleNormalizer

_DEFAULTS = dict(
    task_fn=None, optimizer_fn=None, actor_optimizer_fn=None, critic_optimizer_fn=None,
    network_fn=None, actor_network_fn=None, critic_network_fn=None, replay_fn=None,
    random_process_fn=None, discount=None, target_network_update_freq=None, exploration_steps=None,
    log_level=0, history_length=None, double_q=False, tag="vanilla", num_workers=1, gradient_clip=None,
    entropy_weight=0, use_gae=False, gae_tau=1.0, target_network_mix=0.001, min_memory_size=None,
    max_steps=0, rollout_length=None, value_loss_weight=1.0, iteration_log_interval=30,
    categorical_v_min=None, categorical_v_max=None, categorical_n_atoms=51, num_quantiles=None,
    optimization_epochs=4, mini_batch_size=64, termination_regularizer=0, sgd_update_frequency=None,
    random_action_prob=None, log_interval=int(1e3), save_interval=0, eval_interval=0, eval_episodes=10,
    async_actor=True, tasks=False, decaying_lr=False, shared_repr=False, noisy_linear=False, n_step=1,
)


class Config:
    DEVICE = torch.device("cpu")
    COMPUTE_DTYPE = torch.float32      # torch.bfloat16 = throughput mode of the dense contractions (not in the reference)
    DENSE_BACKEND = "tcgen05"          # bf16 mode: "tcgen05" = csrc/gemm.cu for the whole NatureConvBody, "library" = cuDNN/cuBLAS
    NOISY_LAYER_STD = 0.1
    DEFAULT_REPLAY = "replay"
    PRIORITIZED_REPLAY = "prioritized_replay"

    def __init__(self):
        self.parser = argparse.ArgumentParser()
        for name, value in _DEFAULTS.items():
            setattr(self, name, value)
        self.state_normalizer = RescaleNormalizer()
        self.reward_normalizer = RescaleNormalizer()
        self.replay_type = Config.DEFAULT_REPLAY
        self._eval_env = None

    @property
    def eval_env(self):
        return self._eval_env

    @eval_env.setter
    def eval_env(self, env):
        self._eval_env = env
        self.state_dim, self.action_dim, self.task_name = env.state_dim, env.action_dim, env.name

    def add_argument(self, *args, **kwargs):
        self.parser.add_argument(*args, **kwargs)

    def merge(self, config_dict=None):
        if config_dict is None:
            config_dict = vars(self.parser.parse_args())
        for key, value in config_dict.items():
            setattr(self, key, value)
