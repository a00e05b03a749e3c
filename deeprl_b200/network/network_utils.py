#######################################################################
# This file restates an interface of ShangtongZhang/DeepRL, whose     #
# licence asks that the following declaration stay at the top:        #
#                                                                     #
# Copyright (C) 2017 Shangtong Zhang(zhangshangtong.cpp@gmail.com)    #
# Permission given to modify the code as long as you keep this        #
# declaration at the top                                              #
#######################################################################
"""``layer_init`` / ``NoisyLinear`` / ``BaseNet`` with the reference's names
(``deep_rl/network/network_utils.py:15-83``)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..utils import Config, tensor  # noqa: F401


class BaseNet:
    def reset_noise(self):
        pass


def layer_init(layer, w_scale=1.0):
    """network_utils.py:23-27 -- orthogonal weight (times ``w_scale``), zero bias."""
    nn.init.orthogonal_(layer.weight.data)
    layer.weight.data.mul_(w_scale)
    nn.init.constant_(layer.bias.data, 0)
    return layer


class NoisyLinear(nn.Module):
    """Factorised-Gaussian noisy layer (Fortunato et al. 2017), parameter / buffer names of
    network_utils.py:31-83 so reference ``state_dict``s load."""

    def __init__(self, in_features, out_features, std_init=0.4):
        super().__init__()
        self.in_features, self.out_features, self.std_init = in_features, out_features, std_init
        self.weight_mu = nn.Parameter(torch.zeros(out_features, in_features))
        self.weight_sigma = nn.Parameter(torch.zeros(out_features, in_features))
        self.register_buffer("weight_epsilon", torch.zeros(out_features, in_features))
        self.bias_mu = nn.Parameter(torch.zeros(out_features))
        self.bias_sigma = nn.Parameter(torch.zeros(out_features))
        self.register_buffer("bias_epsilon", torch.zeros(out_features))
        self.register_buffer("noise_in", torch.zeros(in_features))
        self.register_buffer("noise_out_weight", torch.zeros(out_features))
        self.register_buffer("noise_out_bias", torch.zeros(out_features))
        self.reset_parameters()
        self.reset_noise()

    def forward(self, x):
        if self.training:
            w = self.weight_mu + self.weight_sigma * self.weight_epsilon
            b = self.bias_mu + self.bias_sigma * self.bias_epsilon
        else:
            w, b = self.weight_mu, self.bias_mu
        return F.linear(x, w, b)

    def reset_parameters(self):
        r = 1 / math.sqrt(self.weight_mu.size(1))
        self.weight_mu.data.uniform_(-r, r)
        self.weight_sigma.data.fill_(self.std_init / math.sqrt(self.weight_sigma.size(1)))
        self.bias_mu.data.uniform_(-r, r)
        self.bias_sigma.data.fill_(self.std_init / math.sqrt(self.bias_sigma.size(0)))

    def reset_noise(self):
        self.noise_in.normal_(std=Config.NOISY_LAYER_STD)
        self.noise_out_weight.normal_(std=Config.NOISY_LAYER_STD)
        self.noise_out_bias.normal_(std=Config.NOISY_LAYER_STD)
        self.weight_epsilon.copy_(torch.outer(self.transform_noise(self.noise_out_weight),
                                              self.transform_noise(self.noise_in)))
        self.bias_epsilon.copy_(self.transform_noise(self.noise_out_bias))

    @staticmethod
    def transform_noise(x):
        return x.sign().mul(x.abs().sqrt())
