"""TEST INFRASTRUCTURE -- functional torch-CPU restatement of the reference networks on the path
(deep_rl/network/network_bodies.py:10-33,50-73; network_heads.py:11-102,173-255), driven by a
``state_dict`` with the reference's parameter names so weights can be shared with the product.
"""
import torch
import torch.nn.functional as F


def nature_body(sd, x, prefix="body."):                                # network_bodies.py:27-33
    y = F.relu(F.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], stride=4))
    y = F.relu(F.conv2d(y, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], stride=2))
    y = F.relu(F.conv2d(y, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"], stride=1))
    y = y.reshape(y.size(0), -1)
    return F.relu(F.linear(y, sd[prefix + "fc4.weight"], sd[prefix + "fc4.bias"]))


def fc_body(sd, x, prefix, gate=F.relu):                                # network_bodies.py:70-73
    i = 0
    while prefix + "layers.%d.weight" % i in sd:
        x = gate(F.linear(x, sd[prefix + "layers.%d.weight" % i], sd[prefix + "layers.%d.bias" % i]))
        i += 1
    return x


def vanilla_q(sd, phi):                                                 # network_heads.py:18-21
    return F.linear(phi, sd["fc_head.weight"], sd["fc_head.bias"])


def dueling_q(sd, phi):                                                 # network_heads.py:32-37
    v = F.linear(phi, sd["fc_value.weight"], sd["fc_value.bias"])
    a = F.linear(phi, sd["fc_advantage.weight"], sd["fc_advantage.bias"])
    return v.expand_as(a) + (a - a.mean(1, keepdim=True).expand_as(a))


def categorical(sd, phi, action_dim, num_atoms):                        # network_heads.py:49-54
    pre = F.linear(phi, sd["fc_categorical.weight"], sd["fc_categorical.bias"]).view(-1, action_dim, num_atoms)
    return F.softmax(pre, dim=-1), F.log_softmax(pre, dim=-1)


def quantile(sd, phi, action_dim, num_quantiles):                       # network_heads.py:97-102
    return F.linear(phi, sd["fc_quantiles.weight"], sd["fc_quantiles.bias"]).view(-1, action_dim, num_quantiles)


def gaussian_actor_critic(sd, obs, action, gate=torch.tanh):            # network_heads.py:198-214
    phi_a = fc_body(sd, obs, "actor_body.", gate)
    phi_v = fc_body(sd, obs, "critic_body.", gate)
    mean = torch.tanh(F.linear(phi_a, sd["fc_action.weight"], sd["fc_action.bias"]))
    v = F.linear(phi_v, sd["fc_critic.weight"], sd["fc_critic.bias"])
    dist = torch.distributions.Normal(mean, F.softplus(sd["std"]))
    return dict(log_pi_a=dist.log_prob(action).sum(-1).unsqueeze(-1),
                entropy=dist.entropy().sum(-1).unsqueeze(-1), mean=mean, v=v)


def categorical_actor_critic(sd, obs, action=None, gate=torch.tanh):   # network_heads.py:239-255 (shared phi_body)
    phi = fc_body(sd, obs, "phi_body.", gate)
    logits = F.linear(phi, sd["fc_action.weight"], sd["fc_action.bias"])
    v = F.linear(phi, sd["fc_critic.weight"], sd["fc_critic.bias"])
    dist = torch.distributions.Categorical(logits=logits)
    if action is None:
        action = dist.sample()
    return dict(action=action, log_pi_a=dist.log_prob(action).unsqueeze(-1),
                entropy=dist.entropy().unsqueeze(-1), v=v)
