"""TEST INFRASTRUCTURE -- torch-CPU restatements of the per-agent target/loss arithmetic, at the
LOSS-KERNEL BOUNDARY: the inputs are network outputs (q / prob / log_prob / quantile tensors),
not states, so fp32 parity at 1e-5 is meaningful.  Every function
follows the reference's operation order; citations are deep_rl/agent/*.py.
"""
import numpy as np
import torch


def huber(x, k=1.0):                                                    # utils/torch_utils.py:47-48
    return torch.where(x.abs() < k, 0.5 * x.pow(2), k * (x.abs() - 0.5 * k))


# ----------------------------------------------------------------------------------- DQN
def dqn_delta(q, q_next_target, q_next_online, action, reward, mask, discount_n):
    """DQN_agent.py:86-99 -> per-sample ``q_target - q`` (the tensor compute_loss returns)."""
    if q_next_online is not None:                                       # double_q, :88-90
        best = torch.argmax(q_next_online, dim=-1)
        q_next = q_next_target.gather(1, best.unsqueeze(-1)).squeeze(1)
    else:
        q_next = q_next_target.max(1)[0]                                # :92
    q_target = reward + discount_n * q_next * mask                      # :95
    q_sa = q.gather(1, action.long().unsqueeze(-1)).squeeze(-1)         # :97-98
    return q_target - q_sa


def dqn_reduce(loss):                                                   # DQN_agent.py:78-79
    return loss.pow(2).mul(0.5).mean()


def per_block(delta, sampling_prob, beta, eps, alpha):
    """DQN_agent.py:120-127 -> (priorities, IS weights, weighted delta)."""
    prio = delta.abs().add(eps).pow(alpha)
    w = sampling_prob.mul(sampling_prob.size(0)).add(1e-6).pow(-beta)
    w = w / w.max()
    return prio, w, delta.mul(w)


# ----------------------------------------------------------------------------------- C51
def c51_kl(log_prob, prob_next_target, prob_next_online, action, reward, mask, atoms,
           v_min, v_max, discount_n):
    """CategoricalDQN_agent.py:60-86 -> per-sample KL (B,)."""
    B = log_prob.size(0)
    rows = torch.arange(B)
    delta_atom = (v_max - v_min) / float(atoms.numel() - 1)             # :46
    q_next = (prob_next_target * atoms).sum(-1)                         # :66
    if prob_next_online is not None:                                    # :67-68
        a_next = torch.argmax((prob_next_online * atoms).sum(-1), dim=-1)
    else:
        a_next = torch.argmax(q_next, dim=-1)                           # :70
    prob_next = prob_next_target[rows, a_next, :]
    r = reward.unsqueeze(-1)
    m = mask.unsqueeze(-1)
    atoms_target = r + discount_n * m * atoms.view(1, -1)               # :75
    atoms_target = atoms_target.clamp(v_min, v_max).unsqueeze(1)        # :76-77
    w = (1 - (atoms_target - atoms.view(1, -1, 1)).abs() / delta_atom).clamp(0, 1)
    target_prob = (w * prob_next.unsqueeze(1)).sum(-1)                  # :78-80
    lp = log_prob[rows, action.long(), :]                               # :82-84
    return (target_prob * target_prob.add(1e-5).log() - target_prob * lp).sum(-1)   # :85


# ----------------------------------------------------------------------------------- QR-DQN
def qr_loss(quantile, quantile_next_target, action, reward, mask, discount_n):
    """QuantileRegressionDQN_agent.py:55-74 -> vector (N,) indexed by TARGET quantile j."""
    B, _, N = quantile.shape
    rows = torch.arange(B)
    tau = torch.tensor((2 * np.arange(N) + 1) / (2.0 * N), dtype=torch.float32).view(1, -1)   # :44-45
    a_next = torch.argmax(quantile_next_target.sum(-1), dim=-1)         # :60
    qn = quantile_next_target[rows, a_next, :]
    qn = reward.unsqueeze(-1) + discount_n * mask.unsqueeze(-1) * qn    # :65
    q = quantile[rows, action.long(), :]                                # :67-69
    diff = qn.t().unsqueeze(-1) - q                                     # :71-72  (N_j, B, N_i)
    loss = huber(diff) * (tau - (diff.detach() < 0).float()).abs()      # :73
    return loss.sum(-1).mean(1)                                         # :74


# ----------------------------------------------------------------------------------- GAE
def gae(reward, mask, value, discount, tau, use_gae=True):
    """A2C_agent.py:43-53 == PPO_agent.py:51-61.  reward, mask: (T,N,1); value: (T+1,N,1)."""
    T = reward.shape[0]
    adv = torch.zeros_like(value[0])
    ret = value[T].clone()
    advs, rets = [None] * T, [None] * T
    for i in reversed(range(T)):
        ret = reward[i] + discount * mask[i] * ret
        if not use_gae:
            adv = ret - value[i]
        else:
            td = reward[i] + discount * mask[i] * value[i + 1] - value[i]
            adv = adv * tau * discount * mask[i] + td
        advs[i], rets[i] = adv, ret
    return torch.stack(advs), torch.stack(rets)


def normalize_advantage(adv):                                           # PPO_agent.py:66 (unbiased std, no eps)
    return (adv - adv.mean()) / adv.std()


# ----------------------------------------------------------------------------------- PPO / A2C
def ppo_losses(log_pi_a, entropy, v, old_log_pi_a, advantage, ret, clip, entropy_weight):
    """PPO_agent.py:77-86 -> (policy_loss, value_loss, approx_kl)."""
    ratio = (log_pi_a - old_log_pi_a).exp()
    obj = ratio * advantage
    obj_clipped = ratio.clamp(1.0 - clip, 1.0 + clip) * advantage
    policy_loss = -torch.min(obj, obj_clipped).mean() - entropy_weight * entropy.mean()
    value_loss = 0.5 * (ret - v).pow(2).mean()
    approx_kl = (old_log_pi_a - log_pi_a).mean()
    return policy_loss, value_loss, approx_kl


def a2c_loss(log_pi_a, v, ret, advantage, entropy, entropy_weight, value_loss_weight):
    """A2C_agent.py:55-62 -> scalar objective that is back-propagated."""
    policy_loss = -(log_pi_a * advantage).mean()
    value_loss = 0.5 * (ret - v).pow(2).mean()
    entropy_loss = entropy.mean()
    return policy_loss - entropy_weight * entropy_loss + value_loss_weight * value_loss
