"""Golden fixture for n-step Q-learning: ``NStepDQNAgent.step`` (agent/NStepDQN_agent.py:26-70) of the UNMODIFIED
reference, imported through oracle/ref_shim.py in the build container, on the synthetic CartPole task with the
``n_step_dqn_feature`` wiring (examples.py:408-424; a faster epsilon schedule and target sync so both branches are exercised).
Writes tests/golden/nstep.npz: initial weights, the whole env interaction (so a replayed Task reproduces it), the n-step
returns / q-values / actions captured at ``Storage.extract``, and the parameter vector after every step."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shim import import_reference  # noqa: E402

ref = import_reference()
from deeprl_b200.component.envs import Task as SynthTask  # noqa: E402  (host env, duck-typed for the reference)

torch.set_num_threads(1)
ref.select_device(-1)


class NullLogger:
    def info(self, *a, **k): pass
    debug = warning = add_scalar = add_histogram = info


captured = {}
real_extract = ref.Storage.extract


def extract(self, keys):
    T = self.memory_size
    for k in ("q", "action", "ret", "reward", "mask"):
        captured[k] = torch.stack(list(getattr(self, k)[:T])).detach().numpy().copy()
    return real_extract(self, keys)


ref.Storage.extract = extract
np.random.seed(21), torch.manual_seed(21)
N, T = 5, 5
c = ref.Config()
c.merge(dict(tag=None))
c.num_workers = N
c.task_fn = lambda: SynthTask("CartPole-v0", num_envs=N, seed=8)
c.eval_env = SynthTask("CartPole-v0", seed=8)
c.optimizer_fn = lambda p: torch.optim.RMSprop(p, 0.001)
c.network_fn = lambda: ref.VanillaNet(c.action_dim, ref.FCBody(c.state_dim))
c.random_action_prob = ref.LinearSchedule(0.6, 0.1, 200)
c.discount, c.target_network_update_freq, c.rollout_length, c.gradient_clip = 0.99, 12, T, 5
ag = ref.NStepDQNAgent(c)
ag.logger = NullLogger()
out = {"state0": np.asarray(ag.states, np.float32)}
seen = []
real_step = ag.task.step


def task_step(a):
    r = real_step(a)
    seen.append((np.asarray(a).copy(), np.asarray(r[0], np.float32), np.asarray(r[1], np.float64), np.asarray(r[2])))
    return r


ag.task.step = task_step
init = {k: v.detach().numpy().copy() for k, v in ag.network.state_dict().items()}
P, PT = [], []
G = {k: [] for k in ("q", "action", "ret", "reward", "mask")}
for it in range(16):
    ag.step()
    P.append(np.concatenate([p.detach().numpy().ravel() for p in ag.network.parameters()]))
    PT.append(np.concatenate([p.detach().numpy().ravel() for p in ag.target_network.parameters()]))
    for k in G:
        G[k].append(captured[k])
for k, v in init.items():
    out["init." + k] = v
out["keys"] = np.asarray(list(init.keys()))
out["actions"] = np.stack([s[0] for s in seen])
out["next_states"] = np.stack([s[1] for s in seen])
out["rewards"] = np.stack([s[2] for s in seen])
out["dones"] = np.stack([s[3] for s in seen])
out["params"] = np.stack(P)
out["target_params"] = np.stack(PT)
for k in G:
    out["cap_" + k] = np.stack(G[k])
np.savez_compressed(os.path.join(HERE, "nstep.npz"), **out)
print("nstep.npz", len(out), "arrays; terminals seen:", int(out["dones"].sum()), "distinct target snapshots:",
      len({PT[i].tobytes() for i in range(len(PT))}))
