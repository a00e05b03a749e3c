"""Cycles per phase of the persistent PPO kernel (csrc/ppo_persistent.cu) at the benchmark sizes: installs the clock buffer
(b2rl_ppo_set_phase_clocks), runs N minibatch updates in one launch and prints the mean / median cycles between consecutive
phase barriers (thread 0's clock64), plus the kernel's wall time from CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deeprl_b200 as rl  # noqa: E402
from deeprl_b200 import _lib, ops  # noqa: E402
from deeprl_b200.learner import PersistentPPOLearner  # noqa: E402

rl.select_device(0)
rl.Config.COMPUTE_DTYPE = torch.float32
torch.manual_seed(0)
D, A, rows, mb, nb = 17, 6, 32768, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = rl.GaussianActorCriticNet(D, A, actor_body=rl.FCBody(D, gate=torch.tanh), critic_body=rl.FCBody(D, gate=torch.tanh))
a = ops.FlatOptimizer.from_torch(torch.optim.Adam(net.actor_params, 3e-4), net.actor_params)
c = ops.FlatOptimizer.from_torch(torch.optim.Adam(net.critic_params, 1e-3), net.critic_params)
lr = PersistentPPOLearner(net, a, c, rows, D, A, mb, 0.2, 0.01, 0.01, nb)
dev = a.flat.device
state = torch.randn(rows, D, device=dev)
with torch.no_grad():
    pred = net(state)
lr.buf["state"].copy_(state), lr.buf["action"].copy_(pred["action"]), lr.buf["log_pi_a"].copy_(pred["log_pi_a"])
lr.buf["ret"].normal_(), lr.buf["advantage"].normal_()
perm = np.stack([np.random.permutation(rows)[:mb] for _ in range(nb)])
lr.set_batches(list(perm))
lr.run(nb)                                               # warm-up
torch.cuda.synchronize()
clocks = torch.zeros(2 + 9 * nb + 8, dtype=torch.int64, device=dev)
_lib.call("b2rl_ppo_set_phase_clocks", _lib.ptr(clocks))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lr.run(nb)
e1.record()
torch.cuda.synchronize()
_lib.call("b2rl_ppo_set_phase_clocks", None)
ms = e0.elapsed_time(e1)
ck = clocks.cpu().numpy()
per = ck[1:1 + 9 * nb].reshape(nb, 9) - np.concatenate([[ck[0]], ck[1:9 * nb]]).reshape(nb, 9)
names = ["P1 fwd1", "P2 fwd2", "P3 heads", "P4 loss | critic bwd2", "P5 gate | critic bwd1", "P6 head bwd | prefetch",
         "P7 actor bwd2 | critic update", "P8 actor bwd1", "P9 actor update"]
print("kernel: %.2f ms for %d updates = %.2f us / update; actor steps taken %d" % (ms, nb, ms * 1e3 / nb, int(lr.stats[3])))
tot = per.sum(1)
print("cycles / update: mean %.0f median %.0f" % (tot.mean(), np.median(tot)))
for i, n in enumerate(names):
    print("  %-32s mean %7.0f  median %7.0f  (%4.1f %%)" % (n, per[:, i].mean(), np.median(per[:, i]), 100 * per[:, i].mean() / tot.mean()))
