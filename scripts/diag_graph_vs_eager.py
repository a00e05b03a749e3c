"""Isolate which feature makes the captured learner differ from the eager one (diagnostic)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deeprl_b200 as rl  # noqa: E402
from deeprl_b200.learner import GraphedDQNLearner  # noqa: E402

rl.select_device(0)
rl.Config.COMPUTE_DTYPE = torch.bfloat16
bench.CAP = 30_000
dev = torch.device("cuda", 0)


def make(per, dueling, double_q, side_wgrad=True):
    torch.manual_seed(0)
    body = lambda: rl.NatureConvBody(in_channels=4)
    mk = (lambda: rl.DuelingNet(4, body())) if dueling else (lambda: rl.VanillaNet(4, body()))
    net, tgt = mk(), mk()
    tgt.load_state_dict(net.state_dict())
    opt = rl.ops.FlatOptimizer.from_torch(torch.optim.RMSprop(net.parameters(), lr=0.00025, alpha=0.95, eps=0.01, centered=True))
    rp = bench.synthetic_ring(rl, rl.PrioritizedReplay if per else rl.UniformReplay, dev, seed=0)
    lr = GraphedDQNLearner(net, tgt, opt, rp, kind="dqn", double_q=double_q, gradient_clip=5.0, feeds_per_update=4,
                           compute_dtype=torch.bfloat16, target_sync_every=0, prefetch=False)
    return lr


def eager(lr, n, sync=True):
    out = []
    for _ in range(n):
        lr._main(), lr._opt()
        if sync:
            torch.cuda.synchronize()
        out.append(lr.loss.clone())
    torch.cuda.synchronize()
    return [float(x) for x in out]


def graph(lr, n):
    lr.capture(warmup=3)
    out = []
    for _ in range(n):
        lr.update()
        out.append(lr.loss.clone())
    torch.cuda.synchronize()
    return [float(x) for x in out]


for per in (False, True):
    for dueling in (False, True):
        for dq in (False, True):
            e1 = eager(make(per, dueling, dq), 8)
            e2 = eager(make(per, dueling, dq), 8, sync=False)
            g = graph(make(per, dueling, dq), 5)
            d_ee = np.max(np.abs(np.array(e1) - np.array(e2)) / np.abs(e1))
            d_eg = np.max(np.abs(np.array(e1[3:]) - np.array(g)) / np.abs(e1[3:]))
            print("per=%d dueling=%d double=%d | eager(sync) vs eager(async) %.2e | eager vs graph %.2e" % (per, dueling, dq, d_ee, d_eg))
            if d_eg > 1e-2:
                print("   eager:", ["%.5f" % v for v in e1[3:]])
                print("   graph:", ["%.5f" % v for v in g])
