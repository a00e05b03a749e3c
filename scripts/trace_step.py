"""Timeline of ONE captured DQN update from timing events recorded inside the graph (learner.StepTrace): the real overlap of
the branches.  Usage: python scripts/trace_step.py [--workload dqn] [--replay async|sync] [--replays 50]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deeprl_b200 as rl  # noqa: E402
from deeprl_b200.learner import StepTrace  # noqa: E402
from deeprl_b200.network import nature_tc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="dqn")
ap.add_argument("--replay", default="async")
ap.add_argument("--replays", type=int, default=50)
ap.add_argument("--capacity", type=int, default=200_000)
a = ap.parse_args()
rl.select_device(0)
rl.Config.COMPUTE_DTYPE = torch.bfloat16
bench.CAP = a.capacity
lr = bench.build_learner(rl, a.workload, torch.device("cuda", 0), 0, 1, prefetch=(a.replay == "async"))
# warm up eagerly, then capture ONE graph (parity 0) with the trace on
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        lr._main(), lr._opt()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
tr = nature_tc.TRACE = StepTrace()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    lr._main(0 if lr.prefetch else None)
    lr._opt()
nature_tc.TRACE = None
acc = None
for i in range(a.replays):
    g.replay()
    torch.cuda.synchronize()
    t = np.array([x for _, x in tr.timeline()])
    if i >= 5:
        acc = t if acc is None else acc + t
acc /= (a.replays - 5)
names = [n for n, _ in tr.marks]
order = np.argsort(acc)
print("# us after 'start' (mean of %d replays), workload %s, replay %s" % (a.replays - 5, a.workload, a.replay))
for i in order:
    print("%9.1f  %s" % (acc[i], names[i]))
# total per replay, back to back
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("# back-to-back replay period (with the event nodes): %.1f us" % (e0.elapsed_time(e1) * 1e3 / 200))
