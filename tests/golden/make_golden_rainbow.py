"""Golden fixture for the Rainbow pieces: ``NoisyLinear`` (network_utils.py:31-83) and ``RainbowNet``
(network_heads.py:57-86) of the UNMODIFIED reference, imported through oracle/ref_shim.py in the build container.
Writes tests/golden/rainbow.npz (inputs, full state_dicts incl. the noise buffers, outputs in train and eval mode)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shim import import_reference  # noqa: E402

ref = import_reference()
torch.set_num_threads(1)
ref.select_device(-1)
out = {}
torch.manual_seed(7)
np.random.seed(7)
# ---- NoisyLinear: parameters, factorised noise, forward in both modes
nl = ref.NoisyLinear(24, 10)
nl.reset_noise()
x = torch.randn(5, 24)
for k, v in nl.state_dict().items():
    out["nl_sd_" + k] = v.numpy().copy()
out["nl_x"] = x.numpy()
nl.train()
out["nl_y_train"] = nl(x).detach().numpy()
nl.eval()
out["nl_y_eval"] = nl(x).detach().numpy()
# ---- RainbowNet over an FCBody, noisy and plain
for noisy in (True, False):
    tag = "rb%d_" % int(noisy)
    net = ref.RainbowNet(4, 11, ref.FCBody(6, hidden_units=(16,), noisy_linear=noisy), noisy)
    if noisy:
        net.reset_noise()
    s = np.random.randn(7, 6).astype(np.float32)
    net.train()
    o = net(s)
    for k, v in net.state_dict().items():
        out[tag + "sd_" + k] = v.numpy().copy()
    out[tag + "x"] = s
    out[tag + "prob"] = o["prob"].detach().numpy()
    out[tag + "log_prob"] = o["log_prob"].detach().numpy()
np.savez_compressed(os.path.join(HERE, "rainbow.npz"), **out)
print("rainbow.npz", len(out), "arrays")
