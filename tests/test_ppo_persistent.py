"""The persistent PPO minibatch kernel (deeprl_b200/csrc/ppo_persistent.cu: every minibatch update of an iteration in one
launch; reference PPO_agent.py:68-99, non-shared representation) against the oracle's restatement of the same loop
(oracle/agents.py ppo_update, pinned against the real reference by tests/test_oracle_golden.py::test_ppo_step_trajectory).

CPU: the kernel's phase functions (csrc/ppo_phases.h + ppo_sequence.inc) are compiled for the host by tests/host_emul and run
with the block's threads in sequence -- forwards and backwards, which would expose a dependence on intra-phase order, i.e. a race.
GPU: the CUDA build of the same source through the C ABI (``b2rl_ppo_minibatch_updates``) and through ``PPOAgent``.

Tolerance: fp32 sums in another order than MKL / cuBLAS, through up to 48 Adam steps: parameters to 2e-5 absolute."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import agents, nets  # noqa: E402

A_KEYS = ["actor_body.layers.0.weight", "actor_body.layers.0.bias", "actor_body.layers.1.weight", "actor_body.layers.1.bias",
          "fc_action.weight", "fc_action.bias", "std"]
C_KEYS = ["critic_body.layers.0.weight", "critic_body.layers.0.bias", "critic_body.layers.1.weight", "critic_body.layers.1.bias",
          "fc_critic.weight", "fc_critic.bias"]


def make_problem(D, A, H1, H2, rows, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale
    sd = {"actor_body.layers.0.weight": r(H1, D, scale=D ** -0.5), "actor_body.layers.0.bias": r(H1, scale=0.1),
          "actor_body.layers.1.weight": r(H2, H1, scale=H1 ** -0.5), "actor_body.layers.1.bias": r(H2, scale=0.1),
          "fc_action.weight": r(A, H2, scale=0.3 * H2 ** -0.5), "fc_action.bias": r(A, scale=0.05), "std": r(A, scale=0.3),
          "critic_body.layers.0.weight": r(H1, D, scale=D ** -0.5), "critic_body.layers.0.bias": r(H1, scale=0.1),
          "critic_body.layers.1.weight": r(H2, H1, scale=H1 ** -0.5), "critic_body.layers.1.bias": r(H2, scale=0.1),
          "fc_critic.weight": r(1, H2, scale=H2 ** -0.5), "fc_critic.bias": r(1, scale=0.1)}
    states = r(rows, D)
    with torch.no_grad():
        out = nets.gaussian_actor_critic(sd, states, torch.zeros(rows, A))
        actions = out["mean"] + torch.nn.functional.softplus(sd["std"]) * r(rows, A)
        log_pi_old = nets.gaussian_actor_critic(sd, states, actions)["log_pi_a"] + r(rows, 1, scale=0.05)
    ret, adv = r(rows, 1), r(rows, 1)
    return sd, states, actions, log_pi_old, ret, adv


def arena(sd, keys):
    """FlatOptimizer's layout (ops.py): every tensor starts on a multiple of 4 elements."""
    offs, n = [], 0
    for k in keys:
        offs.append(n)
        n += (sd[k].numel() + 3) // 4 * 4
    flat = np.zeros(n, np.float32)
    for k, o in zip(keys, offs):
        flat[o:o + sd[k].numel()] = sd[k].detach().numpy().ravel()
    return flat, np.asarray(offs, np.int32)


def oracle_run(sd0, states, actions, log_pi_old, ret, adv, epochs, mb, clip, ent_w, target_kl, a_lr, c_lr, seed):
    sd = agents.leafify(sd0)
    a_opt = torch.optim.Adam([sd[k] for k in A_KEYS], a_lr)
    c_opt = torch.optim.Adam([sd[k] for k in C_KEYS], c_lr)
    np.random.seed(seed)
    agents.ppo_update(sd, A_KEYS, C_KEYS, a_opt, c_opt, states, actions, log_pi_old, ret, adv.clone(), epochs, mb, clip, ent_w,
                      target_kl)
    a_steps = int(a_opt.state[sd[A_KEYS[0]]]["step"]) if a_opt.state else 0
    return sd, a_steps, a_opt, c_opt


def batches_for(rows, epochs, mb, seed):
    np.random.seed(seed)
    return np.stack([np.asarray(b, np.int64) for _ in range(epochs) for b in agents.random_sample(np.arange(rows), mb)])


F, I64, I32 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)


def fp(x):
    return x.ctypes.data_as(F)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ppo_emul") / "ppo_emul.so")
    subprocess.run(["g++", "-O2", "-fno-strict-aliasing", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "host_emul", "ppo_emul.cpp")],
                   check=True)
    return ctypes.CDLL(out)


CASES = [dict(D=17, A=6, H1=64, H2=64, rows=512, mb=64, epochs=3, target_kl=0.01, a_lr=3e-4),       # examples.py:496-522 shapes
         dict(D=17, A=6, H1=64, H2=64, rows=256, mb=64, epochs=4, target_kl=2e-4, a_lr=3e-3),       # the KL gate closes on the way
         dict(D=11, A=3, H1=32, H2=48, rows=160, mb=32, epochs=2, target_kl=0.01, a_lr=3e-4),       # other shapes, ragged tiles
         dict(D=4, A=1, H1=20, H2=12, rows=96, mb=8, epochs=2, target_kl=0.01, a_lr=3e-4)]


def run_emul(lib, fn, c, problem, threads):
    sd0, states, actions, log_pi_old, ret, adv = problem
    adv_n = ((adv - adv.mean()) / adv.std()).numpy().ravel().copy()                  # PPO_agent.py:66 (the caller's, as in the agent)
    a_flat, a_off = arena(sd0, A_KEYS)
    c_flat, c_off = arena(sd0, C_KEYS)
    a_m, a_v, c_m, c_v = (np.zeros_like(x) for x in (a_flat, a_flat, c_flat, c_flat))
    a_step, c_step = np.zeros(1, np.int64), np.zeros(1, np.int64)
    perm = batches_for(c["rows"], c["epochs"], c["mb"], seed=77)
    stats = np.zeros(4, np.float32)
    st, ac, lp, rt = (np.ascontiguousarray(t.numpy(), np.float32) for t in (states, actions, log_pi_old.ravel(), ret.ravel()))
    getattr(lib, fn)(fp(st), fp(ac), fp(lp), fp(rt), fp(adv_n), c["D"], c["A"], c["H1"], c["H2"], c["mb"],
                     perm.ctypes.data_as(I64), perm.shape[0], fp(a_flat), fp(a_m), fp(a_v), a_step.ctypes.data_as(I64),
                     a_off.ctypes.data_as(I32), fp(c_flat), fp(c_m), fp(c_v), c_step.ctypes.data_as(I64), c_off.ctypes.data_as(I32),
                     ctypes.c_float(c["a_lr"]), ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-8),
                     ctypes.c_float(1e-3), ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(0.2),
                     ctypes.c_float(0.01), ctypes.c_float(1.5 * c["target_kl"]), fp(stats), threads)
    return dict(a_flat=a_flat, a_off=a_off, c_flat=c_flat, c_off=c_off, a_m=a_m, a_v=a_v, c_m=c_m, c_v=c_v, a_step=int(a_step[0]),
                c_step=int(c_step[0]), stats=stats, n_batches=perm.shape[0])


def check_against_oracle(got, c, problem):
    sd0, states, actions, log_pi_old, ret, adv = problem
    sd, a_steps, a_opt, c_opt = oracle_run(sd0, states, actions, log_pi_old, ret, adv, c["epochs"], c["mb"], 0.2, 0.01,
                                           c["target_kl"], c["a_lr"], 1e-3, seed=77)
    assert got["c_step"] == got["n_batches"] and got["a_step"] == a_steps == int(got["stats"][3])
    for keys, flat, off, opt, m, v in ((A_KEYS, got["a_flat"], got["a_off"], a_opt, got["a_m"], got["a_v"]),
                                       (C_KEYS, got["c_flat"], got["c_off"], c_opt, got["c_m"], got["c_v"])):
        for k, o in zip(keys, off):
            want = sd[k].detach().numpy().ravel()
            np.testing.assert_allclose(flat[o:o + want.size], want, rtol=0, atol=2e-5, err_msg=k)
            moved = np.abs(want - sd0[k].numpy().ravel()).max()
            assert moved > 1e-5 or a_steps == 0, k               # the comparison is not between two untouched copies
            if opt.state:
                stt = opt.state[sd[k]]
                np.testing.assert_allclose(m[o:o + want.size], stt["exp_avg"].numpy().ravel(), rtol=1e-3, atol=1e-7, err_msg=k)
                np.testing.assert_allclose(v[o:o + want.size], stt["exp_avg_sq"].numpy().ravel(), rtol=2e-3, atol=1e-10, err_msg=k)
    return a_steps


@pytest.mark.parametrize("case", range(len(CASES)))
def test_phase_functions_match_oracle_on_the_host(emul, case):
    c = CASES[case]
    problem = make_problem(c["D"], c["A"], c["H1"], c["H2"], c["rows"], seed=case)
    got = run_emul(emul, "ppo_emul_minibatch_updates", c, problem, 512)
    a_steps = check_against_oracle(got, c, problem)
    assert a_steps > 0
    if case in (0, 1):
        assert a_steps < got["n_batches"]                        # the gate was open for some minibatches and closed for others
    # no dependence on thread order inside a phase, nor on the number of threads of the block
    rev = run_emul(emul, "ppo_emul_minibatch_updates_reversed", c, problem, 512)
    few = run_emul(emul, "ppo_emul_minibatch_updates", c, problem, 64)
    for k in ("a_flat", "c_flat", "a_m", "a_v", "c_m", "c_v"):
        assert np.array_equal(got[k], rev[k]) and np.array_equal(got[k], few[k]), k
    assert np.array_equal(got["stats"], rev["stats"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
def test_cuda_kernel_matches_oracle(case):
    """The CUDA build of the same phases through the C ABI (one launch for all minibatches of the case)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from deeprl_b200 import _lib
    c = CASES[case]
    problem = make_problem(c["D"], c["A"], c["H1"], c["H2"], c["rows"], seed=case)
    sd0, states, actions, log_pi_old, ret, adv = problem
    adv_n = (adv - adv.mean()) / adv.std()
    a_flat, a_off = arena(sd0, A_KEYS)
    c_flat, c_off = arena(sd0, C_KEYS)
    perm = batches_for(c["rows"], c["epochs"], c["mb"], seed=77)
    dev = torch.device("cuda", 0)
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x)).to(dev)
    t = dict(state=cu(states.numpy()), action=cu(actions.numpy()), lp=cu(log_pi_old.numpy().ravel()), ret=cu(ret.numpy().ravel()),
             adv=cu(adv_n.numpy().ravel()), perm=cu(perm), a_flat=cu(a_flat), c_flat=cu(c_flat))
    for k in ("a", "c"):
        t[k + "_m"], t[k + "_v"] = torch.zeros_like(t[k + "_flat"]), torch.zeros_like(t[k + "_flat"])
        t[k + "_step"] = torch.zeros(1, dtype=torch.int64, device=dev)
    stats = torch.zeros(4, device=dev)
    a_off_t, c_off_t = torch.from_numpy(a_off), torch.from_numpy(c_off)
    _lib.call("b2rl_ppo_minibatch_updates", _lib.ptr(t["state"]), _lib.ptr(t["action"]), _lib.ptr(t["lp"]), _lib.ptr(t["ret"]),
              _lib.ptr(t["adv"]), c["D"], c["A"], c["H1"], c["H2"], c["mb"], _lib.ptr(t["perm"]), perm.shape[0],
              _lib.ptr(t["a_flat"]), _lib.ptr(t["a_m"]), _lib.ptr(t["a_v"]), _lib.ptr(t["a_step"]), _lib.ptr(a_off_t),
              _lib.ptr(t["c_flat"]), _lib.ptr(t["c_m"]), _lib.ptr(t["c_v"]), _lib.ptr(t["c_step"]), _lib.ptr(c_off_t),
              c["a_lr"], 0.9, 0.999, 1e-8, 1e-3, 0.9, 0.999, 1e-8, 0.2, 0.01, 1.5 * c["target_kl"], _lib.ptr(stats), _lib.stream())
    torch.cuda.synchronize()
    got = dict(a_flat=t["a_flat"].cpu().numpy(), a_off=a_off, c_flat=t["c_flat"].cpu().numpy(), c_off=c_off,
               a_m=t["a_m"].cpu().numpy(), a_v=t["a_v"].cpu().numpy(), c_m=t["c_m"].cpu().numpy(), c_v=t["c_v"].cpu().numpy(),
               a_step=int(t["a_step"]), c_step=int(t["c_step"]), stats=stats.cpu().numpy(), n_batches=perm.shape[0])
    check_against_oracle(got, c, problem)


def test_shared_memory_budget_of_the_example_sizes():
    """The examples' sizes (examples.py:496-522) fit one SM; the learner falls back to the graph form when a network does not."""
    from deeprl_b200 import _lib
    L = _lib.lib()
    assert 0 < L.b2rl_ppo_minibatch_smem_bytes(17, 6, 64, 64, 64) <= 227 * 1024
    assert L.b2rl_ppo_minibatch_smem_bytes(64, 32, 128, 128, 128) > 227 * 1024
