"""world_size-2 gloo test (CPU) of the multi-GPU host logic: rank-0 parameter broadcast, SUM all-reduce of
the flat gradient arena with the 1/world scale, identical clipped updates -> parameters stay bit-identical on both
ranks and equal to a single-process run on the concatenated (2 x batch) gradient mean."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from deeprl_b200 import parallel
    w, r, _ = parallel.init("gloo")
    assert (w, r) == (world, rank)
    torch.manual_seed(100 + rank)                           # different init per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    params = list(net.parameters())
    flat = torch.cat([p.detach().reshape(-1) for p in params]).clone()
    parallel.broadcast_parameters(flat)
    off = 0
    for p in params:
        p.data = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    opt = torch.optim.RMSprop(params, lr=1e-2, alpha=0.95, eps=0.01, centered=True)
    gen = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(5, 2 * 8, 6, generator=gen), torch.randn(5, 2 * 8, 3, generator=gen)
    for it in range(5):
        x, y = x_all[it, rank * 8:(rank + 1) * 8], y_all[it, rank * 8:(rank + 1) * 8]     # rank-local shard
        opt.zero_grad()
        (net(x) - y).pow(2).mul(0.5).mean().backward()
        g = torch.cat([p.grad.reshape(-1) for p in params])
        parallel.allreduce_gradients(g)
        g.mul_(parallel.grad_scale())
        off = 0
        for p in params:
            p.grad.copy_(g[off:off + p.numel()].view_as(p))
            off += p.numel()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
    t = parallel.max_over_ranks(rank + 1.0, "cpu")
    assert t == float(world)
    assert parallel.rank_seed(3) != parallel.rank_seed(4)
    out[rank] = flat.numpy().copy()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_keeps_parameters_identical():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert np.array_equal(a, b)                              # bit-identical replicas
    # single-process reference on the full batch: mean over 16 rows == mean of the two 8-row means
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    params = list(net.parameters())
    opt = torch.optim.RMSprop(params, lr=1e-2, alpha=0.95, eps=0.01, centered=True)
    gen = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(5, 16, 6, generator=gen), torch.randn(5, 16, 3, generator=gen)
    for it in range(5):
        opt.zero_grad()
        (net(x_all[it]) - y_all[it]).pow(2).mul(0.5).mean().backward()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in params]).numpy()
    np.testing.assert_allclose(a, ref, rtol=1e-5, atol=1e-6)


def _worker_async_and_leave(rank, world, port, out):
    """The learner's overlapped form of the exchange (learner.py `_allreduce_early` / `_allreduce`): the large slice of the arena
    is all-reduced asynchronously, the remainder afterwards, then the handle is waited for; and ``parallel.leave()`` ends the
    run with every rank at exit code 0 (it must return normally when there is nothing to tear down around: world 1)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from deeprl_b200 import parallel
    parallel.init("gloo")
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    early = parallel.allreduce_gradients(g[100:900], async_op=True)     # "fc4's slice"
    for lo, hi in ((0, 100), (900, 1000)):                               # "the late slices"
        parallel.allreduce_gradients(g[lo:hi])
    early.wait()
    want = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world))
    out[rank] = bool(torch.equal(g, want))
    parallel.leave()                                                     # barrier, flush, os._exit(0)
    raise AssertionError("leave() returned in a multi-rank run")


def test_sliced_async_allreduce_and_leave():
    from deeprl_b200 import parallel
    assert parallel.allreduce_gradients(torch.ones(4)) is None           # no process group: nothing to do, nothing to wait for
    parallel.leave()                                                     # single process: returns
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_async_and_leave, args=(world, port, out), nprocs=world, join=True)    # raises if a rank exits non-zero
    assert out[0] is True and out[1] is True
