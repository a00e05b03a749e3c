"""One learner per GPU: the only exchange step of the path is the gradient all-reduce.

Replay, sum tree and loss kernels are rank-local (each rank owns an independent ring + tree, its own envs and its
own Philox stream seeded with the rank); parameters start identical (rank-0 broadcast) and stay identical because
every rank applies the same clipped update to the same summed gradient.  The collective is NCCL on GPUs
(``backend="nccl"``); the same host logic runs over ``gloo`` on CPU tensors for the world_size-2 tests.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = dict(device_id=torch.device("cuda", local)) if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return world, rank, local


def broadcast_parameters(flat, src=0):
    """Make the flat parameter arena identical on every rank."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src)
    return flat


def allreduce_gradients(flat_grad, async_op=False):
    """SUM the flat gradient arena (or a slice of it) over ranks; the 1/world factor is folded into the clip kernel's
    grad_scale.  ``async_op``: returns the work handle (the learner all-reduces fc4's slice beside the convolution backward and
    waits for it before the optimizer kernels; inside a CUDA-graph capture the handle's wait() becomes a graph dependency)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


def leave():
    """End of a multi-rank run whose collectives were captured in CUDA graphs: the process group's destructor can hang then
    (measured), so the ranks synchronise, flush and exit directly."""
    import sys
    if dist.is_initialized() and dist.get_world_size() > 1:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        sys.stdout.flush(), sys.stderr.flush()
        os._exit(0)


def grad_scale():
    return 1.0 / dist.get_world_size() if dist.is_initialized() else 1.0


def max_over_ranks(value, device):
    """Device-side max over ranks of a host scalar (timings are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def rank_seed(base_seed):
    """Rank-local RNG stream for the replay shard / envs."""
    _, rank, _ = env_world()
    return int(base_seed) * 1000003 + rank
