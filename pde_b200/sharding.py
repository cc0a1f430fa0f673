"""Parameter-set sharding across the GPUs of one node.

Slices are independent and a parameter set's loss needs only its own maturities, so the P
parameter sets are partitioned contiguously across ranks, the (small) option surface is
replicated, and the only collective is an all-gather of per-set results (losses: 8 B/set;
normal-equation blocks: 176 B/set).  One process per GPU, ``torch.distributed`` with the NCCL
backend over NVLink; the same code runs on ``gloo`` for the CPU tests of the host logic.
There is no data-path collective inside the pricing kernels.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(n_sets: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition: the first ``n_sets % world`` ranks get one extra set."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, extra = divmod(n_sets, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_map(fn: Callable, X, group=None):
    """Evaluate ``fn`` on this rank's contiguous shard of the rows of ``X`` ([P, 5], identical on
    every rank) and all-gather the per-row results, returning the full [P, ...] tensor on every
    rank.  ``fn(X_local)`` must return a tensor whose first dimension is ``len(X_local)``.
    Without an initialised process group this is just ``fn(X)``."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return fn(X)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    P = X.shape[0]
    lo, hi = shard_bounds(P, rank, world)
    local = fn(X[lo:hi])
    if P % world == 0:
        # even partition (the usual case: populations are powers of two): gather straight into the result,
        # no padding, no slicing, no concatenation
        out = torch.empty((P,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    width = (P + world - 1) // world  # equal-size buckets for all_gather_into_tensor
    tail = tuple(local.shape[1:])
    padded = torch.zeros((width,) + tail, dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    gathered = torch.empty((world * width,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = []
    for rk in range(world):
        a, b = shard_bounds(P, rk, world)
        parts.append(gathered[rk * width: rk * width + (b - a)])
    return torch.cat(parts, dim=0)


class ShardedPricer:
    """A :class:`pde_b200.BatchPricer` on this rank's GPU plus the all-gather of its results."""

    def __init__(self, pricer, group=None):
        self.pricer = pricer
        self.group = group

    def objective(self, X):
        return sharded_map(self.pricer.objective, X, self.group)

    def normal_equations(self, X):
        return sharded_map(self.pricer.normal_equations, X, self.group)

    def price(self, X):
        return sharded_map(self.pricer.price, X, self.group)


def init_from_env(backend: Optional[str] = None):
    """Initialise torch.distributed from torchrun's environment (RANK, WORLD_SIZE, MASTER_*) and
    bind this process to cuda:LOCAL_RANK.  Returns (rank, world, local_rank)."""
    import os

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank,
                                world_size=world)
    return rank, world, local_rank
