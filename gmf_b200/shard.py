"""Pair sharding across ranks: fragment pairs are independent, so each rank owns a contiguous slice of the
batch and the only cross-rank step is the final host gather of [B,4,4] poses (+ labels).  No collective on the data path.
The training steps (trainer.py, dgr_head.py) are data-parallel replicas with ONE collective: `exchange_gradients`."""
from __future__ import annotations

from typing import List, Tuple

import torch


def shard_range(num_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (num_pairs % world) ranks get one extra pair."""
    base, rem = divmod(num_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_poses(local_trans: torch.Tensor, num_pairs: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Host gather of the per-rank [b,4,4] results into [num_pairs,4,4] on every rank (CPU tensors; gloo or nccl group)."""
    import torch.distributed as dist
    if world == 1:
        return local_trans.cpu()
    sizes = [shard_range(num_pairs, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    # NCCL moves device memory only: host-resident results (the C ABI's host entry point) take a [mx,4,4] staging tensor on the GPU
    dev = local_trans.device
    if dist.get_backend(group) == "nccl" and dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    buf = torch.zeros(mx, 4, 4, dtype=torch.float32, device=dev)
    buf[: local_trans.shape[0]] = local_trans
    outs: List[torch.Tensor] = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)]).cpu()


def exchange_gradients(flat_grads: torch.Tensor, group=None, check_finite: bool = False) -> Tuple[int, bool]:
    """The one collective of a data-parallel training step: all-reduce (sum) of the flat gradient over the process group (NCCL on the GPU
    boxes, gloo in the CPU tests).  Returns (world, ok): the optimiser kernels scale by 1 / world (mean gradient); with `check_finite`, ok is the
    reference's guard (libs/trainer.py:161-166) evaluated on the REDUCED gradient, so every rank takes the same skip decision."""
    import torch.distributed as dist
    world = 1
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    ok = bool(torch.isfinite(flat_grads).all()) if check_finite else True
    return world, ok
