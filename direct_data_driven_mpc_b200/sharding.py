"""Multi-GPU scenario sharding (one process per GPU, torch.distributed).

Closed loops are independent given their controller, so the scenario
axis is cut into contiguous shards, one per rank; each rank rebuilds the
(tiny) controller plan locally and runs its shard with the GLOBAL scenario ids
as Philox streams, so results do not depend on the number of ranks.  The only
collective of a job is the final gather of trajectories / per-loop metrics
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` scenarios owned by `rank` (sizes differ by at most 1)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> Sequence[int]:
    return [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]


def gather_shards(local: torch.Tensor, total: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Concatenate per-rank shards (dim 0) into the full (total, ...) tensor.

    dst=None: every rank receives the result (all_gather); dst=r: only rank r does (gather).
    Shards may have unequal sizes (total not divisible by the world size): they are padded to
    the largest shard for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    mx = max(sizes)
    if dst is None and min(sizes) == mx and local.is_cuda and dist.get_backend(group) == "nccl":
        # equal shards on NCCL: gather straight into the concatenated result (no list of parts, no second copy)
        local = local.contiguous()
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    if dst is None:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
    else:
        parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local, parts, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def run_sharded_closed_loops(controller_set, plant, batch: Dict, n_steps: int, noise_seed: int = 0,
                             noise_eps: Optional[float] = None, group=None, gather: str = "all",
                             tol: float = 1e-8, max_iter: int = 2000):
    """Run the closed loops of `batch` (host arrays x0, u_past0, y_past0, u_s, y_s of the WHOLE job;
    optional w) sharded over the ranks of `group`, then gather.

    gather: "all" (every rank gets the full trajectories), "root" (rank 0 only) or "none"."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = batch["x0"].shape[0]
    lo, hi = shard_bounds(total, world, rank)
    sl = lambda k: batch[k][lo:hi]
    w = batch.get("w")
    u, y, status, iters = controller_set.closed_loop(
        plant, sl("x0"), sl("u_past0"), sl("y_past0"), sl("u_s"), sl("y_s"), n_steps,
        w=None if w is None else w[lo:hi], noise_seed=noise_seed, scenario_id0=lo, noise_eps=noise_eps,
        tol=tol, max_iter=max_iter)
    if gather == "none" or world == 1:
        return u, y, status, iters
    dst = None if gather == "all" else 0
    return tuple(gather_shards(t, total, group, dst) for t in (u, y, status, iters))
