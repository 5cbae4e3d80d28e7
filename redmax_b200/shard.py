"""Batch sharding across GPUs / ranks.

Rollouts are independent (nothing couples them anywhere in simLoop, driverRedMaxBDF1.m:57-91), so the batch is cut into
contiguous shards, one per GPU, and **no collective runs on the data path**.  Two collectives exist off the data path:

  * `gather_trajectories`  -- all-gather of the q(t) shards the north star asks for ("NCCL over NVLink only to gather
                              final trajectories");
  * `reduce_objective`     -- sum of (P, dP/dp) over the batch when all rollouts share one parameter vector (the batched
                              form of driverRedMaxAdjointBDF1.m:39 `taskObjective`).

Both take torch tensors and work on any `torch.distributed` backend (NCCL on the GPU box, gloo in the CPU tests); the
partition rule is the one `rmx_rollout` uses inside one process for `opts.ngpus` devices (rmx_api.cu: b0 = B*g/G).
Nothing here computes dynamics.
"""
from __future__ import annotations


def shard_bounds(B, world, rank):
    """Contiguous shard [lo, hi) of a batch of B rollouts owned by `rank` of `world` (same rule as rmx_rollout)."""
    B, world, rank = int(B), int(world), int(rank)
    if world < 1 or not (0 <= rank < world) or B < 0:
        raise ValueError('shard_bounds: bad (B, world, rank) = (%d, %d, %d)' % (B, world, rank))
    return B * rank // world, B * (rank + 1) // world


def shard_sizes(B, world):
    return [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]


def take_shard(array, world, rank):
    """Rows [lo, hi) of a batch-leading array (numpy or torch)."""
    lo, hi = shard_bounds(len(array), world, rank)
    return array[lo:hi]


def gather_trajectories(local, B=None, group=None, out=None):
    """All-gather batch-leading trajectory shards (B_local x nsteps x nr) into the full (B x nsteps x nr) tensor on every
    rank.  Equal shards use one `all_gather_into_tensor` (a single NCCL all-gather over NVLink on the GPU box); ragged
    shards (B not a multiple of the world size) are padded to the largest shard and trimmed after the gather.  `out`: optional
    preallocated (B x nsteps x nr) tensor for the equal-shard case (keeps the allocation out of a timed region)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if B is None:
        nb = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        dist.all_reduce(nb, group=group)
        B = int(nb.item())
    sizes = shard_sizes(B, world)
    rank = dist.get_rank(group)
    if local.shape[0] != sizes[rank]:
        raise ValueError('gather_trajectories: rank %d holds %d rollouts, partition says %d' % (rank, local.shape[0], sizes[rank]))
    local = local.contiguous()
    if len(set(sizes)) == 1:
        if out is None:
            out = local.new_empty((B,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    m = max(sizes)
    padded = local.new_zeros((m,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    buf = local.new_empty((world * m,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * m:r * m + sizes[r]] for r in range(world)], dim=0)


def reduce_objective(P_local, dPdp_local, group=None):
    """Sum of the per-rollout objectives and gradients over the whole (sharded) batch: (sum_b P_b, sum_b dP/dp_b).
    One all-reduce of np + 1 doubles."""
    import torch
    import torch.distributed as dist
    acc = torch.cat([P_local.sum().reshape(1), dPdp_local.sum(dim=0).reshape(-1)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc[0], acc[1:]
