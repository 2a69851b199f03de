"""Sample sharding of the physics layer over the GPUs of one node.

Every coarse-grained-model solve (bottleneck/ROM.py:83, batched over B) and every virtual-observable
residual (bottleneck/VirtualObservables.py:895, loop over data points) is independent, so the batch is cut
into contiguous slices, one per rank (one process per GPU, torch.distributed); the mesh constants are
replicated in each rank's plans and the data path needs NO collective.  The only exchanges are
  * the sum over data points in the VO precision hyper-update (VirtualObservables.py:983-992): m doubles,
  * timing / result collection for benchmarks and tests.
Works with any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous slice [lo, hi) of a batch of B samples owned by ``rank``; sizes differ by at most one."""
    B, rank, world = int(B), int(rank), int(world)
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, B, rank, world, shared=("V",)):
    """Slices every [B, ...] tensor of a dict to this rank's shard.  Entries named in ``shared`` (the weighting
    matrix V, fields shared by the whole batch) and entries whose leading size is not B pass through."""
    lo, hi = shard_range(B, rank, world)
    out = {}
    for k, t in tensors.items():
        batched = t is not None and k not in shared and t.dim() >= 1 and t.shape[0] == B
        out[k] = t[lo:hi] if batched else t
    return out


def gather_batch(local, B, group=None):
    """Reassembles a [B, ...] tensor from the ranks' shards (all ranks get it).  Shards may differ by one row:
    they are padded to the largest before the all_gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rows = max(shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world))
    pad = local.new_zeros((rows,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: shard_range(B, r, world)[1] - shard_range(B, r, world)[0]] for r in range(world)], 0)


def allreduce_sum_(t, group=None):
    """In-place sum over ranks (the data-point sum of the VO precision hyper-update; scalar diagnostics)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (device-timed milliseconds in bench.py)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def vo_precision_sums(residual, Gamma_sq_vars=None, group=None):
    """beta-sums of VirtualObservablesEnsemble.update_vo_precision (VirtualObservables.py:983-992) for a
    sharded VO set: sum_n r_n^2 (+ sum_n Gamma_n^2 vars_n) over ALL data points = local sum + all-reduce."""
    s = (residual.double() ** 2).sum(dim=0)
    if Gamma_sq_vars is not None:
        s = s + Gamma_sq_vars.double().sum(dim=0)
    return allreduce_sum_(s, group)
