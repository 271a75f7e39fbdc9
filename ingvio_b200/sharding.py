"""Sharding of independent sequences over ranks (one process per GPU) -- SURVEY.md §8e.

The filter is per-sequence and needs no data-path collective: sequences are block-distributed, every
rank generates / receives only its shard, and the only collectives are (i) a MAX reduction of the timed
region and (ii) a gather of per-sequence read-outs (pose 12 + trace(P) + flags), a few KB per frame.
Works with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_range(total, world, rank):
    """Block distribution: ranks [0, total % world) get one extra sequence."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, dist=None, device="cpu"):
    """Max of a scalar over all ranks (the timed region of bench.py)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_readouts(local, total, dist=None, device="cpu"):
    """All-gather of per-sequence read-outs (B_local x K float64) into the global (total x K) array,
    in sequence order. Shards may differ in size by one."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    K = local.shape[1]
    sizes = [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad, K), dtype=torch.float64, device=device)
    buf[:local.shape[0]] = torch.from_numpy(local).to(device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)], 0)


def gather_readouts_device(local, total, dist=None):
    """All-gather of per-sequence read-outs that stays on the device: `local` is a (B_local x K) float64 torch tensor on
    this rank's GPU (or a CPU tensor under gloo), the result is the (total x K) tensor in sequence order on every rank.
    Shards may differ in size by one (block distribution of shard_range); one collective per call, a few KB."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]
    pad, K = max(sizes), local.shape[1]
    if local.shape[0] == pad:
        buf = local.contiguous()
    else:
        buf = torch.zeros((pad, K), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
    out = torch.empty((world * pad, K), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    if all(s == pad for s in sizes):
        return out
    return torch.cat([out[r * pad:r * pad + s] for r, s in enumerate(sizes)], 0)
