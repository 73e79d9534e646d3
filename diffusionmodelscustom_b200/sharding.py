"""Sample sharding across the GPUs of one box.

Every sample's trajectory depends only on its own (x_T, conditioning, z) and the replicated weights, so the batch is cut
into contiguous per-rank blocks with NO data-path collective; the in-kernel Philox noise is keyed by the GLOBAL sample
index (``sample_offset``), which makes the result independent of the number of ranks.  The only collective is one
``all_gather`` of the final fields (NCCL over NVLink on GPUs; gloo in the CPU tests of this plumbing).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank `rank`; the first n_total % world ranks get one extra sample."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t, n_total: int, world: int, rank: int):
    if t is None:
        return None
    lo, hi = shard_range(n_total, world, rank)
    return t[lo:hi]


def gather_fields(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather of ragged per-rank blocks [n_r, C, H, W] -> [n_total, C, H, W] on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((nmax,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def sample_sharded(diffusion, model, x_T, y=None, cond_img=None, lsm_cond=None, topo_cond=None, seed: int = 0, group=None,
                   gather: bool = True):
    """Each rank samples its block of the global batch (inputs are the GLOBAL tensors on the rank's device) and the final
    fields are gathered.  Equivalent to ``diffusion.sample`` on one device with the same seed."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = x_T.shape[0]
    lo, hi = shard_range(n, world, rank)
    sl = lambda t: None if t is None else t[lo:hi]
    x0 = diffusion.sample(x_T[lo:hi], model, sl(y), sl(cond_img), sl(lsm_cond), sl(topo_cond), seed=seed, sample_offset=lo)
    return gather_fields(x0, n, group) if gather else x0
