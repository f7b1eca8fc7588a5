"""Batch-sharded sampling across the GPUs of one box.

Samples are independent (GroupNorm and attention are per-sample, the schedule is shared constants), so the path
shards by sample with NO collective inside the denoising loop; a single all-gather of the decoded range images
(NCCL over NVLink on GPUs, gloo in CPU tests) assembles the global batch.  For bit-level reproducibility against a
single-GPU run, x_T and the per-step noise are drawn for the GLOBAL batch from one seeded generator and sliced.
The reference has no multi-GPU sampling code (every script pins CUDA_VISIBLE_DEVICES=0, README.md:177).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split: rank r takes samples [lo, hi).  Remainders go to the lowest ranks."""
    if global_batch < 0 or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad shard arguments")
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_noise(shape, seed: int, n_steps: int = 0, device="cpu"):
    """x_T (and optional per-step noise) for the GLOBAL batch from one generator; identical on every rank."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    x_T = torch.randn(shape, generator=g)
    noise = torch.randn((n_steps,) + tuple(shape), generator=g) if n_steps > 0 else None
    return x_T.to(device), (noise.to(device) if noise is not None else None)


def local_slice(t: Optional[torch.Tensor], rank: int, world_size: int, batch_dim: int = 0):
    if t is None:
        return None
    lo, hi = shard_range(t.shape[batch_dim], rank, world_size)
    return t.narrow(batch_dim, lo, hi - lo).contiguous()


def all_gather_batch(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """The path's only collective: gather (B_local, ...) tensors into (global_batch, ...), in rank order.
    Handles ragged shards by padding to the largest shard."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < max_n:
        pad = torch.cat([local, local.new_zeros((max_n - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * max_n,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * max_n: r * max_n + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, 0)
