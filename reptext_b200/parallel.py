"""Multi-GPU plumbing for the RepText path: one process per GPU, ``torch.distributed``.

The denoising step shards over independent (prompt, seed) samples: rank ``r`` takes samples ``r::world``, weights are
replicated, nothing is exchanged per step.  The ONE collective is the gather of the output latents (512 KB per
sample at 1024x1024) after the loop - NCCL on GPUs, gloo in the CPU tests.  The reference has no distributed mode
at all ("replicas by hand"), so there is nothing upstream to mirror here.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank) from the torchrun environment; initialises the default group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_indices(n_samples: int, rank: int, world: int) -> List[int]:
    """Round-robin: sample i runs on rank i % world (keeps the per-rank load within one sample)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return list(range(rank, n_samples, world))


def gather_samples(local: torch.Tensor, n_samples: int, group=None) -> torch.Tensor:
    """All-gather per-rank results ``[n_local, ...]`` (rank r holds samples r::world) into ``[n_samples, ...]`` in
    sample order, on every rank.  Ranks with one sample fewer are padded for the collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local.shape[0] != n_samples:
            raise ValueError("single process must hold every sample")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_max = (n_samples + world - 1) // world
    want = len(shard_indices(n_samples, rank, world))
    if local.shape[0] != want:
        raise ValueError(f"rank {rank} holds {local.shape[0]} samples, expected {want}")
    padded = local
    if want < n_max:
        pad = torch.zeros((n_max - want,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    out = torch.empty((n_samples,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_samples, r, world)
        out[idx] = parts[r][: len(idx)]
    return out


def run_sharded(samples: Sequence, fn: Callable[[int, object], torch.Tensor], group=None) -> torch.Tensor:
    """Run ``fn(index, sample) -> Tensor`` on this rank's shard of ``samples`` and gather all results in order.
    ``fn`` is typically one pipeline call with ``output_type="latent"`` (BASELINE.json configs[2])."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_indices(len(samples), rank, world)
    outs = [fn(i, samples[i]) for i in mine]
    if not outs:
        raise ValueError("fewer samples than ranks")
    return gather_samples(torch.stack(outs, dim=0), len(samples), group)
