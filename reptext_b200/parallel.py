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


# ---------------------------------------------------------------------------------------------------------
# Sequence-parallel mode (BASELINE.json configs[4]: one 1536x1536 sample, 9216 image + 512 text tokens, on 8 GPUs)
# ---------------------------------------------------------------------------------------------------------
# Tokens are sharded in contiguous row blocks (rank r owns rows [r*n/P, (r+1)*n/P) of the image tokens and of the text
# tokens); the attention of every block runs head-sharded over the whole sequence.  The head <-> token exchanges are
# peer stores issued by the QKV-GEMM and attention epilogues (include/reptext_rt.h, rt_sp_group) - NCCL only gathers
# the final latents.

def shard_tokens(t: torch.Tensor, rank: int, world: int, dim: int = 1) -> torch.Tensor:
    """This rank's contiguous block of the token dimension."""
    n = t.shape[dim]
    if n % world:
        raise ValueError(f"{n} tokens cannot be split evenly over {world} ranks")
    step = n // world
    return t.narrow(dim, rank * step, step).contiguous()


def gather_tokens(local: torch.Tensor, group=None, dim: int = 1) -> torch.Tensor:
    """Inverse of :func:`shard_tokens` on every rank (all-gather; NCCL on GPUs, gloo in the CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=dim)


class _SpBase:
    world: int
    workspace_bytes: int = 0

    def struct(self, rank=None):
        raise NotImplementedError

    def ensure_workspace(self, nbytes: int) -> None:
        raise NotImplementedError


class LockstepGroup(_SpBase):
    """All ``world`` ranks of a sequence-parallel group driven by ONE process on ONE GPU (the models'
    ``forward_lockstep``).  Every rank has its own workspace, the kernels scatter between them exactly as they do
    between GPUs; stream order stands in for the cross-GPU barriers.  Used to validate the exchange indexing where
    only one GPU is available."""

    def __init__(self, world: int, device="cuda"):
        from . import _lib as L
        if not (1 <= world <= L.SP_MAX_RANKS):
            raise ValueError(f"world must be 1..{L.SP_MAX_RANKS}")
        self.world = world
        self.device = torch.device(device)
        self._bufs: List[torch.Tensor] = []
        self.workspace_bytes = 0

    def ensure_workspace(self, nbytes: int) -> None:
        if nbytes <= self.workspace_bytes:
            return
        self._bufs = [torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
        self.workspace_bytes = nbytes

    def struct(self, rank=None):
        from . import _lib as L
        g = L.SpGroup()
        g.world, g.rank, g.lockstep = self.world, int(rank), 1
        for i, b in enumerate(self._bufs):
            g.peer_workspace[i] = (b.data_ptr() + 255) // 256 * 256
        return g


class SequenceParallelGroup(_SpBase):
    """One process per GPU.  Owns this rank's peer-mappable workspace (``rt_ipc_alloc``), the other ranks' mappings
    of theirs (``rt_ipc_open``; handles travel over ``torch.distributed``) and the barrier flag words."""

    FLAG_BYTES = 4096

    def __init__(self, group=None):
        from . import _lib as L
        if not dist.is_initialized():
            raise RuntimeError("SequenceParallelGroup needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if not (2 <= self.world <= L.SP_MAX_RANKS):
            raise ValueError(f"sequence-parallel world size must be 2..{L.SP_MAX_RANKS}")
        self._own = None
        self._peers: List[Optional[int]] = []
        self.workspace_bytes = 0
        self.ensure_workspace(1 << 20)          # the flag words exist from the start (barrier() works right away)

    def _release(self):
        from . import _lib as L
        lib = L.lib()
        for i, p in enumerate(self._peers):
            if p is not None and i != self.rank:
                lib.rt_ipc_close(p)
        self._peers = []
        if self._own is not None:
            lib.rt_ipc_free(self._own)
            self._own = None
        self.workspace_bytes = 0

    def ensure_workspace(self, nbytes: int) -> None:
        """COLLECTIVE when the buffer has to grow (every rank asks for the same size in the same order)."""
        import ctypes as C
        from . import _lib as L
        if nbytes <= self.workspace_bytes:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)          # nobody still uses the old mapping
        self._release()
        lib = L.lib()
        total = self.FLAG_BYTES + (nbytes + 255) // 256 * 256
        p = C.c_void_p()
        handle = C.create_string_buffer(64)
        L.check(lib.rt_ipc_alloc(total, C.byref(p), handle))
        self._own = p.value
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=self.group)
        self._peers = []
        for i, h in enumerate(handles):
            if i == self.rank:
                self._peers.append(self._own)
            else:
                q = C.c_void_p()
                L.check(lib.rt_ipc_open(h, C.byref(q)))
                self._peers.append(q.value)
        self.workspace_bytes = nbytes
        dist.barrier(group=self.group)          # every mapping exists before the first peer store

    def struct(self, rank=None):
        from . import _lib as L
        g = L.SpGroup()
        g.world, g.rank, g.lockstep = self.world, self.rank, 0
        for i, p in enumerate(self._peers):
            g.peer_flags[i] = p
            g.peer_workspace[i] = p + self.FLAG_BYTES
        return g

    def barrier(self) -> None:
        import ctypes as C
        from . import _lib as L
        L.check(L.lib().rt_sp_barrier(C.byref(self.struct()), L.stream_ptr()))

    def check(self) -> None:
        """Synchronise and raise if a barrier on this rank ever timed out (a peer died or fell out of step)."""
        import ctypes as C
        from . import _lib as L
        bad = C.c_int(0)
        L.check(L.lib().rt_sp_status(C.byref(self.struct()), L.stream_ptr(), C.byref(bad)))
        if bad.value:
            raise RuntimeError("sequence-parallel barrier timed out: a peer rank is not making progress "
                               "(the group stays aborted until every rank calls reset())")

    def reset(self) -> None:
        """COLLECTIVE.  Re-synchronise the group after a failed forward (a rank raised mid-forward, or a barrier timed
        out): every rank drains its stream, then zeroes its own flag block - epochs, epoch counter and the sticky
        abort word - between two host barriers, so that no peer store can land in a block that is being cleared."""
        import ctypes as C
        from . import _lib as L
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        L.check(L.lib().rt_sp_reset(C.byref(self.struct()), L.stream_ptr()))
        dist.barrier(group=self.group)

    def close(self) -> None:
        if self._own is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self._release()
