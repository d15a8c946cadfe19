"""The N > 1 path on the CPU: world_size-2 gloo processes shard independent samples round-robin and gather the
output latents in sample order (the data path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reptext_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_samples, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    samples = [dict(seed=100 + i) for i in range(n_samples)]
    seen = []

    def denoise(i, s):   # stands in for one pipeline call: a deterministic function of the sample alone
        seen.append(i)
        g = torch.Generator().manual_seed(s["seed"])
        return torch.randn(16, 8, generator=g)

    out = parallel.run_sharded(samples, denoise)
    q.put((rank, seen, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [4, 5])
def test_two_ranks_shard_and_gather_in_sample_order(n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_samples, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([torch.randn(16, 8, generator=torch.Generator().manual_seed(100 + i)) for i in range(n_samples)])
    for rank, seen, out in res:
        assert seen == list(range(rank, n_samples, 2))
        assert torch.equal(out, want)


def test_shard_indices_and_single_process():
    assert parallel.shard_indices(64, 3, 8) == list(range(3, 64, 8))
    assert parallel.shard_indices(2, 1, 8) == [1] and parallel.shard_indices(2, 5, 8) == []
    with pytest.raises(ValueError):
        parallel.shard_indices(4, 2, 2)
    x = torch.arange(6.0).reshape(3, 2)
    assert torch.equal(parallel.gather_samples(x, 3), x)
    assert torch.equal(parallel.run_sharded([1, 2, 3], lambda i, s: torch.tensor([float(s)])), torch.tensor([[1.0], [2.0], [3.0]]))
