"""Sequence-parallel mode on the CPU (world_size-2 and -4 gloo processes): the token partition of
reptext_b200.parallel (shard_tokens / gather_tokens) and the head <-> token re-partition the CUDA kernels implement
with peer stores, restated over the ORACLE with torch.distributed collectives.  A sharded run must reproduce the
unsharded oracle: attention does not depend on the order of the keys, RoPE follows the ids each rank holds, and
every other operation of the path is token-wise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reptext_b200 import config, parallel, weights


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ulysses_attention(q, k, v):
    """[B, S_local, H, hd] per rank -> this rank's heads over ALL tokens -> back to this rank's tokens, all heads.
    Row r*S_local + i of the exchanged sequence is rank r's local row i (the layout of rt_sp_group)."""
    import torch.nn.functional as F
    world, rank = dist.get_world_size(), dist.get_rank()
    H = q.shape[2]
    hpr = H // world

    def heads_of_mine_all_tokens(t):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts, dim=1)[:, :, rank * hpr:(rank + 1) * hpr]

    Q, K, V = (heads_of_mine_all_tokens(t) for t in (q, k, v))
    o = F.scaled_dot_product_attention(Q.transpose(1, 2), K.transpose(1, 2), V.transpose(1, 2)).transpose(1, 2)
    parts = [torch.empty_like(o) for _ in range(world)]
    dist.all_gather(parts, o.contiguous())                # parts[r]: rank r's heads, all tokens
    S = q.shape[1]
    mine = torch.cat([p[:, rank * S:(rank + 1) * S] for p in parts], dim=2)   # my tokens, heads in rank order
    return mine.flatten(2, 3).to(q.dtype)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    parallel.init_from_env("gloo")
    from oracle import flux_oracle as O
    from util import synth_inputs
    TR, CN = config.SP8_TRANSFORMER, config.SP8_CONTROLNET
    tr_sd = weights.random_state_dict(TR, "transformer", seed=100)
    cn_sd = weights.random_state_dict(CN, "controlnet", seed=101)
    x = synth_inputs(TR, CN, 128, 256, 16, seed=3, batch=1, n_lines=1)
    t, g = torch.tensor([0.62]), torch.tensor([3.5])
    with torch.no_grad():
        ob, _ = O.controlnet_forward(cn_sd, CN, x["latents"], x["conds"][0], 0.8, x["prompt_embeds"], x["pooled"], t,
                                     x["img_ids"], x["txt_ids"], g)
        want = O.transformer_forward(tr_sd, TR, x["latents"], x["prompt_embeds"], x["pooled"], t, x["img_ids"],
                                     x["txt_ids"], g, ob, None)
        sh = lambda a, dim=1: parallel.shard_tokens(a, rank, world, dim)
        O._attention = _ulysses_attention                 # the exchange, in place of the single-GPU SDPA
        ob_l, _ = O.controlnet_forward(cn_sd, CN, sh(x["latents"]), sh(x["conds"][0]), 0.8, sh(x["prompt_embeds"]),
                                       x["pooled"], t, sh(x["img_ids"], 0), sh(x["txt_ids"], 0), g)
        got_l = O.transformer_forward(tr_sd, TR, sh(x["latents"]), sh(x["prompt_embeds"]), x["pooled"], t,
                                      sh(x["img_ids"], 0), sh(x["txt_ids"], 0), g, ob_l, None)
    got = parallel.gather_tokens(got_l)
    err = float((got - want).norm() / want.norm())
    q.put((rank, err, tuple(got.shape), tuple(want.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_oracle_reproduces_the_unsharded_oracle(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, gs, ws in res:
        assert gs == ws
        assert err < 2e-5, (rank, err)


def test_shard_and_gather_tokens_single_process():
    x = torch.arange(24.0).reshape(1, 12, 2)
    parts = [parallel.shard_tokens(x, r, 4) for r in range(4)]
    assert all(p.shape == (1, 3, 2) and p.is_contiguous() for p in parts)
    assert torch.equal(torch.cat(parts, dim=1), x)
    assert torch.equal(parallel.gather_tokens(parts[0]), parts[0])      # no process group: identity
    with pytest.raises(ValueError):
        parallel.shard_tokens(x, 0, 5)


def test_lockstep_group_validates_world():
    with pytest.raises(ValueError):
        parallel.LockstepGroup(9, "cpu")
