"""Sequence-parallel mode (BASELINE.json configs[4]; include/reptext_rt.h rt_sp_group) on the GPU, through the C-ABI.

On ONE GPU the ranks are driven in lock step (rt_*_forward_lockstep): same kernels, same peer-store indexing as the
multi-process mode, stream order instead of the flag barriers.  Checked against
  * the unsharded run of the same library (same bf16 kernels; only the key order of the softmax differs), and
  * the fp32 oracle (<= 1e-2 rel-L2, the bf16 bar of BASELINE.json).
With >= 2 GPUs the real thing runs too: one process per GPU, CUDA-IPC mapped workspaces, flag barriers, NCCL gather.
"""
import os
import subprocess
import sys

import pytest
import torch

from util import rel_l2, synth_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gemm_scatter_matches_plain_projection():
    """q | k | v segments scattered by head block into `world` exchange buffers == the plain [rows, 3D] projection."""
    from reptext_b200 import _lib as L, ops
    import ctypes as C
    torch.manual_seed(0)
    dt, dev = torch.bfloat16, "cuda"
    world, H, hd, K, rows, rank = 4, 8, 128, 256, 200, 2
    D, Dl, Sg = H * hd, H * hd // world, rows * world
    A = torch.randn(1, rows, K, device=dev, dtype=dt)
    Ws = [torch.randn(D, K, device=dev, dtype=dt) * K ** -0.5 for _ in range(3)]
    bs = [torch.randn(D, device=dev, dtype=dt) for _ in range(3)]
    nw = (1 + 0.1 * torch.randn(hd, device=dev)).to(dt)
    ids = torch.zeros(rows, 3, device=dev)
    ids[:, 1] = torch.arange(rows, device=dev) // 16
    ids[:, 2] = torch.arange(rows, device=dev) % 16
    rope = ops.rope_table(ids, (16, 56, 56))
    plain = torch.zeros(1, rows, 3 * D, device=dev, dtype=dt)
    modes = [L.EPI_QKNORM_ROPE, L.EPI_QKNORM_ROPE, L.EPI_BIAS]
    segs = [ops.Segment(W=Ws[i], bias=bs[i], out=plain, mode=modes[i], out_col0=i * D,
                        norm_w=nw if i < 2 else None) for i in range(3)]
    ops.gemm([ops.Problem(A=A, segs=segs)], 1, dt, rope=rope, head_dim=hd, impl=ops.IMPL_TC1)
    for impl in (ops.IMPL_TC1, ops.IMPL_TC2):
        bufs = [torch.zeros(1, Sg, 3 * Dl, device=dev, dtype=dt) for _ in range(world)]
        segs = [ops.Segment(W=Ws[i], bias=bs[i], out=bufs[0], mode=modes[i], out_col0=i * Dl,
                            norm_w=nw if i < 2 else None, scatter=True) for i in range(3)]
        ops.gemm([ops.Problem(A=A, segs=segs)], 1, dt, rope=rope, head_dim=hd, impl=impl,
                 sp_out=bufs, sp_cols=Dl, sp_row0=rank * rows)
        for r in range(world):
            got = bufs[r][0, rank * rows:(rank + 1) * rows]
            for i in range(3):
                want = plain[0, :, i * D + r * Dl: i * D + (r + 1) * Dl]
                assert torch.equal(got[:, i * Dl:(i + 1) * Dl], want), (impl, r, i)
            other = torch.cat([bufs[r][0, :rank * rows], bufs[r][0, (rank + 1) * rows:]])
            assert not other.any()          # nothing outside this rank's row range was touched


def test_attention_row_scatter_matches_plain_attention():
    from reptext_b200 import ops
    torch.manual_seed(1)
    dt, dev = torch.bfloat16, "cuda"
    world, heads, hd, S_loc, rank = 4, 2, 128, 72, 1
    Sg, Dl, D = S_loc * world, heads * hd, heads * hd * world
    qkv = torch.randn(1, Sg, 3 * Dl, device=dev, dtype=dt)
    plain = ops.attention(qkv, heads, hd, 0, Dl, 2 * Dl, impl=ops.IMPL_TC1)
    outs = [torch.zeros(1, S_loc, 5 * D, device=dev, dtype=dt) for _ in range(world)]
    ops.attention(qkv, heads, hd, 0, Dl, 2 * Dl, out=outs[0], out_col0=rank * Dl, impl=ops.IMPL_TC1,
                  sp_out=outs, sp_rows=S_loc)
    for r in range(world):
        assert torch.equal(outs[r][0, :, rank * Dl:(rank + 1) * Dl], plain[0, r * S_loc:(r + 1) * S_loc])
        assert not outs[r][0, :, :rank * Dl].any() and not outs[r][0, :, (rank + 1) * Dl:].any()


def _models(dtype=torch.bfloat16, seed=100):
    from reptext_b200 import config, models, weights
    TR, CN = config.SP8_TRANSFORMER, config.SP8_CONTROLNET
    tr_sd = {k: v.to(dtype).float() for k, v in weights.random_state_dict(TR, "transformer", seed=seed).items()}
    cn_sd = {k: v.to(dtype).float() for k, v in weights.random_state_dict(CN, "controlnet", seed=seed + 1).items()}
    return TR, CN, models.FluxTransformer2DModel(TR, tr_sd, dtype=dtype), models.FluxControlNetModel(CN, cn_sd, dtype=dtype), tr_sd, cn_sd


def _shard(t, r, w, dim=1):
    from reptext_b200.parallel import shard_tokens
    return shard_tokens(t, r, w, dim)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("batch", [1, 2])
def test_lockstep_forward_matches_unsharded_and_oracle(world, batch):
    from oracle import flux_oracle as O
    from reptext_b200.parallel import LockstepGroup
    dt, dev = torch.bfloat16, "cuda"
    TR, CN, tr, cn, tr_sd, cn_sd = _models()
    H, Wd, T = 512, 256, 64                       # N = 512 image tokens; S = 576 (ragged last key tile)
    x = synth_inputs(TR, CN, H, Wd, T, seed=11, batch=batch, n_lines=1)
    x = {k: ([t.to(dt) for t in v] if isinstance(v, list) else (v.to(dt) if torch.is_tensor(v) else v)) for k, v in x.items()}
    xg = {k: ([t.to(dev) for t in v] if isinstance(v, list) else (v.to(dev) if torch.is_tensor(v) else v)) for k, v in x.items()}
    t = torch.full((batch,), 0.62, device=dev, dtype=dt)
    g = torch.full((batch,), 3.5, device=dev, dtype=dt)
    mask = xg["masks"][0]
    kw = dict(hidden_states=xg["latents"], encoder_hidden_states=xg["prompt_embeds"], pooled_projections=xg["pooled"],
              timestep=t, guidance=g, img_ids=xg["img_ids"], txt_ids=xg["txt_ids"])
    # ---- unsharded run of the same library
    bl, _ = cn(controlnet_cond=xg["conds"][0], conditioning_scale=0.8, regional_mask=mask, return_dict=False, **kw)
    ref = tr(controlnet_block_samples=bl, return_dict=False, **kw)[0]
    # ---- lock-step sequence-parallel run
    grp = LockstepGroup(world, dev)
    per_rank = []
    for r in range(world):
        per_rank.append(dict(
            hidden_states=_shard(xg["latents"], r, world), encoder_hidden_states=_shard(xg["prompt_embeds"], r, world),
            pooled_projections=xg["pooled"], timestep=t, guidance=g, img_ids=_shard(xg["img_ids"], r, world, 0),
            txt_ids=_shard(xg["txt_ids"], r, world, 0)))
    cn_out = cn.forward_lockstep(grp, [dict(p, controlnet_cond=_shard(xg["conds"][0], r, world), conditioning_scale=0.8,
                                            regional_mask=_shard(mask.reshape(1, -1, 1), r, world))
                                       for r, p in enumerate(per_rank)])
    bl_sp = [torch.cat([cn_out[r][0][i] for r in range(world)], dim=1) for i in range(len(bl))]
    for a, b in zip(bl_sp, bl):
        assert rel_l2(a, b) < 4e-3
    outs = tr.forward_lockstep(grp, [dict(p, controlnet_block_samples=cn_out[r][0]) for r, p in enumerate(per_rank)])
    got = torch.cat(outs, dim=1)
    assert torch.isfinite(got.float()).all()
    assert rel_l2(got, ref) < 4e-3, rel_l2(got, ref)
    # ---- fp32 oracle on the same bf16-rounded weights and inputs
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        f = lambda v: v.float()
        # the library embeds `timestep.to(bf16) * 1000` computed IN bf16 (controlnet_flux.py:282-284); so does the oracle
        to, go = (t * 1000).float() / 1000, (g * 1000).float() / 1000
        osd_tr = {k: v.to(dev) for k, v in tr_sd.items()}
        osd_cn = {k: v.to(dev) for k, v in cn_sd.items()}
        with torch.no_grad():
            ob, _ = O.controlnet_forward(osd_cn, CN, f(xg["latents"]), f(xg["conds"][0]), 0.8, f(xg["prompt_embeds"]),
                                         f(xg["pooled"]), to, f(xg["img_ids"]), f(xg["txt_ids"]), go)
            ob = [b * f(mask) for b in ob]
            onp = O.transformer_forward(osd_tr, TR, f(xg["latents"]), f(xg["prompt_embeds"]), f(xg["pooled"]), to,
                                        f(xg["img_ids"]), f(xg["txt_ids"]), go, ob, None)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert rel_l2(got, onp) < 1e-2, rel_l2(got, onp)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_lockstep_forward_is_bit_identical_when_the_shards_align(world):
    """When every rank's text rows are a multiple of 64 and its image rows a multiple of 128, the attention kernel walks
    the keys in the UNSHARDED order (rt_attention_args.sp_txt_rows): same key blocks, same exp2 path per key, same fp32
    summation order.  Every other kernel computes a row independently of its neighbours, so the sharded forward is the
    single-GPU forward bit for bit (ControlNet samples and noise prediction) - BASELINE.json configs[4]'s shapes
    (64 + 1152 rows per rank at 8 ranks) are of this kind; tests/test_fullsize_gpu.py checks them at full size."""
    from reptext_b200.parallel import LockstepGroup
    dt, dev = torch.bfloat16, "cuda"
    TR, CN, tr, cn, tr_sd, cn_sd = _models()
    H, Wd, T = 1024, 512, 512                      # N = 2048 image + 512 text tokens: 64 + 256 rows per rank at world 8
    x = synth_inputs(TR, CN, H, Wd, T, seed=12, batch=1, n_lines=1)
    xg = {k: ([t.to(dev, dt) for t in v] if isinstance(v, list) else (v.to(dev, dt) if torch.is_tensor(v) else v)) for k, v in x.items()}
    t = torch.full((1,), 0.62, device=dev, dtype=dt)
    g = torch.full((1,), 3.5, device=dev, dtype=dt)
    mask = xg["masks"][0]
    kw = dict(hidden_states=xg["latents"], encoder_hidden_states=xg["prompt_embeds"], pooled_projections=xg["pooled"],
              timestep=t, guidance=g, img_ids=xg["img_ids"], txt_ids=xg["txt_ids"])
    bl, _ = cn(controlnet_cond=xg["conds"][0], conditioning_scale=0.8, regional_mask=mask, return_dict=False, **kw)
    ref = tr(controlnet_block_samples=bl, return_dict=False, **kw)[0]
    grp = LockstepGroup(world, dev)
    per_rank = [dict(hidden_states=_shard(xg["latents"], r, world), encoder_hidden_states=_shard(xg["prompt_embeds"], r, world),
                     pooled_projections=xg["pooled"], timestep=t, guidance=g, img_ids=_shard(xg["img_ids"], r, world, 0),
                     txt_ids=_shard(xg["txt_ids"], r, world, 0)) for r in range(world)]
    cn_out = cn.forward_lockstep(grp, [dict(p, controlnet_cond=_shard(xg["conds"][0], r, world), conditioning_scale=0.8,
                                            regional_mask=_shard(mask.reshape(1, -1, 1), r, world))
                                       for r, p in enumerate(per_rank)])
    for i in range(len(bl)):
        assert torch.equal(torch.cat([cn_out[r][0][i] for r in range(world)], dim=1), bl[i]), i
    outs = tr.forward_lockstep(grp, [dict(p, controlnet_block_samples=cn_out[r][0]) for r, p in enumerate(per_rank)])
    assert torch.equal(torch.cat(outs, dim=1), ref)


def test_sequence_parallel_needs_the_tensor_core_path():
    from reptext_b200 import config, models
    from reptext_b200.parallel import LockstepGroup
    tr = models.FluxTransformer2DModel.random_init(config.TINY_TRANSFORMER, dtype=torch.float32)
    grp = LockstepGroup(2, "cuda")
    z = torch.zeros
    with pytest.raises(ValueError, match="sequence-parallel"):
        tr.forward_lockstep(grp, [dict(hidden_states=z(1, 8, 64, device="cuda"), encoder_hidden_states=z(1, 4, 64, device="cuda"),
                                       pooled_projections=z(1, 32, device="cuda"), timestep=z(1, device="cuda"),
                                       guidance=z(1, device="cuda"), img_ids=z(8, 3, device="cuda"),
                                       txt_ids=z(4, 3, device="cuda"))] * 2)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs on one node")
def test_multi_process_sequence_parallel_pipeline_matches_single_gpu():
    """torchrun, one process per GPU: CUDA-IPC workspaces, peer stores from the GEMM / attention epilogues, flag
    barriers, NCCL gather of the latents.  The worker compares against the single-GPU run of the same pipeline."""
    n = min(torch.cuda.device_count(), 8)
    n = {2: 2, 3: 2, 4: 4, 5: 4, 6: 4, 7: 4, 8: 8}[n]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "sp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sp_worker.log"), "w") as fh:
        fh.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SP_WORKER_OK" in r.stdout
