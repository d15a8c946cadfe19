"""Self-consistency invariants of the oracle.  CPU only, tiny config, fp32."""
import numpy as np
import pytest
import torch

from oracle import flux_oracle as O
from reptext_b200 import config, weights
from util import rel_l2, synth_inputs

TR, CN = config.TINY_TRANSFORMER, config.TINY_CONTROLNET
H = W = 128          # N = 64 image tokens
T = 16


@pytest.fixture(scope="module")
def setup():
    tr = weights.random_state_dict(TR, "transformer", seed=1)
    cn = weights.random_state_dict(CN, "controlnet", seed=2)
    cn0 = weights.random_state_dict(CN, "controlnet", seed=2, zero_init=True)
    x = synth_inputs(TR, CN, H, W, T, seed=3)
    return tr, cn, cn0, x


def _cn(cn, x, scale=1.0, lat=None, cond=None, img_ids=None):
    t = torch.tensor([0.7])
    g = torch.tensor([3.5])
    return O.controlnet_forward(cn, CN, x["latents"] if lat is None else lat, x["conds"][0] if cond is None else cond,
                                scale, x["prompt_embeds"], x["pooled"], t, x["img_ids"] if img_ids is None else img_ids,
                                x["txt_ids"], g)


def _tr(tr, x, blk=None, sgl=None, lat=None, img_ids=None):
    t = torch.tensor([0.7])
    g = torch.tensor([3.5])
    return O.transformer_forward(tr, TR, x["latents"] if lat is None else lat, x["prompt_embeds"], x["pooled"], t,
                                 x["img_ids"] if img_ids is None else img_ids, x["txt_ids"], g, blk, sgl)


def test_zero_init_controlnet_is_zero(setup):                       # A.10 (i)
    tr, cn, cn0, x = setup
    blk, sgl = _cn(cn0, x)
    assert sgl is None and len(blk) == CN["num_layers"]
    assert all(float(b.abs().max()) == 0.0 for b in blk)
    assert torch.equal(_tr(tr, x, blk), _tr(tr, x, None))


def test_conditioning_scale_linear(setup):                          # A.10 (ii)
    tr, cn, cn0, x = setup
    a, _ = _cn(cn, x, 1.0)
    b, _ = _cn(cn, x, 0.25)
    for u, v in zip(a, b):
        assert rel_l2(v, 0.25 * u) < 1e-6


def test_mask_ones_and_zeros(setup):                                # A.10 (iii)
    tr, cn, cn0, x = setup
    ts, sg = O.make_sigmas(2, x["N"])
    common = dict(latents=x["latents"], prompt_embeds=x["prompt_embeds"], pooled=x["pooled"],
                  control_image_list=x["conds"], text_ids=x["txt_ids"], img_ids=x["img_ids"],
                  timesteps=ts, sigmas=sg, guidance_scale=3.5)
    ones = [torch.ones(1, x["N"], 1)]
    zeros = [torch.zeros(1, x["N"], 1)]
    a = O.denoise_t2i(tr, TR, cn, CN, control_mask_list=ones, **common)
    b = O.denoise_t2i(tr, TR, cn, CN, control_mask_list=[], **common)
    assert torch.equal(a, b)
    c = O.denoise_t2i(tr, TR, cn, CN, control_mask_list=zeros, **common)
    d = O.denoise_t2i(tr, TR, cn0, CN, control_mask_list=[], **common)
    assert rel_l2(c, d) < 1e-6
    assert rel_l2(a, c) > 1e-3      # the ControlNet does something


def test_token_permutation_equivariance(setup):                     # A.10 (iv)
    tr, cn, cn0, x = setup
    perm = torch.randperm(x["N"], generator=torch.Generator().manual_seed(5))
    ref = _tr(tr, x)
    out = _tr(tr, x, lat=x["latents"][:, perm], img_ids=x["img_ids"][perm])
    assert rel_l2(out, ref[:, perm]) < 1e-5


def test_pack_unpack_roundtrip():                                   # A.10 (v)
    z = torch.randn(2, 16, 16, 24)
    p = O.pack_latents(z)
    assert p.shape == (2, 8 * 12, 64)
    assert torch.equal(O.unpack_latents(p, 16 * 8, 24 * 8), z)


def test_text_rows_rope_invariant():                                # A.10 (vi)
    ids = torch.cat([torch.zeros(4, 3), O.prepare_latent_image_ids(8, 8)])
    cos, sin = O.rope_table(ids, (16, 24, 24))
    assert cos.shape == (4 + 16, 64) and cos.dtype == torch.float32
    assert torch.all(cos[:4] == 1) and torch.all(sin[:4] == 0)
    assert torch.all(cos[:, :16] == 1) and torch.all(sin[:, :16] == 0)   # axis 0 ids are all zero
    q = torch.randn(1, 20, 2, 64)
    assert torch.equal(O.apply_rope(q, (cos, sin))[:, :4], q[:, :4])


def test_cfg_batch_broadcast(setup):                                # A.10 (vii)
    tr, cn, cn0, x = setup
    pe2 = torch.cat([x["prompt_embeds"] * 0.5, x["prompt_embeds"]])
    po2 = torch.cat([x["pooled"] * 0.5, x["pooled"]])
    t, g = torch.tensor([0.7]), torch.tensor([3.5])
    a = O.transformer_forward(tr, TR, x["latents"], pe2, po2, t, x["img_ids"], x["txt_ids"], g)
    b = O.transformer_forward(tr, TR, x["latents"].repeat(2, 1, 1), pe2, po2, t.repeat(2), x["img_ids"], x["txt_ids"],
                              g.repeat(2))
    assert a.shape[0] == 2 and rel_l2(a, b) < 1e-6


def test_euler_zero_velocity_and_promotion():                       # A.10 (viii) + A.7 promotion
    x = torch.randn(1, 8, 64)
    ts, sg = O.make_sigmas(4, 256)
    assert torch.equal(O.euler_step(torch.zeros_like(x), sg[0], sg[1], x), x)
    xb, vb = x.bfloat16(), torch.randn(1, 8, 64).bfloat16()
    out = O.euler_step(vb, sg[0], sg[1], xb)
    # torch CPU promotion: the 0-dim fp32 dt is cast to bf16, the product is rounded to bf16, the add is fp32.
    # (On CUDA with the scheduler's CPU-resident sigmas dt stays fp32: the product is bf16(dt_f32 * v);
    #  the CUDA kernel follows that form -- tests/test_kernels_gpu.py checks it bit-exactly.)
    dt = (sg[1] - sg[0]).bfloat16().float()
    want = (xb.float() + (dt * vb.float()).bfloat16().float()).bfloat16()
    assert out.dtype == torch.bfloat16 and torch.equal(out, want)


def test_sigmas_schedule():
    ts, sg = O.make_sigmas(28, 4096)
    assert ts.shape == (28,) and sg.shape == (29,) and sg[-1] == 0 and abs(float(sg[0]) - 1.0) < 1e-6
    assert abs(O.calculate_shift(4096, 256, 4096, 0.5, 1.15) - 1.15) < 1e-9
    assert abs(O.calculate_shift(9216, 256, 4096, 0.5, 1.15) - 2.0167) < 1e-3
    assert torch.all(sg[:-1] > sg[1:])
    # mu=1.15: sigma' = e^mu / (e^mu + 1/sigma - 1)
    s = 1 - 13 / 28 * (1 - 1 / 28) * 28 / 27
    assert abs(float(sg[13]) - np.exp(1.15) / (np.exp(1.15) + 1 / np.linspace(1, 1 / 28, 28)[13] - 1)) < 1e-6


def test_controlnet_block6_unused_by_base(setup):
    """A.6: with 19 base blocks and 6 samples, interval=4 -> sample 5 is never consumed."""
    nl, n = 19, 6
    interval = int(np.ceil(nl / n))
    assert sorted({i // interval for i in range(nl)}) == [0, 1, 2, 3, 4]


def test_glyph_init_live_vs_dead():
    img = torch.zeros(1, 3, 64, 64)
    img[:, :, 16:32, 16:48] = 1.0
    z = torch.randn(1, 16, 8, 8)
    noise = torch.randn(1, 16, 8, 8)
    dead = O.glyph_latent_init(img, z, noise, live=False)
    live = O.glyph_latent_init(img, z, noise, live=True)
    assert torch.equal(dead, O.pack_latents(noise))
    assert not torch.equal(live, dead)
    diff = O.unpack_latents(live - dead, 64, 64)
    assert float(diff[:, :, 0, 0].abs().max()) == 0.0
    assert rel_l2(diff[:, :, 3, 3], 0.10 * z[:, :, 3, 3]) < 1e-6
