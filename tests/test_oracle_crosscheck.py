"""Cross-check the oracle's block math against an INDEPENDENT implementation that is on this box:
torchtitan's port of the Black-Forest-Labs FLUX blocks.  Weight remap:
img_attn.qkv = cat(to_q,to_k,to_v), txt_attn.qkv = cat(add_q,add_k,add_v), img_attn.proj = to_out.0, ...
"""
import pytest
import torch

from oracle import flux_oracle as O
from reptext_b200 import config, weights
from util import rel_l2

layers = pytest.importorskip("torchtitan.experiments.flux.model.layers")
tmath = pytest.importorskip("torchtitan.experiments.flux.model.math")

CFG = config.TINY_TRANSFORMER
HD, H = CFG["attention_head_dim"], CFG["num_attention_heads"]
D = HD * H


def _inputs():
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 48, D, generator=g)
    c = torch.randn(2, 16, D, generator=g)
    temb = torch.randn(2, D, generator=g)
    ids = torch.cat([torch.zeros(16, 3), O.prepare_latent_image_ids(12, 16)])
    return x, c, temb, ids


def _bfl_pe(ids):
    axes = CFG["axes_dims_rope"]
    pe = torch.cat([tmath.rope(ids[None, :, i].double(), axes[i], 10000) for i in range(3)], dim=-3)
    return pe.unsqueeze(1)          # [1,1,S,hd/2,2,2]


def test_double_block_matches_bfl():
    sd = weights.random_state_dict(CFG, "transformer", seed=4)
    p = "transformer_blocks.0."
    x, c, temb, ids = _inputs()
    blk = layers.DoubleStreamBlock(D, H, mlp_ratio=4.0, qkv_bias=True)
    with torch.no_grad():
        blk.img_mod.lin.weight.copy_(sd[p + "norm1.linear.weight"]); blk.img_mod.lin.bias.copy_(sd[p + "norm1.linear.bias"])
        blk.txt_mod.lin.weight.copy_(sd[p + "norm1_context.linear.weight"]); blk.txt_mod.lin.bias.copy_(sd[p + "norm1_context.linear.bias"])
        blk.img_attn.qkv.weight.copy_(torch.cat([sd[p + f"attn.{n}.weight"] for n in ("to_q", "to_k", "to_v")]))
        blk.img_attn.qkv.bias.copy_(torch.cat([sd[p + f"attn.{n}.bias"] for n in ("to_q", "to_k", "to_v")]))
        blk.txt_attn.qkv.weight.copy_(torch.cat([sd[p + f"attn.{n}.weight"] for n in ("add_q_proj", "add_k_proj", "add_v_proj")]))
        blk.txt_attn.qkv.bias.copy_(torch.cat([sd[p + f"attn.{n}.bias"] for n in ("add_q_proj", "add_k_proj", "add_v_proj")]))
        blk.img_attn.proj.weight.copy_(sd[p + "attn.to_out.0.weight"]); blk.img_attn.proj.bias.copy_(sd[p + "attn.to_out.0.bias"])
        blk.txt_attn.proj.weight.copy_(sd[p + "attn.to_add_out.weight"]); blk.txt_attn.proj.bias.copy_(sd[p + "attn.to_add_out.bias"])
        for a, nq, nk in ((blk.img_attn, "norm_q", "norm_k"), (blk.txt_attn, "norm_added_q", "norm_added_k")):
            a.norm.query_norm.weight.copy_(sd[p + f"attn.{nq}.weight"]); a.norm.query_norm.eps = 1e-6
            a.norm.key_norm.weight.copy_(sd[p + f"attn.{nk}.weight"]); a.norm.key_norm.eps = 1e-6
        for m, ff in ((blk.img_mlp, "ff"), (blk.txt_mlp, "ff_context")):
            m[0].weight.copy_(sd[p + f"{ff}.net.0.proj.weight"]); m[0].bias.copy_(sd[p + f"{ff}.net.0.proj.bias"])
            m[2].weight.copy_(sd[p + f"{ff}.net.2.weight"]); m[2].bias.copy_(sd[p + f"{ff}.net.2.bias"])
        img, txt = blk(x, c, temb, _bfl_pe(ids))
    rope = O.rope_table(ids, CFG["axes_dims_rope"])
    c2, x2 = O.double_block(sd, p, x, c, temb, rope, H)
    assert rel_l2(x2, img) < 2e-5 and rel_l2(c2, txt) < 2e-5


def test_single_block_matches_bfl():
    sd = weights.random_state_dict(CFG, "transformer", seed=4)
    p = "single_transformer_blocks.1."
    x, c, temb, ids = _inputs()
    blk = layers.SingleStreamBlock(D, H, mlp_ratio=4.0)
    with torch.no_grad():
        blk.modulation.lin.weight.copy_(sd[p + "norm.linear.weight"]); blk.modulation.lin.bias.copy_(sd[p + "norm.linear.bias"])
        blk.linear1.weight.copy_(torch.cat([sd[p + f"attn.{n}.weight"] for n in ("to_q", "to_k", "to_v")] + [sd[p + "proj_mlp.weight"]]))
        blk.linear1.bias.copy_(torch.cat([sd[p + f"attn.{n}.bias"] for n in ("to_q", "to_k", "to_v")] + [sd[p + "proj_mlp.bias"]]))
        blk.linear2.weight.copy_(sd[p + "proj_out.weight"]); blk.linear2.bias.copy_(sd[p + "proj_out.bias"])
        blk.norm.query_norm.weight.copy_(sd[p + "attn.norm_q.weight"]); blk.norm.query_norm.eps = 1e-6
        blk.norm.key_norm.weight.copy_(sd[p + "attn.norm_k.weight"]); blk.norm.key_norm.eps = 1e-6
        h = blk(torch.cat([c, x], dim=1), temb, _bfl_pe(ids))
    rope = O.rope_table(ids, CFG["axes_dims_rope"])
    c2, x2 = O.single_block(sd, p, x, c, temb, rope, H)
    assert rel_l2(torch.cat([c2, x2], dim=1), h) < 2e-5


def test_last_layer_matches_bfl():
    """LastLayer chunks (shift, scale); diffusers norm_out.linear chunks (scale, shift)."""
    sd = weights.random_state_dict(CFG, "transformer", seed=4)
    x, c, temb, ids = _inputs()
    ll = layers.LastLayer(D, 1, 64)
    w, b = sd["norm_out.linear.weight"], sd["norm_out.linear.bias"]
    with torch.no_grad():
        ll.adaLN_modulation[1].weight.copy_(torch.cat([w[D:], w[:D]])); ll.adaLN_modulation[1].bias.copy_(torch.cat([b[D:], b[:D]]))
        ll.linear.weight.copy_(sd["proj_out.weight"]); ll.linear.bias.copy_(sd["proj_out.bias"])
        ref = ll(x, temb)
    import torch.nn.functional as F
    emb = F.linear(F.silu(temb), w, b)
    scale, shift = emb.chunk(2, dim=1)
    out = F.linear(O._ln(x) * (1 + scale)[:, None] + shift[:, None], sd["proj_out.weight"], sd["proj_out.bias"])
    assert rel_l2(out, ref) < 2e-5


# ---------------------------------------------------------------------------------------------------------
# Whole-model and pipeline-level pins against the same independent implementation (BFL / torchtitan)
# ---------------------------------------------------------------------------------------------------------
def _copy_linear(dst, sd, name):
    dst.weight.copy_(sd[name + ".weight"])
    dst.bias.copy_(sd[name + ".bias"])


def _load_double(blk, sd, p):
    _copy_linear(blk.img_mod.lin, sd, p + "norm1.linear")
    _copy_linear(blk.txt_mod.lin, sd, p + "norm1_context.linear")
    for attn, names in ((blk.img_attn, ("to_q", "to_k", "to_v")), (blk.txt_attn, ("add_q_proj", "add_k_proj", "add_v_proj"))):
        attn.qkv.weight.copy_(torch.cat([sd[p + f"attn.{n}.weight"] for n in names]))
        attn.qkv.bias.copy_(torch.cat([sd[p + f"attn.{n}.bias"] for n in names]))
    _copy_linear(blk.img_attn.proj, sd, p + "attn.to_out.0")
    _copy_linear(blk.txt_attn.proj, sd, p + "attn.to_add_out")
    for a, nq, nk in ((blk.img_attn, "norm_q", "norm_k"), (blk.txt_attn, "norm_added_q", "norm_added_k")):
        a.norm.query_norm.weight.copy_(sd[p + f"attn.{nq}.weight"]); a.norm.query_norm.eps = 1e-6
        a.norm.key_norm.weight.copy_(sd[p + f"attn.{nk}.weight"]); a.norm.key_norm.eps = 1e-6
    for m, ff in ((blk.img_mlp, "ff"), (blk.txt_mlp, "ff_context")):
        _copy_linear(m[0], sd, p + f"{ff}.net.0.proj")
        _copy_linear(m[2], sd, p + f"{ff}.net.2")


def _load_single(blk, sd, p):
    _copy_linear(blk.modulation.lin, sd, p + "norm.linear")
    blk.linear1.weight.copy_(torch.cat([sd[p + f"attn.{n}.weight"] for n in ("to_q", "to_k", "to_v")] + [sd[p + "proj_mlp.weight"]]))
    blk.linear1.bias.copy_(torch.cat([sd[p + f"attn.{n}.bias"] for n in ("to_q", "to_k", "to_v")] + [sd[p + "proj_mlp.bias"]]))
    _copy_linear(blk.linear2, sd, p + "proj_out")
    blk.norm.query_norm.weight.copy_(sd[p + "attn.norm_q.weight"]); blk.norm.query_norm.eps = 1e-6
    blk.norm.key_norm.weight.copy_(sd[p + "attn.norm_k.weight"]); blk.norm.key_norm.eps = 1e-6


def test_whole_transformer_matches_bfl_flux_model():
    """Embedders (timestep sinusoid x1000, pooled-text MLP), FluxPosEmbed, the double / single block loops and the final
    AdaLayerNormContinuous + proj_out, end to end, against torchtitan's FluxModel (no guidance embedder there, so the
    oracle runs with guidance_embeds=False; the guidance branch is the same MLP as the timestep one)."""
    model_mod = pytest.importorskip("torchtitan.experiments.flux.model.model")
    args_mod = pytest.importorskip("torchtitan.experiments.flux.model.args")
    cfg = dict(CFG, guidance_embeds=False)
    sd = weights.random_state_dict(cfg, "transformer", seed=21)
    m = model_mod.FluxModel(args_mod.FluxModelArgs(
        in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], vec_in_dim=cfg["pooled_projection_dim"],
        context_in_dim=cfg["joint_attention_dim"], hidden_size=D, mlp_ratio=4.0, num_heads=H, depth=cfg["num_layers"],
        depth_single_blocks=cfg["num_single_layers"], axes_dim=tuple(cfg["axes_dims_rope"]), theta=10000, qkv_bias=True))
    with torch.no_grad():
        _copy_linear(m.img_in, sd, "x_embedder")
        _copy_linear(m.txt_in, sd, "context_embedder")
        _copy_linear(m.time_in.in_layer, sd, "time_text_embed.timestep_embedder.linear_1")
        _copy_linear(m.time_in.out_layer, sd, "time_text_embed.timestep_embedder.linear_2")
        _copy_linear(m.vector_in.in_layer, sd, "time_text_embed.text_embedder.linear_1")
        _copy_linear(m.vector_in.out_layer, sd, "time_text_embed.text_embedder.linear_2")
        for i, blk in enumerate(m.double_blocks):
            _load_double(blk, sd, f"transformer_blocks.{i}.")
        for j, blk in enumerate(m.single_blocks):
            _load_single(blk, sd, f"single_transformer_blocks.{j}.")
        w, b = sd["norm_out.linear.weight"], sd["norm_out.linear.bias"]
        m.final_layer.adaLN_modulation[1].weight.copy_(torch.cat([w[D:], w[:D]]))       # BFL chunks (shift, scale)
        m.final_layer.adaLN_modulation[1].bias.copy_(torch.cat([b[D:], b[:D]]))
        _copy_linear(m.final_layer.linear, sd, "proj_out")
        g = torch.Generator().manual_seed(3)
        B, T, lh, lw = 2, 24, 12, 16
        N = (lh // 2) * (lw // 2)
        lat = torch.randn(B, N, cfg["in_channels"], generator=g)
        txt = torch.randn(B, T, cfg["joint_attention_dim"], generator=g)
        pooled = torch.randn(B, cfg["pooled_projection_dim"], generator=g)
        t = torch.tensor([0.83, 0.31])
        img_ids, txt_ids = O.prepare_latent_image_ids(lh, lw), torch.zeros(T, 3)
        ref = m(img=lat, img_ids=img_ids[None].expand(B, -1, -1), txt=txt, txt_ids=txt_ids[None].expand(B, -1, -1),
                timesteps=t, y=pooled)
        got = O.transformer_forward(sd, cfg, lat, txt, pooled, t, img_ids, txt_ids, None, None, None)
    assert rel_l2(got, ref) < 5e-5, rel_l2(got, ref)


def test_schedule_pack_and_ids_match_bfl():
    """calculate_shift + the shifted sigma grid (RepText/pipeline_flux_controlnet.py:948-967, :78-88), the 2x2 patch
    packing (:550-570) and the latent position ids (:535-546) against BFL's sampling utilities."""
    sampling = pytest.importorskip("torchtitan.experiments.flux.sampling")
    utils = pytest.importorskip("torchtitan.experiments.flux.utils")
    for n, tokens in ((28, 4096), (4, 256), (30, 9216), (7, 1024)):
        ts, sg = O.make_sigmas(n, tokens)
        ref = torch.tensor(sampling.get_schedule(n, tokens, base_shift=0.5, max_shift=1.15, shift=True))
        assert torch.allclose(sg.double(), ref.double(), atol=2e-6), (n, tokens)
        assert torch.allclose(ts.double(), ref[:-1].double() * 1000, atol=2e-3)
    z = torch.randn(2, 16, 12, 20, generator=torch.Generator().manual_seed(0))
    assert torch.equal(O.pack_latents(z), utils.pack_latents(z))
    assert torch.equal(O.unpack_latents(O.pack_latents(z), 12 * 8, 20 * 8, 16), utils.unpack_latents(utils.pack_latents(z), 12, 20))
    ids = utils.create_position_encoding_for_latents(1, 12, 20)[0]
    assert torch.equal(O.prepare_latent_image_ids(12, 20), ids.to(torch.float32))
