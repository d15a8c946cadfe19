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
