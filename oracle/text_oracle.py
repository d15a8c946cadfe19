"""CPU oracle for the prompt encoders - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU leg may import this module; the product path
(``reptext_b200/``) never does.

The RepText pipelines call ``self.text_encoder_2(input_ids)[0]`` (transformers ``T5EncoderModel``, T5-v1.1-XXL for
FLUX.1-dev; no attention mask: padding tokens attend and are attended to) and ``self.text_encoder(input_ids)
.pooler_output`` (transformers ``CLIPTextModel``, CLIP ViT-L/14) - ``RepText/pipeline_flux_controlnet.py:232-347``.
The arithmetic lives in transformers (pinned here: the version in this image, 5.5.0).  This file restates it over the
state dicts with transformers' parameter names; ``tests/test_text_oracle.py`` PINS it against the real
``transformers`` modules (random-init small configs and a one-layer full-width T5) in this container.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

T5_XXL_CONFIG = dict(vocab_size=32128, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                     relative_attention_num_buckets=32, relative_attention_max_distance=128, layer_norm_epsilon=1e-6)
CLIP_L_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                     num_attention_heads=12, max_position_embeddings=77, layer_norm_eps=1e-5, eos_token_id=2)


# ------------------------------------------------------------------------------------------------ T5 encoder
def t5_relative_buckets(S: int, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """transformers ``T5Attention._relative_position_bucket`` (bidirectional) for relative positions
    key - query in [-(S-1), S-1]: returns [2S-1] bucket ids, index = key - query + S - 1."""
    rel = torch.arange(-(S - 1), S)
    nb = num_buckets // 2
    out = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(rel < max_exact, rel, large)


def t5_position_bias(sd: SD, cfg: dict, S: int) -> Tensor:
    """[heads, S, S] additive bias (``compute_bias``): the table of block 0, shared by every block."""
    lut = t5_relative_buckets(S, cfg["relative_attention_num_buckets"], cfg["relative_attention_max_distance"])
    table = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]        # [buckets, heads]
    i = torch.arange(S)
    idx = lut.to(table.device)[(i[None, :] - i[:, None]) + S - 1]                            # [query, key]
    return table[idx].permute(2, 0, 1)


def t5_layer_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """``T5LayerNorm``: RMS norm, no mean subtraction, no bias, variance in fp32."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(var + eps)).to(x.dtype)


def t5_encoder(sd: SD, cfg: dict, input_ids: Tensor) -> Tensor:
    """``T5EncoderModel(input_ids)[0]``: [B, S] -> [B, S, d_model].  No mask, no dropout, gated-GELU (gelu_new) MLP,
    attention scores NOT scaled, final RMS norm."""
    H, dk, eps = cfg["num_heads"], cfg["d_kv"], cfg["layer_norm_epsilon"]
    x = sd["shared.weight"][input_ids]
    B, S, _ = x.shape
    bias = t5_position_bias(sd, cfg, S).to(x.dtype)
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer."
        h = t5_layer_norm(x, sd[p + "0.layer_norm.weight"], eps)
        q = F.linear(h, sd[p + "0.SelfAttention.q.weight"]).view(B, S, H, dk).transpose(1, 2)
        k = F.linear(h, sd[p + "0.SelfAttention.k.weight"]).view(B, S, H, dk).transpose(1, 2)
        v = F.linear(h, sd[p + "0.SelfAttention.v.weight"]).view(B, S, H, dk).transpose(1, 2)
        w = torch.softmax((q @ k.transpose(-1, -2) + bias[None]).float(), dim=-1).to(x.dtype)
        o = (w @ v).transpose(1, 2).reshape(B, S, H * dk)
        x = x + F.linear(o, sd[p + "0.SelfAttention.o.weight"])
        h = t5_layer_norm(x, sd[p + "1.layer_norm.weight"], eps)
        g = F.gelu(F.linear(h, sd[p + "1.DenseReluDense.wi_0.weight"]), approximate="tanh")        # gelu_new
        x = x + F.linear(g * F.linear(h, sd[p + "1.DenseReluDense.wi_1.weight"]), sd[p + "1.DenseReluDense.wo.weight"])
    return t5_layer_norm(x, sd["encoder.final_layer_norm.weight"], eps)


# ------------------------------------------------------------------------------------------------ CLIP text model
def clip_text(sd: SD, cfg: dict, input_ids: Tensor) -> Tuple[Tensor, Tensor]:
    """``CLIPTextModel(input_ids)`` -> (last_hidden_state [B, S, D], pooler_output [B, D]).  Pre-LN blocks, causal mask,
    q scaled by head_dim ** -0.5, quick-GELU MLP, final LayerNorm; pooled = the state at the EOS token (the arg-max id when
    ``eos_token_id == 2``, the legacy configs FLUX's CLIP-L ships with; else the first ``eos_token_id``)."""
    H, eps = cfg["num_attention_heads"], cfg["layer_norm_eps"]
    P = "text_model."
    B, S = input_ids.shape
    x = sd[P + "embeddings.token_embedding.weight"][input_ids] + sd[P + "embeddings.position_embedding.weight"][:S]
    D = x.shape[-1]
    hd = D // H
    mask = torch.full((S, S), float("-inf"), device=x.device).triu(1)

    def lin(t, name):
        return F.linear(t, sd[name + ".weight"], sd[name + ".bias"])

    def ln(t, name):
        return F.layer_norm(t, (D,), sd[name + ".weight"], sd[name + ".bias"], eps)

    for i in range(cfg["num_hidden_layers"]):
        p = f"{P}encoder.layers.{i}."
        h = ln(x, p + "layer_norm1")
        q = (lin(h, p + "self_attn.q_proj") * hd ** -0.5).view(B, S, H, hd).transpose(1, 2)
        k = lin(h, p + "self_attn.k_proj").view(B, S, H, hd).transpose(1, 2)
        v = lin(h, p + "self_attn.v_proj").view(B, S, H, hd).transpose(1, 2)
        w = torch.softmax((q @ k.transpose(-1, -2)).float() + mask, dim=-1).to(x.dtype)
        x = x + lin((w @ v).transpose(1, 2).reshape(B, S, D), p + "self_attn.out_proj")
        h = lin(ln(x, p + "layer_norm2"), p + "mlp.fc1")
        x = x + lin(h * torch.sigmoid(1.702 * h), p + "mlp.fc2")
    x = ln(x, P + "final_layer_norm")
    if cfg.get("eos_token_id", 2) == 2:
        idx = input_ids.to(torch.int).argmax(dim=-1)
    else:
        idx = (input_ids == cfg["eos_token_id"]).int().argmax(dim=-1)
    return x, x[torch.arange(B, device=x.device), idx.to(x.device)]


# ------------------------------------------------------------------------------------------------ random weights
def t5_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    d, inner, ff = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]
    out = {"shared.weight": (cfg["vocab_size"], d)}
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer."
        for n in ("q", "k", "v"):
            out[p + f"0.SelfAttention.{n}.weight"] = (inner, d)
        out[p + "0.SelfAttention.o.weight"] = (d, inner)
        if i == 0:
            out[p + "0.SelfAttention.relative_attention_bias.weight"] = (cfg["relative_attention_num_buckets"], cfg["num_heads"])
        out[p + "0.layer_norm.weight"] = (d,)
        out[p + "1.DenseReluDense.wi_0.weight"] = (ff, d)
        out[p + "1.DenseReluDense.wi_1.weight"] = (ff, d)
        out[p + "1.DenseReluDense.wo.weight"] = (d, ff)
        out[p + "1.layer_norm.weight"] = (d,)
    out["encoder.final_layer_norm.weight"] = (d,)
    return out


def clip_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    d, ff = cfg["hidden_size"], cfg["intermediate_size"]
    P = "text_model."
    out = {P + "embeddings.token_embedding.weight": (cfg["vocab_size"], d),
           P + "embeddings.position_embedding.weight": (cfg["max_position_embeddings"], d)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"{P}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out[p + f"self_attn.{n}.weight"] = (d, d)
            out[p + f"self_attn.{n}.bias"] = (d,)
        for n in ("layer_norm1", "layer_norm2"):
            out[p + n + ".weight"] = (d,)
            out[p + n + ".bias"] = (d,)
        out[p + "mlp.fc1.weight"], out[p + "mlp.fc1.bias"] = (ff, d), (ff,)
        out[p + "mlp.fc2.weight"], out[p + "mlp.fc2.bias"] = (d, ff), (d,)
    out[P + "final_layer_norm.weight"], out[P + "final_layer_norm.bias"] = (d,), (d,)
    return out


def random_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> SD:
    """Seeded weights at a scale that keeps activations O(1): linears fan-in scaled, norms 1 + 0.1 n, biases 0.02 n,
    embeddings n, relative bias n."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, s in shapes.items():
        if "relative_attention_bias" in k or "embedding" in k or k == "shared.weight":
            t = torch.randn(s, generator=g) * (1.0 if "relative" in k or k == "shared.weight" else 0.5)
        elif k.endswith(".bias"):
            t = 0.02 * torch.randn(s, generator=g)
        elif len(s) == 1:
            t = 1.0 + 0.1 * torch.randn(s, generator=g)
        else:
            t = torch.randn(s, generator=g) * s[1] ** -0.5
            if k.endswith("SelfAttention.q.weight"):
                # T5 does not scale q k^T; its own initialisation (transformers T5PreTrainedModel._init_weights) puts the
                # 1 / sqrt(d_kv) into q's weights instead, which keeps the scores O(1)
                t = t * 64 ** -0.5
        sd[k] = t
    return sd
