"""Parameter tables and seeded random initialisation with the diffusers state-dict key names.

There is no network for checkpoints, so every model in this repo is random-init with the
FLUX.1-dev / RepText architecture (SURVEY.md A.8 lists the key names; a real checkpoint with
those names would load through the same ``load_state_dict``).

Init is variance preserving (``std = fan_in ** -0.5``); the "zero" linears of the ControlNet
(``zero_module`` at ``RepText/controlnet_flux.py:101-114``) are re-randomised by default, because a
true zero init makes the ControlNet output identically 0 and would hide every bug in it.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterator, Tuple

import torch


def _double_block(p: str, D: int, hd: int) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    yield p + "norm1.linear.weight", (6 * D, D), "w"
    yield p + "norm1.linear.bias", (6 * D,), "b"
    yield p + "norm1_context.linear.weight", (6 * D, D), "w"
    yield p + "norm1_context.linear.bias", (6 * D,), "b"
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        yield p + f"attn.{n}.weight", (D, D), "w"
        yield p + f"attn.{n}.bias", (D,), "b"
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        yield p + f"attn.{n}.weight", (hd,), "n"
    for ff in ("ff", "ff_context"):
        yield p + f"{ff}.net.0.proj.weight", (4 * D, D), "w"
        yield p + f"{ff}.net.0.proj.bias", (4 * D,), "b"
        yield p + f"{ff}.net.2.weight", (D, 4 * D), "w"
        yield p + f"{ff}.net.2.bias", (D,), "b"


def _single_block(p: str, D: int, hd: int) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    yield p + "norm.linear.weight", (3 * D, D), "w"
    yield p + "norm.linear.bias", (3 * D,), "b"
    for n in ("to_q", "to_k", "to_v"):
        yield p + f"attn.{n}.weight", (D, D), "w"
        yield p + f"attn.{n}.bias", (D,), "b"
    for n in ("norm_q", "norm_k"):
        yield p + f"attn.{n}.weight", (hd,), "n"
    yield p + "proj_mlp.weight", (4 * D, D), "w"
    yield p + "proj_mlp.bias", (4 * D,), "b"
    yield p + "proj_out.weight", (D, 5 * D), "w"
    yield p + "proj_out.bias", (D,), "b"


def param_table(cfg: dict, kind: str) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (key, shape, role) for a ``"transformer"`` or ``"controlnet"``; role in {w, b, n, z}."""
    hd = cfg["attention_head_dim"]
    D = hd * cfg["num_attention_heads"]
    cin = cfg["in_channels"]
    yield "x_embedder.weight", (D, cin), "w"
    yield "x_embedder.bias", (D,), "b"
    yield "context_embedder.weight", (D, cfg["joint_attention_dim"]), "w"
    yield "context_embedder.bias", (D,), "b"
    emb = ["timestep_embedder"] + (["guidance_embedder"] if cfg.get("guidance_embeds") else [])
    for e in emb:
        yield f"time_text_embed.{e}.linear_1.weight", (D, 256), "w"
        yield f"time_text_embed.{e}.linear_1.bias", (D,), "b"
        yield f"time_text_embed.{e}.linear_2.weight", (D, D), "w"
        yield f"time_text_embed.{e}.linear_2.bias", (D,), "b"
    yield "time_text_embed.text_embedder.linear_1.weight", (D, cfg["pooled_projection_dim"]), "w"
    yield "time_text_embed.text_embedder.linear_1.bias", (D,), "b"
    yield "time_text_embed.text_embedder.linear_2.weight", (D, D), "w"
    yield "time_text_embed.text_embedder.linear_2.bias", (D,), "b"
    for i in range(cfg["num_layers"]):
        yield from _double_block(f"transformer_blocks.{i}.", D, hd)
    for j in range(cfg["num_single_layers"]):
        yield from _single_block(f"single_transformer_blocks.{j}.", D, hd)
    if kind == "transformer":
        yield "norm_out.linear.weight", (2 * D, D), "w"
        yield "norm_out.linear.bias", (2 * D,), "b"
        cout = (cfg.get("out_channels") or cin) * cfg.get("patch_size", 1) ** 2
        yield "proj_out.weight", (cout, D), "w"
        yield "proj_out.bias", (cout,), "b"
    elif kind == "controlnet":
        for i in range(cfg["num_layers"]):
            yield f"controlnet_blocks.{i}.weight", (D, D), "z"
            yield f"controlnet_blocks.{i}.bias", (D,), "zb"
        for j in range(cfg["num_single_layers"]):
            yield f"controlnet_single_blocks.{j}.weight", (D, D), "z"
            yield f"controlnet_single_blocks.{j}.bias", (D,), "zb"
        yield "controlnet_x_embedder.weight", (D, cin + cfg.get("extra_condition_channels", 0)), "z"
        yield "controlnet_x_embedder.bias", (D,), "zb"
    else:
        raise ValueError(f"unknown kind {kind!r}")


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def random_state_dict(cfg: dict, kind: str, seed: int = 0, dtype=torch.float32, device="cpu",
                      zero_init: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded random-init state dict.  Each tensor has its own generator seeded from (seed, key), so
    the values do not depend on iteration order.  Values are drawn in fp32 ON THE CPU GENERATOR when
    ``device`` is cpu, and on the device generator otherwise (the two streams differ; a parity test
    builds the dict once and hands the same tensors to both sides)."""
    dev = torch.device(device)
    out: Dict[str, torch.Tensor] = {}
    for key, shape, role in param_table(cfg, kind):
        g = torch.Generator(device=dev)
        g.manual_seed(_key_seed(seed, key))
        if role in ("z", "zb") and zero_init:
            t = torch.zeros(shape, dtype=torch.float32, device=dev)
        elif role in ("w", "z"):
            t = torch.randn(shape, generator=g, dtype=torch.float32, device=dev) * (shape[-1] ** -0.5)
        elif role in ("b", "zb"):
            t = torch.randn(shape, generator=g, dtype=torch.float32, device=dev) * 0.02
        elif role == "n":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32, device=dev)
        else:
            raise AssertionError(role)
        out[key] = t.to(dtype)
    return out


def num_params(cfg: dict, kind: str) -> int:
    n = 0
    for _, shape, _ in param_table(cfg, kind):
        k = 1
        for s in shape:
            k *= s
        n += k
    return n
