"""CPU oracle for the FLUX VAE (diffusers ``AutoencoderKL`` with the FLUX.1-dev config) - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU leg may import this module; the product path
(``reptext_b200/``) never does.

SURVEY.md 8(f) ranks the VAE first among the "next" rows: the RepText pipelines call it twice per text line to encode
the Canny / position images (``RepText/pipeline_flux_controlnet.py:705-715``) and once per image to decode
(``:1136-1140``).  The arithmetic lives in diffusers (``AutoencoderKL`` / ``Encoder`` / ``Decoder`` / ``ResnetBlock2D`` /
``Attention`` / ``Downsample2D`` / ``Upsample2D``), which is not installable here; this file restates it over a
state dict with diffusers' parameter names:

    {encoder,decoder}.conv_in, .conv_norm_out, .conv_out
    encoder.down_blocks.{i}.resnets.{j}.{norm1,conv1,norm2,conv2[,conv_shortcut]}, .downsamplers.0.conv
    decoder.up_blocks.{i}.resnets.{j}..., .upsamplers.0.conv
    {encoder,decoder}.mid_block.resnets.{0,1}..., .mid_block.attentions.0.{group_norm,to_q,to_k,to_v,to_out.0}

FLUX.1-dev ``vae/config.json`` (from memory; the oracle itself is pinned against the BFL autoencoder, tests/test_vae_oracle.py): block_out_channels (128, 256, 512, 512),
layers_per_block 2, latent_channels 16, norm_num_groups 32, eps 1e-6, mid-block attention, no quant / post-quant
conv, scaling_factor 0.3611, shift_factor 0.1159.  PINNED against the independent Black-Forest-Labs autoencoder that
torchtitan ships on this box (``tests/test_vae_oracle.py``, weight remap BFL <-> diffusers names).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

FLUX_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, scaling_factor=0.3611, shift_factor=0.1159)


def _gn(sd: SD, name: str, x: Tensor, groups: int) -> Tensor:
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def _conv(sd: SD, name: str, x: Tensor, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def resnet_block(sd: SD, p: str, x: Tensor, groups: int) -> Tensor:
    """diffusers ``ResnetBlock2D`` (no time embedding, output_scale_factor 1): norm-silu-conv twice + shortcut."""
    h = _conv(sd, p + "conv1", F.silu(_gn(sd, p + "norm1", x, groups)))
    h = _conv(sd, p + "conv2", F.silu(_gn(sd, p + "norm2", h, groups)))
    if (p + "conv_shortcut.weight") in sd:
        x = _conv(sd, p + "conv_shortcut", x, padding=0)
    return x + h


def attention_block(sd: SD, p: str, x: Tensor, groups: int) -> Tensor:
    """diffusers ``Attention`` as the VAE mid block builds it: one head over all H*W positions, GroupNorm first,
    residual connection, linear q / k / v / out (the converted 1x1 convolutions of the LDM autoencoder)."""
    b, c, h, w = x.shape
    t = _gn(sd, p + "group_norm", x, groups).flatten(2).transpose(1, 2)          # [B, HW, C]
    q = F.linear(t, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
    k = F.linear(t, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
    v = F.linear(t, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]  # scale 1/sqrt(C)
    o = F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, h, w)


def _mid(sd: SD, p: str, x: Tensor, groups: int) -> Tensor:
    x = resnet_block(sd, p + "mid_block.resnets.0.", x, groups)
    x = attention_block(sd, p + "mid_block.attentions.0.", x, groups)
    return resnet_block(sd, p + "mid_block.resnets.1.", x, groups)


def encode_moments(sd: SD, cfg: dict, x: Tensor) -> Tensor:
    """``AutoencoderKL.encode`` up to the moments: [B, 3, H, W] in [-1, 1] -> [B, 2 * latent, H/8, W/8]."""
    g, nb = cfg["norm_num_groups"], len(cfg["block_out_channels"])
    h = _conv(sd, "encoder.conv_in", x)
    for i in range(nb):
        for j in range(cfg["layers_per_block"]):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}.", h, g)
        if i != nb - 1:  # Downsample2D(padding=0): pad right / bottom by one, 3x3 stride-2 convolution
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _mid(sd, "encoder.", h, g)
    return _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, g)))


def sample_posterior(moments: Tensor, noise: Optional[Tensor] = None, generator=None) -> Tensor:
    """``DiagonalGaussianDistribution.sample``: logvar clamped to [-30, 20]; ``noise`` overrides the draw."""
    mean, logvar = moments.chunk(2, dim=1)
    std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
    if noise is None:
        noise = torch.randn(mean.shape, generator=generator, dtype=mean.dtype, device=mean.device)
    return mean + std * noise


def decode(sd: SD, cfg: dict, z: Tensor) -> Tensor:
    """``AutoencoderKL.decode``: [B, latent, h, w] -> [B, 3, 8h, 8w]."""
    g, nb = cfg["norm_num_groups"], len(cfg["block_out_channels"])
    h = _conv(sd, "decoder.conv_in", z)
    h = _mid(sd, "decoder.", h, g)
    for i in range(nb):
        for j in range(cfg["layers_per_block"] + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}.", h, g)
        if i != nb - 1:  # Upsample2D: nearest x2, then a 3x3 convolution
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, g)))


def encode_for_pipeline(sd: SD, cfg: dict, image: Tensor, noise: Optional[Tensor] = None) -> Tensor:
    """``RepText/pipeline_flux_controlnet.py:705-708``: sample, then (z - shift) * scale."""
    z = sample_posterior(encode_moments(sd, cfg, image), noise)
    return (z - cfg["shift_factor"]) * cfg["scaling_factor"]


def decode_for_pipeline(sd: SD, cfg: dict, latents: Tensor) -> Tensor:
    """``RepText/pipeline_flux_controlnet.py:1137-1139``: z / scale + shift, then decode."""
    return decode(sd, cfg, latents / cfg["scaling_factor"] + cfg["shift_factor"])


def param_shapes(cfg: dict) -> Dict[str, Sequence[int]]:
    """Every parameter of the FLUX-style AutoencoderKL with its shape (diffusers names)."""
    boc, lpb, lat = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]
    out: Dict[str, Sequence[int]] = {}

    def conv(name, cin, cout, k=3):
        out[name + ".weight"] = (cout, cin, k, k)
        out[name + ".bias"] = (cout,)

    def norm(name, c):
        out[name + ".weight"] = (c,)
        out[name + ".bias"] = (c,)

    def resnet(p, cin, cout):
        norm(p + "norm1", cin); conv(p + "conv1", cin, cout)
        norm(p + "norm2", cout); conv(p + "conv2", cout, cout)
        if cin != cout:
            conv(p + "conv_shortcut", cin, cout, 1)

    def mid(p, c):
        resnet(p + "mid_block.resnets.0.", c, c)
        a = p + "mid_block.attentions.0."
        norm(a + "group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            out[a + n + ".weight"] = (c, c)
            out[a + n + ".bias"] = (c,)
        resnet(p + "mid_block.resnets.1.", c, c)

    conv("encoder.conv_in", cfg["in_channels"], boc[0])
    c = boc[0]
    for i, co in enumerate(boc):
        for j in range(lpb):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", c, co)
            c = co
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c)
    mid("encoder.", c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", c, 2 * lat)

    rev = list(reversed(boc))
    conv("decoder.conv_in", lat, rev[0])
    mid("decoder.", rev[0])
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(lpb + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", c, co)
            c = co
        if i != len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c)
    norm("decoder.conv_norm_out", c)
    conv("decoder.conv_out", c, cfg["out_channels"])
    return out


def random_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> SD:
    """Seeded random weights at a scale that keeps activations O(1) through the stack."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for k, shape in param_shapes(cfg).items():
        if k.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif len(shape) == 1:       # GroupNorm weight
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * fan_in ** -0.5
        sd[k] = t.to(dtype)
    return sd
