"""Model configurations for the RepText hot path.

Field names are the ``register_to_config`` names of the reference
(``RepText/controlnet_flux.py:44-60``) and of diffusers' ``FluxTransformer2DModel``.
"""
from __future__ import annotations

import copy

# FLUX.1-dev transformer
FLUX_DEV = dict(
    patch_size=1, in_channels=64, out_channels=64, num_layers=19, num_single_layers=38,
    attention_head_dim=128, num_attention_heads=24, joint_attention_dim=4096,
    pooled_projection_dim=768, guidance_embeds=True, axes_dims_rope=(16, 56, 56),
)

# Shakker-Labs/RepText ControlNet: 6 double + 0 single blocks, canny+position latents (64+64 features)
REPTEXT_CONTROLNET = dict(
    patch_size=1, in_channels=64, num_layers=6, num_single_layers=0,
    attention_head_dim=128, num_attention_heads=24, joint_attention_dim=4096,
    pooled_projection_dim=768, guidance_embeds=True, axes_dims_rope=(16, 56, 56),
    num_mode=None, extra_conditioning_channels=0, extra_condition_channels=64,
)

# alimama FLUX inpainting ControlNet: masked-image latents (16ch) + mask (1ch) -> 68 packed features
INPAINT_CONTROLNET = dict(REPTEXT_CONTROLNET, extra_condition_channels=4)

# BASELINE.json config 1: tiny random-init pair, CPU-runnable in fp32
TINY_TRANSFORMER = dict(
    patch_size=1, in_channels=64, out_channels=64, num_layers=2, num_single_layers=4,
    attention_head_dim=64, num_attention_heads=4, joint_attention_dim=64,
    pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 24, 24),
)
TINY_CONTROLNET = dict(
    patch_size=1, in_channels=64, num_layers=2, num_single_layers=0,
    attention_head_dim=64, num_attention_heads=4, joint_attention_dim=64,
    pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 24, 24),
    num_mode=None, extra_conditioning_channels=0, extra_condition_channels=64,
)
TINY_INPAINT_CONTROLNET = dict(TINY_CONTROLNET, extra_condition_channels=4)

# a tiny pair whose head_dim is 128 and whose token counts are multiples of 128, so that the
# tcgen05 kernels (not the SIMT ones) run in small bf16 tests
SMALL128_TRANSFORMER = dict(
    patch_size=1, in_channels=64, out_channels=64, num_layers=2, num_single_layers=2,
    attention_head_dim=128, num_attention_heads=2, joint_attention_dim=128,
    pooled_projection_dim=64, guidance_embeds=True, axes_dims_rope=(16, 56, 56),
)
SMALL128_CONTROLNET = dict(
    patch_size=1, in_channels=64, num_layers=2, num_single_layers=0,
    attention_head_dim=128, num_attention_heads=2, joint_attention_dim=128,
    pooled_projection_dim=64, guidance_embeds=True, axes_dims_rope=(16, 56, 56),
    num_mode=None, extra_conditioning_channels=0, extra_condition_channels=64,
)

# eight heads of 128: the smallest pair whose heads shard over 2, 4 and 8 ranks (sequence-parallel tests)
SP8_TRANSFORMER = dict(SMALL128_TRANSFORMER, num_attention_heads=8, num_layers=1, num_single_layers=2)
SP8_CONTROLNET = dict(SMALL128_CONTROLNET, num_attention_heads=8, num_layers=1)
SP8_INPAINT_CONTROLNET = dict(SP8_CONTROLNET, extra_condition_channels=4)

SMALL128_INPAINT_CONTROLNET = dict(SMALL128_CONTROLNET, extra_condition_channels=4)

# FLUX.1-dev scheduler_config.json
SCHEDULER = dict(
    num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True,
    base_shift=0.5, max_shift=1.15, base_image_seq_len=256, max_image_seq_len=4096,
)


def clone(cfg: dict, **overrides) -> dict:
    out = copy.deepcopy(cfg)
    out.update(overrides)
    return out
