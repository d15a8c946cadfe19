"""Text-to-image RepText pipeline: drop-in for ``RepText/pipeline_flux_controlnet.py`` (class
``FluxControlNetPipeline``, ``__call__`` at ``:749-1148``) with the denoising step on the B200 runtime.

Same constructor modules, same ``__call__`` keyword arguments (``control_image`` = per-line Canny glyph images,
``control_position`` = per-line position masks, ``control_mask`` = per-line regional masks, ``control_glyph``,
``controlnet_conditioning_scale``, ``controlnet_conditioning_step``, ...), same return types.  Glyph rendering
(PIL), Canny (cv2) and mask construction stay in the caller's Python, as in ``RepText/infer.py:64-104``.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import numpy as np
import torch

from ._pipeline_common import RepTextPipelineBase, calculate_shift, retrieve_timesteps  # noqa: F401
from .models import FluxControlNetModel, FluxTransformer2DModel  # noqa: F401
from .pipeline_utils import FluxPipelineOutput  # noqa: F401


class FluxControlNetPipeline(RepTextPipelineBase):
    model_cpu_offload_seq = "text_encoder->text_encoder_2->transformer->vae"
    _optional_components: List[str] = []
    _inpaint = False

    def __init__(self, scheduler, vae, text_encoder, tokenizer, text_encoder_2, tokenizer_2, transformer, controlnet):
        super().__init__()
        self.register_modules(vae=vae, text_encoder=text_encoder, text_encoder_2=text_encoder_2, tokenizer=tokenizer,
                              tokenizer_2=tokenizer_2, transformer=transformer, scheduler=scheduler,
                              controlnet=controlnet)
        self._setup()

    def encode_prompt(self, prompt, prompt_2, device=None, num_images_per_prompt: int = 1, prompt_embeds=None,
                      pooled_prompt_embeds=None, max_sequence_length: int = 512, lora_scale=None,
                      get_text_to_render: bool = False):
        """``:349-456``: returns (prompt_embeds [B, L, 4096], pooled [B, 768], text_ids [L, 3] zeros), followed by the T5
        token span (start, end) of the quoted text when ``get_text_to_render`` is set (``:423-430``, ``:453-454``; as
        upstream the span only exists when the prompt is encoded here, not with ``prompt_embeds`` passed in)."""
        device = device or self._execution_device
        span = ()
        if prompt_embeds is None:
            prompt_embeds, pooled_prompt_embeds, *span = self._encode_text(
                prompt_2 or prompt, num_images_per_prompt, max_sequence_length, clip_prompt=prompt,
                get_text_to_render=get_text_to_render)
        elif get_text_to_render:
            raise ValueError("`get_text_to_render` needs the prompt to be encoded here (upstream fails with an unbound "
                             "`t5_start_index` when `prompt_embeds` is passed)")
        return (prompt_embeds, pooled_prompt_embeds, self._text_ids(prompt_embeds.shape[1], device), *span)

    @torch.no_grad()
    def __call__(
        self,
        prompt: Union[str, List[str]] = None,
        prompt_2: Optional[Union[str, List[str]]] = None,
        height: Optional[int] = None,
        width: Optional[int] = None,
        num_inference_steps: int = 28,
        timesteps: List[int] = None,
        guidance_scale: float = 7.0,
        control_guidance_start: Union[float, List[float]] = 0.0,
        control_guidance_end: Union[float, List[float]] = 1.0,
        control_image=None,
        control_mode: Optional[Union[int, List[int]]] = None,
        controlnet_conditioning_scale: Union[float, List[float]] = 1.0,
        controlnet_conditioning_step: int = 30,
        num_images_per_prompt: Optional[int] = 1,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        latents: Optional[torch.FloatTensor] = None,
        prompt_embeds: Optional[torch.FloatTensor] = None,
        pooled_prompt_embeds: Optional[torch.FloatTensor] = None,
        output_type: Optional[str] = "pil",
        return_dict: bool = True,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
        callback_on_step_end_tensor_inputs: List[str] = ["latents"],
        max_sequence_length: int = 512,
        control_mask=None,
        control_position=None,
        control_glyph=None,
    ):
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        # control_guidance_start / control_guidance_end are accepted and, as upstream (:999-1005), never used.
        self.check_inputs(prompt, prompt_2, height, width, prompt_embeds=prompt_embeds,
                          pooled_prompt_embeds=pooled_prompt_embeds,
                          callback_on_step_end_tensor_inputs=callback_on_step_end_tensor_inputs,
                          max_sequence_length=max_sequence_length)
        self._guidance_scale = guidance_scale
        self._joint_attention_kwargs = joint_attention_kwargs
        self._interrupt = False

        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self._execution_device
        dtype = self.transformer.dtype
        if prompt_embeds is not None:
            prompt_embeds = prompt_embeds.to(device=device, dtype=dtype, non_blocking=True)
            pooled_prompt_embeds = pooled_prompt_embeds.to(device=device, dtype=dtype, non_blocking=True)
        prompt_embeds, pooled_prompt_embeds, text_ids = self.encode_prompt(
            prompt=prompt, prompt_2=prompt_2, prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds,
            device=device, num_images_per_prompt=num_images_per_prompt, max_sequence_length=max_sequence_length)

        self._require_controlnet(self.controlnet, "controlnet")
        control_image_list = []
        for image_, position_ in zip(control_image, control_position):
            packed, height, width = self.prepare_image(
                image=image_, image_position=position_, width=width, height=height,
                batch_size=batch_size * num_images_per_prompt, num_images_per_prompt=num_images_per_prompt,
                device=device, dtype=dtype)
            control_image_list.append(packed)

        num_channels_latents = self.transformer.config.in_channels // 4
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        image_seq_len = (int(height) // self.vae_scale_factor) * (int(width) // self.vae_scale_factor)
        sc = self.scheduler.config
        mu = calculate_shift(image_seq_len, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps,
                                                            sigmas, mu=mu)
        if control_glyph is not None:
            init_image = self.image_processor.preprocess(control_glyph, height=height, width=width).to(torch.float32)
            latents, latent_image_ids = self.prepare_latents_reptext(
                init_image, batch_size * num_images_per_prompt, num_channels_latents, height, width,
                prompt_embeds.dtype, device, generator, None)
        else:
            latents, latent_image_ids = self.prepare_latents(
                batch_size * num_images_per_prompt, num_channels_latents, height, width, prompt_embeds.dtype, device,
                generator, latents)
        self._num_timesteps = len(timesteps)
        control_mask_list = self._regional_masks(control_mask, latents.device, latents.dtype)

        latents = self._denoise(
            latents=latents, latent_image_ids=latent_image_ids, text_ids=text_ids, prompt_embeds=prompt_embeds,
            pooled_prompt_embeds=pooled_prompt_embeds, timesteps=timesteps, num_inference_steps=num_inference_steps,
            guidance_scale=guidance_scale, control_image_list=control_image_list, control_mask_list=control_mask_list,
            control_mode=control_mode, controlnet_conditioning_scale=controlnet_conditioning_scale,
            controlnet_conditioning_step=controlnet_conditioning_step, callback_on_step_end=callback_on_step_end,
            callback_on_step_end_tensor_inputs=callback_on_step_end_tensor_inputs)
        return self._finish(latents, height, width, output_type, return_dict)
