"""Inpainting RepText pipeline: drop-in for ``RepText/pipeline_flux_controlnet_inpaint.py`` (class
``FluxControlNetPipeline``, ``__call__`` at ``:846-1313``).

Adds to the text-to-image pipeline: a second ControlNet (``controlnet_inpaint``; condition = masked-image
latents + (1 - mask), 68 packed features), negative prompts with TRUE classifier-free guidance (effective batch 2
with batch-1 latents), and the live glyph-latent init.  ``control_image_inpaint``, ``control_mask_inpaint`` and
``controlnet_conditioning_scale_inpaint`` keep their upstream meaning.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from ._pipeline_common import RepTextPipelineBase, calculate_shift, retrieve_timesteps  # noqa: F401
from .models import FluxControlNetModel, FluxTransformer2DModel  # noqa: F401
from .pipeline_utils import FluxPipelineOutput, VaeImageProcessor  # noqa: F401

DEFAULT_NEGATIVE_PROMPT = "bad quality, worst quality, text, signature, watermark, extra words"


class FluxControlNetPipeline(RepTextPipelineBase):
    model_cpu_offload_seq = "text_encoder->text_encoder_2->transformer->vae"
    _optional_components: List[str] = []
    _inpaint = True

    def __init__(self, scheduler, vae, text_encoder, tokenizer, text_encoder_2, tokenizer_2, transformer, controlnet,
                 controlnet_inpaint):
        super().__init__()
        self.register_modules(vae=vae, text_encoder=text_encoder, text_encoder_2=text_encoder_2, tokenizer=tokenizer,
                              tokenizer_2=tokenizer_2, transformer=transformer, scheduler=scheduler,
                              controlnet=controlnet, controlnet_inpaint=controlnet_inpaint)
        self._setup()
        self.mask_processor = VaeImageProcessor(vae_scale_factor=self.vae_scale_factor, do_resize=True,
                                                do_convert_grayscale=True, do_normalize=False, do_binarize=True)

    def encode_prompt(self, prompt, prompt_2, device=None, num_images_per_prompt: int = 1,
                      do_classifier_free_guidance: bool = True, negative_prompt=None, negative_prompt_2=None,
                      prompt_embeds=None, pooled_prompt_embeds=None, max_sequence_length: int = 512, lora_scale=None,
                      negative_prompt_embeds=None, negative_pooled_prompt_embeds=None):
        """``:333-448``, same positional order.  CLIP encodes ``prompt`` / ``negative_prompt`` (pooled), T5 encodes
        ``prompt_2`` / ``negative_prompt_2`` (``:403-429``).  ``negative_prompt_embeds`` / ``negative_pooled_prompt_embeds``
        (keyword only in practice: they come after every upstream parameter) are an extension so that the pipeline also
        runs without text encoders."""
        device = device or self._execution_device
        if prompt_embeds is None:
            prompt_embeds, pooled_prompt_embeds = self._encode_text(prompt_2 or prompt, num_images_per_prompt,
                                                                    max_sequence_length, clip_prompt=prompt)
        if do_classifier_free_guidance:
            if negative_prompt_embeds is None:
                negative_prompt = negative_prompt or DEFAULT_NEGATIVE_PROMPT
                negative_prompt_2 = negative_prompt_2 or negative_prompt
                n = prompt_embeds.shape[0] // num_images_per_prompt
                rep = lambda p: [p] * n if isinstance(p, str) else list(p)
                negative_prompt_embeds, negative_pooled_prompt_embeds = self._encode_text(
                    rep(negative_prompt_2), num_images_per_prompt, prompt_embeds.shape[1],
                    clip_prompt=rep(negative_prompt))
        else:
            negative_prompt_embeds = negative_pooled_prompt_embeds = None
        return (prompt_embeds, pooled_prompt_embeds, negative_prompt_embeds, negative_pooled_prompt_embeds,
                self._text_ids(prompt_embeds.shape[1], device))

    def prepare_image_with_mask(self, image, mask, width, height, batch_size, num_images_per_prompt, device, dtype,
                                do_classifier_free_guidance=False):
        """``:761-826``: masked image (= -1 inside the mask) -> VAE latents, concat (1 - mask) resized to the latent
        grid (nearest), pack -> [B, N, 68]."""
        if not isinstance(image, torch.Tensor):
            image = self.image_processor.preprocess(image, height=height, width=width)
        repeat_by = batch_size if image.shape[0] == 1 else num_images_per_prompt
        image = image.repeat_interleave(repeat_by, dim=0).to(device=device, dtype=dtype)
        if not isinstance(mask, torch.Tensor):
            mask = self.mask_processor.preprocess(mask, height=height, width=width)
        mask = mask.repeat_interleave(repeat_by, dim=0).to(device=device, dtype=dtype)
        masked = image.clone()
        masked[(mask > 0.5).repeat(1, 3, 1, 1)] = -1
        cfg = self.vae.config
        z = self.vae.encode(masked.to(self.vae.dtype)).latent_dist.sample()
        z = ((z - cfg.shift_factor) * cfg.scaling_factor).to(dtype)
        m = F.interpolate(mask, size=(height // self.vae_scale_factor * 2, width // self.vae_scale_factor * 2))
        control = torch.cat([z, 1 - m], dim=1)
        packed = self._pack_latents(control, batch_size * num_images_per_prompt, control.shape[1], control.shape[2],
                                    control.shape[3])
        if do_classifier_free_guidance:
            packed = torch.cat([packed] * 2)
        return packed, height, width

    @torch.no_grad()
    def __call__(
        self,
        prompt: Union[str, List[str]] = None,
        prompt_2: Optional[Union[str, List[str]]] = None,
        true_guidance_scale: float = 3.5,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        negative_prompt_2: Optional[Union[str, List[str]]] = None,
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
        control_image_inpaint=None,
        control_mask_inpaint=None,
        controlnet_conditioning_scale_inpaint: Union[float, List[float]] = 1.0,
        negative_prompt_embeds: Optional[torch.FloatTensor] = None,
        negative_pooled_prompt_embeds: Optional[torch.FloatTensor] = None,
    ):
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
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
        mv = lambda x: None if x is None else x.to(device=device, dtype=dtype, non_blocking=True)
        do_cfg = self.do_classifier_free_guidance      # guidance_scale > 1 (:241-242), NOT true_guidance_scale
        (prompt_embeds, pooled_prompt_embeds, negative_prompt_embeds, negative_pooled_prompt_embeds,
         text_ids) = self.encode_prompt(
            prompt=prompt, prompt_2=prompt_2, prompt_embeds=mv(prompt_embeds),
            pooled_prompt_embeds=mv(pooled_prompt_embeds), do_classifier_free_guidance=do_cfg,
            negative_prompt=negative_prompt, negative_prompt_2=negative_prompt_2, device=device,
            num_images_per_prompt=num_images_per_prompt, max_sequence_length=max_sequence_length,
            negative_prompt_embeds=mv(negative_prompt_embeds),
            negative_pooled_prompt_embeds=mv(negative_pooled_prompt_embeds))
        if do_cfg:   # :1033-1035 - the latents are NOT doubled (:1145)
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
            pooled_prompt_embeds = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], dim=0)

        self._require_controlnet(self.controlnet, "controlnet")
        self._require_controlnet(self.controlnet_inpaint, "controlnet_inpaint")
        control_image_list = []
        for image_, position_ in zip(control_image, control_position):
            packed, height, width = self.prepare_image(
                image=image_, image_position=position_, width=width, height=height,
                batch_size=batch_size * num_images_per_prompt, num_images_per_prompt=num_images_per_prompt,
                device=device, dtype=dtype, do_classifier_free_guidance=do_cfg)
            control_image_list.append(packed)
        control_image_inpaint, height, width = self.prepare_image_with_mask(
            image=control_image_inpaint, mask=control_mask_inpaint, width=width, height=height,
            batch_size=batch_size * num_images_per_prompt, num_images_per_prompt=num_images_per_prompt, device=device,
            dtype=dtype, do_classifier_free_guidance=do_cfg)

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
            callback_on_step_end_tensor_inputs=callback_on_step_end_tensor_inputs,
            control_image_inpaint=control_image_inpaint,
            controlnet_conditioning_scale_inpaint=controlnet_conditioning_scale_inpaint,
            true_guidance_scale=true_guidance_scale)
        return self._finish(latents, height, width, output_type, return_dict)
