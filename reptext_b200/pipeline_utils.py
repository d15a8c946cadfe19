"""Minimal host-side pieces the reference pipelines take from diffusers (which is not installable here):
a ``DiffusionPipeline`` base (``register_modules``, ``_execution_device``, ``progress_bar``,
``maybe_free_model_hooks``), ``VaeImageProcessor`` (PIL / numpy / tensor -> [-1, 1] NCHW and back),
``randn_tensor``, and stand-ins for the modules OUTSIDE the hot path - the VAE and the CLIP / T5 text
encoders (SURVEY.md section 8f "next") - so that the reference-shaped ``__call__`` runs end to end on synthetic
data.  The stand-ins are deterministic and cheap; they make no claim to be FLUX's VAE or T5.
"""
from __future__ import annotations

import contextlib
import hashlib
import warnings
from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F

from .models import FrozenConfig

try:  # PIL is host-side glyph rendering's library; present in this image
    import PIL.Image
except Exception:  # pragma: no cover
    PIL = None


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: draw on the generator's device, then move."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    if isinstance(generator, (list, tuple)):
        shape1 = (1,) + tuple(shape[1:])
        parts = [randn_tensor(shape1, g, device, dtype) for g in generator]
        return torch.cat(parts, dim=0)
    gdev = generator.device if generator is not None else device
    t = torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype)
    return t.to(device)


@dataclass
class FluxPipelineOutput:
    images: Union[List, np.ndarray, torch.Tensor]


class _Progress:
    def __init__(self, total, disable):
        self.total, self.n, self.disable = total, 0, disable

    def update(self, k=1):
        self.n += k

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class DiffusionPipeline:
    """The slice of diffusers' DiffusionPipeline the reference touches (pipeline_flux_controlnet.py:209-218,
    :906, :1016, :1143)."""

    def __init__(self):
        self._modules_registered: List[str] = []
        self._progress_bar_config = {}

    def register_modules(self, **modules):
        for name, m in modules.items():
            setattr(self, name, m)
            self._modules_registered.append(name)

    @property
    def components(self):
        return {k: getattr(self, k) for k in self._modules_registered}

    @property
    def _execution_device(self) -> torch.device:
        for name in ("transformer", "controlnet"):
            m = getattr(self, name, None)
            if m is not None and hasattr(m, "device"):
                return torch.device(m.device)
        return torch.device("cuda")

    @property
    def device(self):
        return self._execution_device

    def to(self, *args, **kwargs):
        for name in self._modules_registered:
            m = getattr(self, name)
            if hasattr(m, "to"):
                m.to(*args, **kwargs)
        return self

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def progress_bar(self, iterable=None, total=None):
        return _Progress(total, self._progress_bar_config.get("disable", True))

    def maybe_free_model_hooks(self):
        pass

    def enable_model_cpu_offload(self, *a, **k):
        raise NotImplementedError("weights stay resident in HBM (180 GB per B200); CPU offload is not provided")


class VaeImageProcessor:
    """diffusers' ``VaeImageProcessor`` on the calls the RepText pipelines make (``image_processor``: RGB in [-1, 1];
    ``mask_processor``: grayscale, binarised, not normalised).  Semantics kept from diffusers 0.36:

    * the target size is rounded DOWN to a multiple of ``vae_scale_factor`` (16 here, so 1000 x 760 -> 992 x 752);
    * PIL inputs are resized with PIL's Lanczos filter, tensors and numpy arrays with ``F.interpolate`` (nearest);
    * numpy arrays are taken as HWC (or NHWC) floats ALREADY in [0, 1]; tensors as CHW (or NCHW) in [0, 1]; a tensor
      with 4 channels is taken for latents and returned untouched; an input that already has negative values is not
      normalised again."""

    def __init__(self, vae_scale_factor: int = 8, do_resize: bool = True, do_normalize: bool = True,
                 do_binarize: bool = False, do_convert_grayscale: bool = False, do_convert_rgb: bool = False,
                 vae_latent_channels: int = 4):
        if do_convert_rgb and do_convert_grayscale:
            raise ValueError("`do_convert_rgb` and `do_convert_grayscale` can not both be set to `True`")
        self.config = FrozenConfig(vae_scale_factor=vae_scale_factor, do_resize=do_resize, do_normalize=do_normalize,
                                   do_binarize=do_binarize, do_convert_grayscale=do_convert_grayscale,
                                   do_convert_rgb=do_convert_rgb, vae_latent_channels=vae_latent_channels,
                                   resample="lanczos")

    def _target(self, height, width, default_h, default_w):
        f = self.config.vae_scale_factor
        height = default_h if height is None else height
        width = default_w if width is None else width
        return height - height % f, width - width % f

    def _from_pil(self, imgs, height, width) -> torch.Tensor:
        c = self.config
        if c.do_resize:
            height, width = self._target(height, width, imgs[0].height, imgs[0].width)
            imgs = [im.resize((width, height), resample=PIL.Image.Resampling.LANCZOS) for im in imgs]
        if c.do_convert_rgb:
            imgs = [im.convert("RGB") for im in imgs]
        elif c.do_convert_grayscale:
            imgs = [im.convert("L") for im in imgs]
        arr = np.stack([np.array(im).astype(np.float32) / 255.0 for im in imgs], axis=0)
        if arr.ndim == 3:
            arr = arr[..., None]
        return torch.from_numpy(arr.transpose(0, 3, 1, 2))

    def preprocess(self, image, height: Optional[int] = None, width: Optional[int] = None) -> torch.Tensor:
        c = self.config
        kinds = (np.ndarray, torch.Tensor) if PIL is None else (PIL.Image.Image, np.ndarray, torch.Tensor)
        if c.do_convert_grayscale and isinstance(image, (torch.Tensor, np.ndarray)) and image.ndim == 3:
            image = image.unsqueeze(1) if isinstance(image, torch.Tensor) else np.expand_dims(image, axis=-1)
        imgs = [image] if isinstance(image, kinds) else image
        if not (isinstance(imgs, (list, tuple)) and len(imgs) > 0 and all(isinstance(i, kinds) for i in imgs)):
            raise ValueError(f"unsupported image input {type(image)}: expected PIL images, numpy arrays or tensors")
        imgs = list(imgs)
        first = imgs[0]
        if PIL is not None and isinstance(first, PIL.Image.Image):
            x = self._from_pil(imgs, height, width)
        else:
            if isinstance(first, np.ndarray):
                arr = np.concatenate(imgs, axis=0) if first.ndim == 4 else np.stack(imgs, axis=0)
                if arr.ndim == 3:
                    arr = arr[..., None]
                x = torch.from_numpy(np.ascontiguousarray(arr.transpose(0, 3, 1, 2)))
            else:
                x = torch.cat(imgs, dim=0) if first.ndim == 4 else torch.stack(imgs, dim=0)
                if c.do_convert_grayscale and x.ndim == 3:
                    x = x.unsqueeze(1)
                if x.ndim == 4 and x.shape[1] == c.vae_latent_channels:          # latents: nothing to do
                    return x
            if x.ndim != 4:
                raise ValueError(f"image tensors must be CHW or NCHW (HWC / NHWC for numpy), got {x.ndim} dimensions")
            if c.do_resize:
                height, width = self._target(height, width, x.shape[2], x.shape[3])
                x = F.interpolate(x, size=(height, width))
        normalize = c.do_normalize
        if normalize and x.min() < 0:
            warnings.warn("image values are already in [-1, 1]; expected [0, 1] - not normalising again", FutureWarning)
            normalize = False
        if normalize:
            x = 2.0 * x - 1.0
        if c.do_binarize:
            x = (x >= 0.5).to(x.dtype)
        return x

    def postprocess(self, image: torch.Tensor, output_type: str = "pil"):
        if output_type in ("latent", "pt"):
            return image
        x = (image.float() / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).cpu().numpy()
        if output_type == "np":
            return x
        if output_type == "pil":
            return [PIL.Image.fromarray((im * 255).round().astype("uint8").squeeze()) for im in x]
        raise ValueError(f"unknown output_type {output_type}")


# ------------------------------------------------------------------------------------------------------
# Stand-ins for modules outside the hot path
# ------------------------------------------------------------------------------------------------------
class _LatentDist:
    def __init__(self, mean: torch.Tensor, std: float):
        self.mean, self.std = mean, std

    def sample(self, generator=None) -> torch.Tensor:
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class SyntheticVAE:
    """Stand-in with AutoencoderKL's interface (``encode(x).latent_dist.sample()``, ``decode(z)``,
    ``config.{shift_factor, scaling_factor, block_out_channels}``, ``dtype``): 8x average pooling and a fixed
    3 -> 16 channel mix.  NOT the FLUX VAE (out of scope)."""

    def __init__(self, dtype=torch.bfloat16, device="cuda", latent_channels: int = 16, posterior_std: float = 0.0):
        self.dtype, self.device = dtype, torch.device(device)
        self.config = FrozenConfig(shift_factor=0.1159, scaling_factor=0.3611, block_out_channels=(128, 256, 512, 512),
                                   latent_channels=latent_channels)
        g = torch.Generator().manual_seed(1234)
        self._mix = torch.randn(latent_channels, 3, generator=g).to(self.device)
        self._std = posterior_std

    def to(self, *a, **k):
        return self

    def encode(self, x: torch.Tensor):
        x = x.to(self.device, torch.float32)
        if x.shape[1] == 1:
            x = x.repeat(1, 3, 1, 1)
        p = F.avg_pool2d(x, 8)
        z = torch.einsum("oc,bchw->bohw", self._mix, p)
        return FrozenConfig(latent_dist=_LatentDist(z.to(self.dtype), self._std))

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        z = z.to(self.device, torch.float32)
        img = torch.einsum("oc,bohw->bchw", self._mix, z) / self._mix.shape[0]
        img = F.interpolate(img, scale_factor=8, mode="nearest").clamp(-1, 1).to(self.dtype)
        return (img,) if not return_dict else FrozenConfig(sample=img)


class SyntheticTextEncoders:
    """Stand-in for CLIP + T5: prompt -> deterministic pseudo-random (prompt_embeds [B, L, joint_dim],
    pooled [B, pooled_dim]) seeded by a hash of the prompt.  NOT a language model (out of scope, 8f.3)."""

    def __init__(self, joint_attention_dim: int, pooled_projection_dim: int, dtype=torch.bfloat16, device="cuda"):
        self.joint, self.pooled = joint_attention_dim, pooled_projection_dim
        self.dtype, self.device = dtype, torch.device(device)

    def to(self, *a, **k):
        return self

    def _gen(self, text: str) -> torch.Generator:
        seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:7], "little")
        return torch.Generator().manual_seed(seed)

    def encode(self, prompts: Sequence[str], max_sequence_length: int):
        pe, po = [], []
        for p in prompts:
            g = self._gen(p)
            pe.append(torch.randn(max_sequence_length, self.joint, generator=g))
            po.append(torch.randn(self.pooled, generator=g))
        return (torch.stack(pe).to(self.device, self.dtype), torch.stack(po).to(self.device, self.dtype))


class _TokenizerOutput(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


class SyntheticTokenizer:
    """Stand-in for the CLIP BPE / T5 sentencepiece tokenizers (their vocabulary files are checkpoints; there is no
    network): whitespace-split words hashed into the id range, with the framing of the real ones - ``kind="clip"``:
    BOS, words, EOS (the LARGEST id, as in CLIP's vocabulary, which ``pooler_output`` relies on), padded with the
    EOS / pad id; ``kind="t5"``: words, EOS = 1, padded with 0.  Same call signature and ``.input_ids`` result as a
    transformers tokenizer.  NOT a language tokenizer."""

    def __init__(self, kind: str, vocab_size: int, model_max_length: int):
        assert kind in ("clip", "t5")
        self.kind, self.vocab_size, self.model_max_length = kind, vocab_size, model_max_length
        if kind == "clip":
            self.bos_token_id, self.eos_token_id, self.pad_token_id = vocab_size - 2, vocab_size - 1, vocab_size - 1
        else:
            self.bos_token_id, self.eos_token_id, self.pad_token_id = None, 1, 0

    def _words(self, text: str) -> List[int]:
        lo, hi = (0, self.vocab_size - 2) if self.kind == "clip" else (2, self.vocab_size)
        out = []
        for w in text.split():
            h = int.from_bytes(hashlib.sha256(w.encode("utf-8")).digest()[:6], "little")
            out.append(lo + h % (hi - lo))
        return out

    def __call__(self, text, padding="longest", max_length=None, truncation=False, return_tensors="pt", **kw):
        texts = [text] if isinstance(text, str) else list(text)
        rows = []
        for t in texts:
            ids = self._words(t)
            ids = ([self.bos_token_id] if self.kind == "clip" else []) + ids + [self.eos_token_id]
            if truncation and max_length is not None and len(ids) > max_length:
                ids = ids[:max_length - 1] + [self.eos_token_id]
            rows.append(ids)
        width = max_length if padding == "max_length" and max_length is not None else max(len(r) for r in rows)
        rows = [r + [self.pad_token_id] * (width - len(r)) for r in rows]
        return _TokenizerOutput(input_ids=torch.tensor(rows, dtype=torch.long))

    def batch_decode(self, ids, **kw):
        return ["<%d tokens>" % len(r) for r in ids]


@contextlib.contextmanager
def no_grad():
    with torch.no_grad():
        yield
