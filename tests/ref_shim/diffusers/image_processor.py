"""VaeImageProcessor (diffusers 0.36.0 ``image_processor.py``), the paths the reference pipelines reach:
``preprocess`` of PIL / numpy / tensor inputs (resize to a multiple of ``vae_scale_factor``: PIL with Lanczos, tensors
with ``F.interpolate`` nearest; [0, 1] -> [-1, 1]; optional grayscale + binarize for masks) and ``postprocess``."""
import warnings
from typing import List, Union

import numpy as np
import PIL.Image
import torch

from .configuration_utils import ConfigMixin, register_to_config

PipelineImageInput = Union[PIL.Image.Image, np.ndarray, torch.Tensor, List[PIL.Image.Image], List[np.ndarray],
                           List[torch.Tensor]]
_RESAMPLE = {"linear": PIL.Image.Resampling.BILINEAR, "bilinear": PIL.Image.Resampling.BILINEAR,
             "bicubic": PIL.Image.Resampling.BICUBIC, "lanczos": PIL.Image.Resampling.LANCZOS,
             "nearest": PIL.Image.Resampling.NEAREST}


class VaeImageProcessor(ConfigMixin):
    @register_to_config
    def __init__(self, do_resize=True, vae_scale_factor=8, vae_latent_channels=4, resample="lanczos",
                 reducing_gap=None, do_normalize=True, do_binarize=False, do_convert_rgb=False,
                 do_convert_grayscale=False):
        super().__init__()
        if do_convert_rgb and do_convert_grayscale:
            raise ValueError("`do_convert_rgb` and `do_convert_grayscale` can not both be set to `True`")

    @staticmethod
    def numpy_to_pil(images):
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        if images.shape[-1] == 1:
            return [PIL.Image.fromarray(image.squeeze(), mode="L") for image in images]
        return [PIL.Image.fromarray(image) for image in images]

    @staticmethod
    def pil_to_numpy(images):
        if not isinstance(images, list):
            images = [images]
        return np.stack([np.array(image).astype(np.float32) / 255.0 for image in images], axis=0)

    @staticmethod
    def numpy_to_pt(images):
        if images.ndim == 3:
            images = images[..., None]
        return torch.from_numpy(images.transpose(0, 3, 1, 2))

    @staticmethod
    def pt_to_numpy(images):
        return images.cpu().permute(0, 2, 3, 1).float().numpy()

    @staticmethod
    def normalize(images):
        return 2.0 * images - 1.0

    @staticmethod
    def denormalize(images):
        return (images * 0.5 + 0.5).clamp(0, 1)

    def get_default_height_width(self, image, height=None, width=None):
        if height is None:
            height = image.height if isinstance(image, PIL.Image.Image) else (
                image.shape[2] if isinstance(image, torch.Tensor) else image.shape[1])
        if width is None:
            width = image.width if isinstance(image, PIL.Image.Image) else (
                image.shape[3] if isinstance(image, torch.Tensor) else image.shape[2])
        width, height = (x - x % self.config.vae_scale_factor for x in (width, height))
        return height, width

    def resize(self, image, height, width, resize_mode="default"):
        if isinstance(image, PIL.Image.Image):
            return image.resize((width, height), resample=_RESAMPLE[self.config.resample],
                                reducing_gap=self.config.reducing_gap)
        if isinstance(image, torch.Tensor):
            return torch.nn.functional.interpolate(image, size=(height, width))
        image = self.numpy_to_pt(image)
        image = torch.nn.functional.interpolate(image, size=(height, width))
        return self.pt_to_numpy(image)

    def binarize(self, image):
        image[image < 0.5] = 0
        image[image >= 0.5] = 1
        return image

    def preprocess(self, image, height=None, width=None, resize_mode="default", crops_coords=None):
        supported = (PIL.Image.Image, np.ndarray, torch.Tensor)
        if self.config.do_convert_grayscale and isinstance(image, (torch.Tensor, np.ndarray)) and image.ndim == 3:
            image = image.unsqueeze(1) if isinstance(image, torch.Tensor) else np.expand_dims(image, axis=-1)
        if isinstance(image, list) and isinstance(image[0], np.ndarray) and image[0].ndim == 4:
            image = np.concatenate(image, axis=0)
        if isinstance(image, list) and isinstance(image[0], torch.Tensor) and image[0].ndim == 4:
            image = torch.cat(image, axis=0)
        if isinstance(image, supported):
            image = [image]
        elif not (isinstance(image, list) and all(isinstance(i, supported) for i in image)):
            raise ValueError(f"Input is in incorrect format. Currently, we only support {', '.join(str(x) for x in supported)}")

        if isinstance(image[0], PIL.Image.Image):
            if self.config.do_resize:
                height, width = self.get_default_height_width(image[0], height, width)
                image = [self.resize(i, height, width, resize_mode=resize_mode) for i in image]
            if self.config.do_convert_rgb:
                image = [i.convert("RGB") for i in image]
            elif self.config.do_convert_grayscale:
                image = [i.convert("L") for i in image]
            image = self.numpy_to_pt(self.pil_to_numpy(image))
        elif isinstance(image[0], np.ndarray):
            image = np.concatenate(image, axis=0) if image[0].ndim == 4 else np.stack(image, axis=0)
            image = self.numpy_to_pt(image)
            height, width = self.get_default_height_width(image, height, width)
            if self.config.do_resize:
                image = self.resize(image, height, width)
        elif isinstance(image[0], torch.Tensor):
            image = torch.cat(image, axis=0) if image[0].ndim == 4 else torch.stack(image, axis=0)
            if self.config.do_convert_grayscale and image.ndim == 3:
                image = image.unsqueeze(1)
            if image.shape[1] == self.config.vae_latent_channels:     # latents need no preprocessing
                return image
            height, width = self.get_default_height_width(image, height, width)
            if self.config.do_resize:
                image = self.resize(image, height, width)

        do_normalize = self.config.do_normalize
        if do_normalize and image.min() < 0:
            warnings.warn("Passing `image` as torch tensor with value range in [-1,1] is deprecated. The expected value "
                          f"range for image tensor is [0,1] when passing as pytorch tensor or numpy Array. You passed "
                          f"`image` with value range [{image.min()},{image.max()}]", FutureWarning)
            do_normalize = False
        if do_normalize:
            image = self.normalize(image)
        if self.config.do_binarize:
            image = self.binarize(image)
        return image

    def postprocess(self, image, output_type="pil", do_denormalize=None):
        if output_type == "latent":
            return image
        if do_denormalize is None:
            do_denormalize = [self.config.do_normalize] * image.shape[0]
        image = torch.stack([self.denormalize(image[i]) if do_denormalize[i] else image[i] for i in range(image.shape[0])])
        if output_type == "pt":
            return image
        image = self.pt_to_numpy(image)
        if output_type == "np":
            return image
        return self.numpy_to_pil(image)
