"""AutoencoderKL with diffusers' interface (``encode(x).latent_dist.sample()``, ``decode(z, return_dict=False)[0]``,
``config.{shift_factor, scaling_factor, block_out_channels}``) over the Black-Forest-Labs autoencoder torchtitan ships.
The FLUX VAE has no quant / post-quant convolutions, so ``encoder(x)`` IS the moments tensor."""
from dataclasses import dataclass

import torch
from torchtitan.experiments.flux.model import autoencoder as bfl_ae

from ...configuration_utils import ConfigMixin, FrozenDict
from ...utils import BaseOutput
from ...utils.torch_utils import randn_tensor
from ..modeling_utils import ModelMixin


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.parameters.device,
                             dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


@dataclass
class AutoencoderKLOutput(BaseOutput):
    latent_dist: DiagonalGaussianDistribution = None


@dataclass
class DecoderOutput(BaseOutput):
    sample: torch.Tensor = None


class AutoencoderKL(ModelMixin, ConfigMixin):
    def __init__(self, block_out_channels=(128, 256, 512, 512), latent_channels=16, layers_per_block=2,
                 scaling_factor=0.3611, shift_factor=0.1159, in_channels=3, out_channels=3):
        super().__init__()
        ch = block_out_channels[0]
        assert all(c % ch == 0 for c in block_out_channels)
        self.ae = bfl_ae.AutoEncoder(bfl_ae.AutoEncoderParams(
            resolution=256, in_channels=in_channels, ch=ch, out_ch=out_channels,
            ch_mult=tuple(c // ch for c in block_out_channels), num_res_blocks=layers_per_block,
            z_channels=latent_channels, scale_factor=scaling_factor, shift_factor=shift_factor))
        self._internal_dict = FrozenDict(block_out_channels=tuple(block_out_channels), latent_channels=latent_channels,
                                         layers_per_block=layers_per_block, scaling_factor=scaling_factor,
                                         shift_factor=shift_factor, in_channels=in_channels, out_channels=out_channels,
                                         use_quant_conv=False, use_post_quant_conv=False)

    def encode(self, x, return_dict=True):
        post = DiagonalGaussianDistribution(self.ae.encoder(x))
        return AutoencoderKLOutput(latent_dist=post) if return_dict else (post,)

    def decode(self, z, return_dict=True, generator=None):
        img = self.ae.decoder(z)
        return DecoderOutput(sample=img) if return_dict else (img,)
