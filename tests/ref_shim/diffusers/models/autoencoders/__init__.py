from .autoencoder_kl import AutoencoderKL  # noqa: F401
