"""reptext_b200 — B200-native RepText denoising step behind the reference's pipeline API."""
from . import config, weights  # noqa: F401

__all__ = ["config", "weights"]
