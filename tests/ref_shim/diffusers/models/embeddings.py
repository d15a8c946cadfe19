"""diffusers.models.embeddings for FLUX, as adapters over the Black-Forest-Labs modules torchtitan ships.

Parameter names are diffusers' (``timestep_embedder.linear_1`` ...); the arithmetic is BFL's
(``timestep_embedding``, ``MLPEmbedder``-shaped in -> SiLU -> out, ``EmbedND`` / ``rope``)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchtitan.experiments.flux.model import layers as bfl


class _MLP(nn.Module):
    """diffusers ``TimestepEmbedding`` / ``PixArtAlphaTextProjection(act_fn="silu")``: linear_1 -> SiLU -> linear_2,
    evaluated by BFL's ``MLPEmbedder.forward`` on these parameters."""

    def __init__(self, in_dim, hidden):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, hidden)
        self.linear_2 = nn.Linear(hidden, hidden)

    def forward(self, x):
        m = bfl.MLPEmbedder.__new__(bfl.MLPEmbedder)
        nn.Module.__init__(m)
        m.in_layer, m.silu, m.out_layer = self.linear_1, nn.SiLU(), self.linear_2
        return bfl.MLPEmbedder.forward(m, x)


def _time_proj(t):
    """``Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0)`` == BFL ``timestep_embedding``
    with time_factor 1 (the caller has already multiplied by 1000); computed in fp32 like diffusers."""
    return bfl.timestep_embedding(t.float(), 256, time_factor=1.0)


class CombinedTimestepTextProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.timestep_embedder = _MLP(256, embedding_dim)
        self.text_embedder = _MLP(pooled_projection_dim, embedding_dim)

    def forward(self, timestep, pooled_projection):
        t = self.timestep_embedder(_time_proj(timestep).to(dtype=pooled_projection.dtype))
        return t + self.text_embedder(pooled_projection)


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.timestep_embedder = _MLP(256, embedding_dim)
        self.guidance_embedder = _MLP(256, embedding_dim)
        self.text_embedder = _MLP(pooled_projection_dim, embedding_dim)

    def forward(self, timestep, guidance, pooled_projection):
        t = self.timestep_embedder(_time_proj(timestep).to(dtype=pooled_projection.dtype))
        g = self.guidance_embedder(_time_proj(guidance).to(dtype=pooled_projection.dtype))
        return (t + g) + self.text_embedder(pooled_projection)


class FluxPosEmbed(nn.Module):
    """Returns BFL's rotation-matrix table ``[1, 1, S, hd/2, 2, 2]`` (float32, angles in float64 like diffusers'
    ``freqs_dtype``); the shim's blocks hand it to BFL's ``attention`` unchanged.  The reference only passes the
    object through (controlnet_flux.py:316-317, :347, :379)."""

    def __init__(self, theta, axes_dim):
        super().__init__()
        self.theta, self.axes_dim = theta, list(axes_dim)
        self._nd = bfl.EmbedND(sum(self.axes_dim), theta, self.axes_dim)

    def forward(self, ids):
        return self._nd(ids[None].double())
