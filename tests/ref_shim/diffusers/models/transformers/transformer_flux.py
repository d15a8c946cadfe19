"""diffusers.models.transformers.transformer_flux, as adapters over the Black-Forest-Labs blocks torchtitan ships.

Each block owns nn.Modules with DIFFUSERS' parameter names (so a diffusers-keyed state dict loads with
``load_state_dict``) and evaluates them with BFL's own ``DoubleStreamBlock.forward`` / ``SingleStreamBlock.forward`` /
``LastLayer.forward`` after the weight remap of SURVEY.md A.9.  The model-level loop (``FluxTransformer2DModel.forward``)
restates diffusers 0.36.0, including where the ControlNet residuals are added."""
import numpy as np
import torch
import torch.nn as nn
from torchtitan.experiments.flux.model import layers as bfl

from ...configuration_utils import ConfigMixin, register_to_config
from ..embeddings import (CombinedTimestepGuidanceTextProjEmbeddings, CombinedTimestepTextProjEmbeddings,
                          FluxPosEmbed)
from ..modeling_outputs import Transformer2DModelOutput
from ..modeling_utils import ModelMixin


class _AdaNorm(nn.Module):           # AdaLayerNormZero / AdaLayerNormZeroSingle / AdaLayerNormContinuous: only `.linear`
    def __init__(self, dim, mult):
        super().__init__()
        self.linear = nn.Linear(dim, mult * dim)


class _RMS(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _FluxAttention(nn.Module):
    def __init__(self, dim, head_dim, joint: bool):
        super().__init__()
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.norm_q, self.norm_k = _RMS(head_dim), _RMS(head_dim)
        if joint:
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
            self.norm_added_q, self.norm_added_k = _RMS(head_dim), _RMS(head_dim)
            self.to_add_out = nn.Linear(dim, dim)


class _GELUProj(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner)


class _FeedForward(nn.Module):       # FeedForward(dim, dim_out=dim, activation_fn="gelu-approximate"): net.0.proj, net.2
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])


def _cat(*mods, attr):
    return torch.cat([getattr(m, attr) for m in mods])


def _set_linear(dst: nn.Linear, w, b):
    dst.weight, dst.bias = nn.Parameter(w, requires_grad=False), nn.Parameter(b, requires_grad=False)


def _set_qknorm(norm: "bfl.QKNorm", q: _RMS, k: _RMS):
    norm.query_norm.weight, norm.key_norm.weight = q.weight, k.weight
    norm.query_norm.eps = norm.key_norm.eps = 1e-6          # diffusers: RMSNorm(head_dim, eps=1e-6)


class FluxTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, qk_norm="rms_norm", eps=1e-6):
        super().__init__()
        self.dim, self.heads = dim, num_attention_heads
        self.norm1, self.norm1_context = _AdaNorm(dim, 6), _AdaNorm(dim, 6)
        self.attn = _FluxAttention(dim, attention_head_dim, joint=True)
        self.ff, self.ff_context = _FeedForward(dim), _FeedForward(dim)

    def _bfl(self):
        with torch.device("meta"):
            blk = bfl.DoubleStreamBlock(self.dim, self.heads, mlp_ratio=4.0, qkv_bias=True)
        a = self.attn
        blk.img_mod.lin, blk.txt_mod.lin = self.norm1.linear, self.norm1_context.linear
        _set_linear(blk.img_attn.qkv, _cat(a.to_q, a.to_k, a.to_v, attr="weight"), _cat(a.to_q, a.to_k, a.to_v, attr="bias"))
        _set_linear(blk.txt_attn.qkv, _cat(a.add_q_proj, a.add_k_proj, a.add_v_proj, attr="weight"),
                    _cat(a.add_q_proj, a.add_k_proj, a.add_v_proj, attr="bias"))
        blk.img_attn.proj, blk.txt_attn.proj = a.to_out[0], a.to_add_out
        blk.img_attn.norm, blk.txt_attn.norm = bfl.QKNorm(a.norm_q.weight.shape[0]), bfl.QKNorm(a.norm_q.weight.shape[0])
        _set_qknorm(blk.img_attn.norm, a.norm_q, a.norm_k)
        _set_qknorm(blk.txt_attn.norm, a.norm_added_q, a.norm_added_k)
        blk.img_mlp = nn.Sequential(self.ff.net[0].proj, nn.GELU(approximate="tanh"), self.ff.net[2])
        blk.txt_mlp = nn.Sequential(self.ff_context.net[0].proj, nn.GELU(approximate="tanh"), self.ff_context.net[2])
        return blk

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None):
        img, txt = self._bfl()(hidden_states, encoder_hidden_states, temb, image_rotary_emb)
        if txt.dtype == torch.float16:
            txt = txt.clip(-65504, 65504)
        return txt, img                                      # diffusers returns (encoder_hidden_states, hidden_states)


class FluxSingleTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, mlp_ratio=4.0):
        super().__init__()
        self.dim, self.heads = dim, num_attention_heads
        self.norm = _AdaNorm(dim, 3)
        self.proj_mlp = nn.Linear(dim, int(dim * mlp_ratio))
        self.proj_out = nn.Linear(dim + int(dim * mlp_ratio), dim)
        self.attn = _FluxAttention(dim, attention_head_dim, joint=False)

    def _bfl(self):
        with torch.device("meta"):
            blk = bfl.SingleStreamBlock(self.dim, self.heads, mlp_ratio=4.0)
        a = self.attn
        blk.modulation.lin = self.norm.linear
        _set_linear(blk.linear1, torch.cat([a.to_q.weight, a.to_k.weight, a.to_v.weight, self.proj_mlp.weight]),
                    torch.cat([a.to_q.bias, a.to_k.bias, a.to_v.bias, self.proj_mlp.bias]))
        blk.linear2 = self.proj_out
        blk.norm = bfl.QKNorm(a.norm_q.weight.shape[0])
        _set_qknorm(blk.norm, a.norm_q, a.norm_k)
        return blk

    def forward(self, hidden_states, encoder_hidden_states=None, temb=None, image_rotary_emb=None,
                joint_attention_kwargs=None):
        """diffusers >= 0.35: ``(hidden_states, encoder_hidden_states, temb, rope) -> (encoder_hidden_states,
        hidden_states)``.  With ``encoder_hidden_states=None`` this is the <= 0.34 form the reference's ControlNet still
        calls (controlnet_flux.py:376-380: already-concatenated tokens in, one tensor out); diffusers 0.36.0 itself
        raises a TypeError there, which is why that loop only works with ``num_single_layers=0`` upstream."""
        if encoder_hidden_states is None:
            out = self._bfl()(hidden_states, temb, image_rotary_emb)
            return out.clip(-65504, 65504) if out.dtype == torch.float16 else out
        t = encoder_hidden_states.shape[1]
        out = self._bfl()(torch.cat([encoder_hidden_states, hidden_states], dim=1), temb, image_rotary_emb)
        if out.dtype == torch.float16:
            out = out.clip(-65504, 65504)
        return out[:, :t], out[:, t:]


class FluxTransformer2DModel(ModelMixin, ConfigMixin):
    @register_to_config
    def __init__(self, patch_size=1, in_channels=64, out_channels=None, num_layers=19, num_single_layers=38,
                 attention_head_dim=128, num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768,
                 guidance_embeds=False, axes_dims_rope=(16, 56, 56)):
        super().__init__()
        self.out_channels = out_channels or in_channels
        self.inner_dim = num_attention_heads * attention_head_dim
        self.pos_embed = FluxPosEmbed(theta=10000, axes_dim=axes_dims_rope)
        cls = CombinedTimestepGuidanceTextProjEmbeddings if guidance_embeds else CombinedTimestepTextProjEmbeddings
        self.time_text_embed = cls(embedding_dim=self.inner_dim, pooled_projection_dim=pooled_projection_dim)
        self.context_embedder = nn.Linear(joint_attention_dim, self.inner_dim)
        self.x_embedder = nn.Linear(in_channels, self.inner_dim)
        self.transformer_blocks = nn.ModuleList(
            [FluxTransformerBlock(self.inner_dim, num_attention_heads, attention_head_dim) for _ in range(num_layers)])
        self.single_transformer_blocks = nn.ModuleList(
            [FluxSingleTransformerBlock(self.inner_dim, num_attention_heads, attention_head_dim)
             for _ in range(num_single_layers)])
        self.norm_out = _AdaNorm(self.inner_dim, 2)
        self.proj_out = nn.Linear(self.inner_dim, patch_size * patch_size * self.out_channels)

    def _last_layer(self):
        d = self.inner_dim
        with torch.device("meta"):
            ll = bfl.LastLayer(d, 1, self.out_channels)
        w, b = self.norm_out.linear.weight, self.norm_out.linear.bias
        # AdaLayerNormContinuous chunks (scale, shift); BFL's LastLayer chunks (shift, scale): swap the halves
        lin = nn.Linear(d, 2 * d)
        _set_linear(lin, torch.cat([w[d:], w[:d]]), torch.cat([b[d:], b[:d]]))
        ll.adaLN_modulation = nn.Sequential(nn.SiLU(), lin)
        ll.linear = self.proj_out
        return ll

    def forward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None, img_ids=None,
                txt_ids=None, guidance=None, joint_attention_kwargs=None, controlnet_block_samples=None,
                controlnet_single_block_samples=None, return_dict=True, controlnet_blocks_repeat=False):
        hidden_states = self.x_embedder(hidden_states)
        timestep = timestep.to(hidden_states.dtype) * 1000
        if guidance is not None:
            guidance = guidance.to(hidden_states.dtype) * 1000
        temb = (self.time_text_embed(timestep, pooled_projections) if guidance is None
                else self.time_text_embed(timestep, guidance, pooled_projections))
        encoder_hidden_states = self.context_embedder(encoder_hidden_states)
        if txt_ids.ndim == 3:
            txt_ids = txt_ids[0]
        if img_ids.ndim == 3:
            img_ids = img_ids[0]
        image_rotary_emb = self.pos_embed(torch.cat((txt_ids, img_ids), dim=0))

        for index_block, block in enumerate(self.transformer_blocks):
            encoder_hidden_states, hidden_states = block(hidden_states=hidden_states,
                                                         encoder_hidden_states=encoder_hidden_states, temb=temb,
                                                         image_rotary_emb=image_rotary_emb)
            if controlnet_block_samples is not None:
                interval_control = int(np.ceil(len(self.transformer_blocks) / len(controlnet_block_samples)))
                if controlnet_blocks_repeat:
                    hidden_states = hidden_states + controlnet_block_samples[index_block % len(controlnet_block_samples)]
                else:
                    hidden_states = hidden_states + controlnet_block_samples[index_block // interval_control]

        for index_block, block in enumerate(self.single_transformer_blocks):
            encoder_hidden_states, hidden_states = block(hidden_states=hidden_states,
                                                         encoder_hidden_states=encoder_hidden_states, temb=temb,
                                                         image_rotary_emb=image_rotary_emb)
            if controlnet_single_block_samples is not None:
                interval_control = int(np.ceil(len(self.single_transformer_blocks) / len(controlnet_single_block_samples)))
                hidden_states = hidden_states + controlnet_single_block_samples[index_block // interval_control]

        output = self._last_layer()(hidden_states, temb)
        if not return_dict:
            return (output,)
        return Transformer2DModelOutput(sample=output)
