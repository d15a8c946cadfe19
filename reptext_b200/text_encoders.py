"""The two prompt encoders of the RepText pipelines on the C-ABI runtime: drop-ins for
transformers' ``T5EncoderModel`` (``text_encoder_2``: T5-v1.1-XXL, ``RepText/pipeline_flux_controlnet.py:289-291``) and
``CLIPTextModel`` (``text_encoder``: CLIP ViT-L/14, ``:330-333``) on exactly the calls the pipelines make:

    prompt_embeds = text_encoder_2(input_ids, output_hidden_states=False)[0]        # [B, 512, 4096]
    pooled        = text_encoder(input_ids, output_hidden_states=False).pooler_output  # [B, 768]

State dicts use transformers' parameter names.  The projections and MLPs run on the tcgen05 GEMM (``rt_gemm``): q | k | v as
three column segments of ONE launch, T5's ``wi_0`` (GELU-tanh epilogue) | ``wi_1`` as two segments of one launch, every
output projection with the residual add fused (in place on the residual stream).  Norms, the head_dim-64 attention with
T5's relative-position bias / CLIP's causal mask, the gated activation and the embedding gather are the kernels of
``csrc/text_kernels.cu``.  bf16 storage, fp32 accumulation.  No CPU path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import _lib as L
from . import ops
from .models import FrozenConfig

T5_XXL_CONFIG = dict(vocab_size=32128, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                     relative_attention_num_buckets=32, relative_attention_max_distance=128, layer_norm_epsilon=1e-6,
                     feed_forward_proj="gated-gelu")
CLIP_L_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                     num_attention_heads=12, max_position_embeddings=77, layer_norm_eps=1e-5, eos_token_id=2,
                     hidden_act="quick_gelu")


class _Encoder:
    def __init__(self, config: dict, defaults: dict, state_dict: Dict[str, torch.Tensor], dtype, device):
        if dtype != torch.bfloat16:
            raise ValueError("the prompt-encoder path computes in bfloat16 (fp32 accumulation)")
        cfg = dict(defaults)
        cfg.update(config or {})
        self.config = FrozenConfig(**cfg)
        self.dtype, self.device = dtype, torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("reptext_b200 needs a CUDA device (there is no CPU path)")
        L.lib()
        self._w: Dict[str, torch.Tensor] = {}
        for k, v in state_dict.items():
            self._w[k] = v.detach().to(self.device, self.dtype).contiguous()

    _class_name = ""
    _shapes = staticmethod(lambda cfg: {})

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, torch_dtype=torch.bfloat16,
                        device="cuda", variant: Optional[str] = None, **unused):
        """transformers' layout (``config.json`` + ``model.safetensors`` or its shards; ``black-forest-labs/FLUX.1-dev``'s
        ``text_encoder/`` and ``text_encoder_2/``) from a LOCAL directory (:mod:`reptext_b200.checkpoint`).  Every parameter
        the encoder reads must be there with its shape; extra tensors (a T5 decoder, a CLIP vision tower) are ignored."""
        from . import checkpoint as ck
        d = ck.resolve_dir(pretrained_model_name_or_path, subfolder)
        raw, stored_cls = ck.read_config(d)
        if stored_cls is not None and cls._class_name not in stored_cls and stored_cls not in cls._also:
            raise ValueError(f"{d!r} holds a {stored_cls}, not a {cls._class_name}")
        sd = ck.load_state_dict(d, (ck.TRANSFORMERS_STEM,), variant)
        if "shared.weight" not in sd and "encoder.embed_tokens.weight" in sd:
            sd["shared.weight"] = sd["encoder.embed_tokens.weight"]
        cfg = dict(cls._defaults)
        cfg.update({k: v for k, v in raw.items() if k in cls._defaults})
        want = cls._shapes(cfg)
        bad = [k for k in want if k not in sd or tuple(sd[k].shape) != tuple(want[k])]
        if bad:
            raise RuntimeError(f"{d!r}: parameters missing or of the wrong shape for this config: {bad[:4]}")
        return cls(cfg, {k: sd[k] for k in want}, dtype=torch_dtype, device=device)

    def save_pretrained(self, save_directory) -> None:
        from . import checkpoint as ck
        cfg = {k: getattr(self.config, k) for k in self._defaults}
        ck.write_config(save_directory, cfg, self._class_name, transformers=True)
        ck.save_state_dict(save_directory, {k: v for k, v in self._w.items() if k in self._shapes(cfg)}, ck.TRANSFORMERS_STEM)

    def _p(self, name: str) -> torch.Tensor:
        try:
            return self._w[name]
        except KeyError:
            raise ValueError(f"{type(self).__name__} state dict has no '{name}'") from None

    def to(self, *a, **k):
        return self

    def _ids(self, input_ids) -> torch.Tensor:
        if input_ids is None or input_ids.dim() != 2:
            raise ValueError("input_ids must be a [batch, sequence] tensor of token ids")
        return input_ids

    @staticmethod
    def _proj_residual(x, a, w, b=None):
        """x += a @ w^T (+ b), in the GEMM epilogue."""
        ops.gemm([ops.Problem(A=a, segs=[ops.Segment(W=w, bias=b, out=x, mode=L.EPI_GATE_RESID)])], a.shape[0], a.dtype)


class _Output(tuple):
    """``BaseModelOutput``-like: indexable (``out[0]``) with attributes."""

    def __new__(cls, last_hidden_state, pooler_output=None):
        o = super().__new__(cls, (last_hidden_state,) if pooler_output is None else (last_hidden_state, pooler_output))
        o.last_hidden_state, o.pooler_output = last_hidden_state, pooler_output
        return o


def t5_relative_buckets(S: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """Bucket id of every relative position key - query in [-(S - 1), S - 1] (T5's bidirectional bucketing: half of the
    buckets per sign, exact up to num_buckets / 4, logarithmic up to max_distance)."""
    rel = torch.arange(-(S - 1), S)
    nb = num_buckets // 2
    sign = (rel > 0).long() * nb
    dist = rel.abs()
    exact = nb // 2
    far = exact + (torch.log(dist.float() / exact) / math.log(max_distance / exact) * (nb - exact)).long()
    far = torch.clamp(far, max=nb - 1)
    return sign + torch.where(dist < exact, dist, far)


class T5EncoderModel(_Encoder):
    """``transformers.T5EncoderModel`` (v1.1: gated-GELU, no biases, RMS norms, unscaled scores + relative bias)."""
    _class_name, _also = "T5EncoderModel", ("T5ForConditionalGeneration", "T5Model")
    _defaults = T5_XXL_CONFIG
    _shapes = staticmethod(lambda cfg: t5_param_shapes(cfg))

    def __init__(self, config: Optional[dict], state_dict: Dict[str, torch.Tensor], dtype=torch.bfloat16, device="cuda"):
        super().__init__(config, T5_XXL_CONFIG, state_dict, dtype, device)
        c = self.config
        if c.d_kv != 64:
            raise ValueError("the attention kernel of this path takes head_dim 64 (T5-v1.1-XXL)")
        if c.feed_forward_proj != "gated-gelu":
            raise ValueError("only the gated-GELU feed-forward of T5 v1.1 is implemented")
        if "shared.weight" not in self._w and "encoder.embed_tokens.weight" in self._w:
            self._w["shared.weight"] = self._w["encoder.embed_tokens.weight"]
        self._bias_lut: Dict[int, torch.Tensor] = {}

    def _rel_bias(self, S: int) -> torch.Tensor:
        """[heads, 2 S - 1] fp32: bias of (key - query), from block 0's table (shared by all blocks)."""
        if S not in self._bias_lut:
            c = self.config
            lut = t5_relative_buckets(S, c.relative_attention_num_buckets, c.relative_attention_max_distance)
            table = self._p("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight").float()
            self._bias_lut[S] = table[lut.to(self.device)].t().contiguous()
        return self._bias_lut[S]

    @torch.no_grad()
    def __call__(self, input_ids=None, attention_mask=None, output_hidden_states=False, return_dict=True, **kw):
        if attention_mask is not None:
            raise ValueError("the RepText pipelines call the T5 encoder without a mask (padding attends); masks are "
                             "not implemented")
        ids = self._ids(input_ids)
        c = self.config
        B, S = ids.shape
        D, H, Fd = c.d_model, c.num_heads, c.d_ff
        inner = H * c.d_kv
        x = ops.embedding(self._p("shared.weight"), ids)
        bias = self._rel_bias(S)
        qkv = torch.empty(B, S, 3 * inner, dtype=self.dtype, device=self.device)
        hid = torch.empty(B, S, 2 * Fd, dtype=self.dtype, device=self.device)
        for i in range(c.num_layers):
            p = f"encoder.block.{i}.layer."
            h = ops.norm_rows(x, self._p(p + "0.layer_norm.weight"), None, c.layer_norm_epsilon, False)
            ops.gemm([ops.Problem(A=h, segs=[
                ops.Segment(W=self._p(p + f"0.SelfAttention.{n}.weight"), out=qkv, out_col0=j * inner)
                for j, n in enumerate("qkv")])], B, self.dtype)
            o = ops.text_attention(qkv, H, 1.0, rel_bias=bias)
            self._proj_residual(x, o, self._p(p + "0.SelfAttention.o.weight"))
            h = ops.norm_rows(x, self._p(p + "1.layer_norm.weight"), None, c.layer_norm_epsilon, False)
            ops.gemm([ops.Problem(A=h, segs=[
                ops.Segment(W=self._p(p + "1.DenseReluDense.wi_0.weight"), out=hid, out_col0=0, mode=L.EPI_GELU),
                ops.Segment(W=self._p(p + "1.DenseReluDense.wi_1.weight"), out=hid, out_col0=Fd)])], B, self.dtype)
            g = ops.glu_act(hid, Fd, 0)
            self._proj_residual(x, g, self._p(p + "1.DenseReluDense.wo.weight"))
        out = ops.norm_rows(x, self._p("encoder.final_layer_norm.weight"), None, c.layer_norm_epsilon, False)
        return _Output(out)


class CLIPTextModel(_Encoder):
    """``transformers.CLIPTextModel``: pre-LN blocks, causal mask, quick-GELU; ``pooler_output`` = state at the EOS token."""
    _class_name, _also = "CLIPTextModel", ("CLIPModel", "CLIPTextModelWithProjection")
    _defaults = CLIP_L_CONFIG
    _shapes = staticmethod(lambda cfg: clip_param_shapes(cfg))

    def __init__(self, config: Optional[dict], state_dict: Dict[str, torch.Tensor], dtype=torch.bfloat16, device="cuda"):
        super().__init__(config, CLIP_L_CONFIG, state_dict, dtype, device)
        c = self.config
        if c.hidden_size // c.num_attention_heads != 64:
            raise ValueError("the attention kernel of this path takes head_dim 64 (CLIP ViT-L/14 text model)")
        if c.hidden_act != "quick_gelu":
            raise ValueError("only quick_gelu (CLIP ViT-L/14) is implemented")

    @torch.no_grad()
    def __call__(self, input_ids=None, attention_mask=None, output_hidden_states=False, return_dict=True, **kw):
        if attention_mask is not None:
            raise ValueError("the RepText pipelines call the CLIP text model without a padding mask; not implemented")
        ids = self._ids(input_ids)
        c = self.config
        B, S = ids.shape
        if S > c.max_position_embeddings:
            raise ValueError(f"sequence length {S} exceeds max_position_embeddings {c.max_position_embeddings}")
        D, H, Fd, eps = c.hidden_size, c.num_attention_heads, c.intermediate_size, c.layer_norm_eps
        P = "text_model."
        x = ops.embedding(self._p(P + "embeddings.token_embedding.weight"), ids,
                          self._p(P + "embeddings.position_embedding.weight"))
        qkv = torch.empty(B, S, 3 * D, dtype=self.dtype, device=self.device)
        hid = torch.empty(B, S, Fd, dtype=self.dtype, device=self.device)
        for i in range(c.num_hidden_layers):
            p = f"{P}encoder.layers.{i}."
            h = ops.norm_rows(x, self._p(p + "layer_norm1.weight"), self._p(p + "layer_norm1.bias"), eps, True)
            ops.gemm([ops.Problem(A=h, segs=[
                ops.Segment(W=self._p(p + f"self_attn.{n}_proj.weight"), bias=self._p(p + f"self_attn.{n}_proj.bias"),
                            out=qkv, out_col0=j * D) for j, n in enumerate("qkv")])], B, self.dtype)
            o = ops.text_attention(qkv, H, 64 ** -0.5, causal=True)
            self._proj_residual(x, o, self._p(p + "self_attn.out_proj.weight"), self._p(p + "self_attn.out_proj.bias"))
            h = ops.norm_rows(x, self._p(p + "layer_norm2.weight"), self._p(p + "layer_norm2.bias"), eps, True)
            ops.gemm([ops.Problem(A=h, segs=[ops.Segment(W=self._p(p + "mlp.fc1.weight"), bias=self._p(p + "mlp.fc1.bias"),
                                                         out=hid)])], B, self.dtype)
            g = ops.glu_act(hid, Fd, 1)
            self._proj_residual(x, g, self._p(p + "mlp.fc2.weight"), self._p(p + "mlp.fc2.bias"))
        last = ops.norm_rows(x, self._p(P + "final_layer_norm.weight"), self._p(P + "final_layer_norm.bias"), eps, True)
        # pooled = the state at the EOS token: the largest id with the legacy eos_token_id == 2 configs (FLUX's CLIP-L),
        # else the first eos_token_id.  Host-side index arithmetic on the ids; the row itself is gathered on the device.
        ids_c = ids.to("cpu")
        if c.eos_token_id == 2:
            idx = ids_c.to(torch.int).argmax(dim=-1)
        else:
            idx = (ids_c == c.eos_token_id).int().argmax(dim=-1)
        rows = (torch.arange(B) * S + idx).view(B, 1)
        pooled = ops.embedding(last.view(B * S, D), rows)[:, 0]
        return _Output(last, pooled)


# ------------------------------------------------------------------------------------------------ random weights
def t5_param_shapes(cfg: dict) -> Dict[str, tuple]:
    d, inner, ff = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]
    out = {"shared.weight": (cfg["vocab_size"], d)}
    for i in range(cfg["num_layers"]):
        a, f = f"encoder.block.{i}.layer.0.", f"encoder.block.{i}.layer.1."
        out.update({a + "SelfAttention.q.weight": (inner, d), a + "SelfAttention.k.weight": (inner, d),
                    a + "SelfAttention.v.weight": (inner, d), a + "SelfAttention.o.weight": (d, inner),
                    a + "layer_norm.weight": (d,), f + "DenseReluDense.wi_0.weight": (ff, d),
                    f + "DenseReluDense.wi_1.weight": (ff, d), f + "DenseReluDense.wo.weight": (d, ff),
                    f + "layer_norm.weight": (d,)})
    out["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"] = (
        cfg["relative_attention_num_buckets"], cfg["num_heads"])
    out["encoder.final_layer_norm.weight"] = (d,)
    return out


def clip_param_shapes(cfg: dict) -> Dict[str, tuple]:
    d, ff = cfg["hidden_size"], cfg["intermediate_size"]
    P = "text_model."
    out = {P + "embeddings.token_embedding.weight": (cfg["vocab_size"], d),
           P + "embeddings.position_embedding.weight": (cfg["max_position_embeddings"], d),
           P + "final_layer_norm.weight": (d,), P + "final_layer_norm.bias": (d,)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"{P}encoder.layers.{i}."
        for n, (o, k) in {"self_attn.q_proj": (d, d), "self_attn.k_proj": (d, d), "self_attn.v_proj": (d, d),
                          "self_attn.out_proj": (d, d), "mlp.fc1": (ff, d), "mlp.fc2": (d, ff)}.items():
            out[p + n + ".weight"], out[p + n + ".bias"] = (o, k), (o,)
        for n in ("layer_norm1", "layer_norm2"):
            out[p + n + ".weight"], out[p + n + ".bias"] = (d,), (d,)
    return out


def random_weights(shapes: Dict[str, tuple], seed: int, device, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Random weights of the named architecture, drawn ON THE DEVICE (T5-XXL is 4.7 B parameters): linears fan-in scaled,
    norms near 1, small biases, unit embeddings."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, s in shapes.items():
        t = torch.randn(s, generator=g, device=device, dtype=torch.float32)
        if "relative_attention_bias" in k or k == "shared.weight":
            pass
        elif "embedding" in k:
            t *= 0.5
        elif k.endswith(".bias"):
            t *= 0.02
        elif len(s) == 1:
            t = 1.0 + 0.1 * t
        else:
            t *= s[1] ** -0.5
            if k.endswith("SelfAttention.q.weight"):
                t *= 64 ** -0.5      # T5 keeps its unscaled scores O(1) through q's initialisation (transformers' init)
        sd[k] = t.to(dtype)
    return sd
