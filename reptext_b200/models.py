"""Host-side mirrors of the two modules the reference pipelines call every step.

* :class:`FluxControlNetModel` keeps the call signature of ``RepText/controlnet_flux.py:216-229`` and the
  return convention of ``:398-413``.
* :class:`FluxTransformer2DModel` keeps the signature diffusers 0.36.0 exposes and the reference uses at
  ``RepText/pipeline_flux_controlnet.py:1092-1104``.

Both are thin: they validate arguments (raising what the reference raises), hand raw device pointers to the
C-ABI (``include/reptext_rt.h``) and wrap the result buffers as tensors.  All arithmetic is in
``csrc/*.cu``; there is no PyTorch fallback - without the built library, or on a CPU tensor, they raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from . import _lib as L
from . import weights as W

_WORKSPACES: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """One scratch buffer per (device, stream): the pipelines run the ControlNet and the transformer
    back to back on one stream, so they share it.  It only ever grows."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.0) + 256, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


@dataclass
class FluxControlNetOutput:
    """``RepText/controlnet_flux.py:35-38``."""
    controlnet_block_samples: Optional[List[torch.Tensor]]
    controlnet_single_block_samples: Optional[List[torch.Tensor]]


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


class FrozenConfig(SimpleNamespace):
    """Attribute- and item-style access, like diffusers' ``FrozenDict`` config."""

    def __getitem__(self, k):
        return getattr(self, k)

    def get(self, k, default=None):
        return getattr(self, k, default)


class _RuntimeModel:
    """Common part: owns the parameter tensors (diffusers state-dict names) and the native handle."""

    _kind = "transformer"

    def __init__(self, config: dict, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 dtype: torch.dtype = torch.bfloat16, device: Union[str, torch.device] = "cuda"):
        self.config = FrozenConfig(**config)
        self._cfg = dict(config)
        self._dtype = dtype
        self._device = torch.device(device)
        self._handle: Optional[C.c_void_p] = None
        self._params: Dict[str, torch.Tensor] = {}
        self._inv_on = False          # step-invariant cache (set_step_invariant_cache)
        self._inv_key = None
        self._inv_conv = None
        self._inv_src = None
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ---- nn.Module-like surface the pipelines touch -------------------------------------------------
    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    @property
    def device(self) -> torch.device:
        return self._device

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a != self._dtype:
                raise ValueError("reptext_b200 models are built for one dtype; pass dtype= at construction")
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise ValueError("reptext_b200 models live on a CUDA device (there is no CPU path)")
        return self

    def eval(self):
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._params)

    def parameters(self):
        return iter(self._params.values())

    @classmethod
    def from_config(cls, config: dict, **kw):
        return cls(config, **kw)

    # ---- checkpoints in the reference's directory layout (RepText/infer.py:27-33) --------------------
    _class_name = "FluxTransformer2DModel"
    _config_defaults: Dict[str, Any] = {}

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None,
                        torch_dtype: torch.dtype = torch.bfloat16, device: Union[str, torch.device] = "cuda",
                        variant: Optional[str] = None, **unused):
        """``<dir>/config.json`` + ``<dir>/diffusion_pytorch_model.safetensors`` (single file, or the shards of
        ``...safetensors.index.json``), as ``ModelMixin.from_pretrained`` reads them - from a LOCAL directory only
        (:mod:`reptext_b200.checkpoint`).  Keys missing from ``config.json`` take the constructor defaults of the class
        the directory was saved from (``RepText/controlnet_flux.py:44-60`` / diffusers' ``FluxTransformer2DModel``)."""
        from . import checkpoint as ck
        d = ck.resolve_dir(pretrained_model_name_or_path, subfolder)
        raw, stored_cls = ck.read_config(d)
        if stored_cls not in (None, cls._class_name):
            raise ValueError(f"{d!r} holds a {stored_cls}, not a {cls._class_name}")
        cfg = dict(cls._config_defaults)
        cfg.update(raw)
        cfg["axes_dims_rope"] = tuple(cfg["axes_dims_rope"])
        return cls(cfg, ck.load_state_dict(d, (ck.DIFFUSERS_STEM,), variant), dtype=torch_dtype, device=device)

    def save_pretrained(self, save_directory, max_shard_size: int = 10 * 2 ** 30) -> None:
        """Write ``config.json`` and the weights (sharded above ``max_shard_size`` bytes) so that
        :meth:`from_pretrained` - and diffusers' own loader, the key names are its - reads them back."""
        from . import checkpoint as ck
        ck.write_config(save_directory, self._cfg, self._class_name)
        ck.save_state_dict(save_directory, self.state_dict(), ck.DIFFUSERS_STEM, max_shard_size)

    @classmethod
    def random_init(cls, config: dict, seed: int = 0, dtype: torch.dtype = torch.bfloat16, device="cuda",
                    zero_init: bool = False):
        """Seeded random-init weights with the FLUX.1-dev / RepText architecture (no network for checkpoints)."""
        sd = W.random_state_dict(config, cls._kind, seed=seed, dtype=dtype, device=device, zero_init=zero_init)
        return cls(config, sd, dtype=dtype, device=device)

    # ---- native handle ------------------------------------------------------------------------------
    def _native_config(self) -> L.ModelConfig:
        c = self._cfg
        nc = L.ModelConfig()
        nc.kind = L.RT_CONTROLNET if self._kind == "controlnet" else L.RT_TRANSFORMER
        nc.dtype = L.dtype_code(self._dtype)
        nc.in_channels = c["in_channels"]
        nc.cond_channels = c["in_channels"] + c.get("extra_condition_channels", 0)
        nc.out_channels = (c.get("out_channels") or c["in_channels"]) * c.get("patch_size", 1) ** 2
        nc.num_layers = c["num_layers"]
        nc.num_single_layers = c["num_single_layers"]
        nc.num_attention_heads = c["num_attention_heads"]
        nc.attention_head_dim = c["attention_head_dim"]
        nc.joint_attention_dim = c["joint_attention_dim"]
        nc.pooled_projection_dim = c["pooled_projection_dim"]
        nc.guidance_embeds = int(bool(c.get("guidance_embeds", False)))
        for i, a in enumerate(c["axes_dims_rope"]):
            nc.axes_dims_rope[i] = int(a)
        return nc

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        lib = L.lib()
        want = {k: tuple(s) for k, s, _ in W.param_table(self._cfg, self._kind)}
        missing = [k for k in want if k not in sd]
        unexpected = [k for k in sd if k not in want]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing keys {missing[:4]}..., unexpected keys {unexpected[:4]}...")
        if self._handle is not None:
            L.check(lib.rt_model_destroy(self._handle))
            self._handle = None
        h = C.c_void_p()
        nc = self._native_config()
        L.check(lib.rt_model_create(C.byref(nc), C.byref(h)))
        self._params = {}
        for k, shape in want.items():
            t = sd[k]
            if tuple(t.shape) != shape:
                raise RuntimeError(f"load_state_dict: {k} has shape {tuple(t.shape)}, expected {shape}")
            t = t.detach().to(device=self._device, dtype=self._dtype).contiguous()
            self._params[k] = t
            shp = (C.c_int64 * t.dim())(*t.shape)
            L.check(lib.rt_model_set_weight(h, k.encode(), L.ptr(t), shp, t.dim()))
        L.check(lib.rt_model_finalize(h, L.stream_ptr()))
        self._handle = h
        self.release_step_invariants()

    # ---- step-invariant inputs -----------------------------------------------------
    def set_step_invariant_cache(self, on: bool) -> None:
        """``on``: keep what a forward derives from ``encoder_hidden_states`` / ``pooled_projections`` / ``guidance`` /
        ``txt_ids`` / ``img_ids`` alone - ``context_embedder(enc)`` (RepText/controlnet_flux.py:292), the rotary table
        (:316-317), the first guidance / pooled-text linears (:282-291), and the dtype conversions of those tensors -
        across calls (``rt_model_set_step_invariant_cache``).  A call whose tensors are the SAME objects' storage in the
        same version (``data_ptr`` / ``_version`` / shape / strides / dtype) reuses them; anything else recomputes.  The
        tensors are referenced while cached (their memory cannot be recycled under the key).  Results are
        bit-identical.  The RepText pipelines switch it on for their denoising loop
        (``pipe.cache_step_invariants``) and release it afterwards; a bare model computes everything every call."""
        self._inv_on = bool(on)
        self.release_step_invariants()

    def release_step_invariants(self) -> None:
        """Forget the cached step invariants (and the references that pin their inputs); the mode stays."""
        self._inv_key = self._inv_conv = self._inv_src = None
        if self._handle is not None:
            L.check(L.lib().rt_model_set_step_invariant_cache(self._handle, int(self._inv_on)))

    # ---- the AdaLN vectors of every step of an image in one pass ---------------------------------------
    def build_modulation_table(self, timesteps: torch.Tensor, guidance: Optional[torch.Tensor],
                               pooled_projections: torch.Tensor) -> int:
        """``timesteps``: [steps, batch or 1] - what each step of a denoising loop will pass as ``timestep`` (already
        divided by 1000, RepText/pipeline_flux_controlnet.py:1043); ``guidance``: [batch] or [1] (or None);
        ``pooled_projections``: [batch, P].  Runs the time / guidance / pooled-text embedders and the AdaLN linears of
        every block (RepText/controlnet_flux.py:282-291 and the blocks' ``norm1`` / ``norm``) for all steps at once
        (``rt_model_build_modulation_table``): the 6.5 GB of modulation weights are read once per image instead of once per
        step.  ``select_modulation(i)`` then makes the forwards use step i's vectors; results are bit-identical to
        forwards that compute them.  Returns the number of steps."""
        if self._handle is None:
            raise RuntimeError("model has no weights: call load_state_dict() or use random_init()")
        c = self._cfg
        if timesteps.dim() != 2:
            raise ValueError("`timesteps` must be [steps, batch]")
        K, Bt = timesteps.shape
        pooled = self._check("pooled_projections", pooled_projections, c["pooled_projection_dim"])
        B = pooled.shape[0]
        if Bt not in (1, B):   # (batch-1 latents against batch-2 embeddings: the inpaint pipeline's true-CFG call)
            raise ValueError("timesteps must have 1 or batch elements per step")
        ts = self._check("timesteps", timesteps.expand(K, B)).reshape(-1)
        pooled = pooled.repeat(K, 1).contiguous()
        g = None
        if c.get("guidance_embeds", False):
            if guidance is None:
                raise ValueError("this model was built with guidance_embeds=True: `guidance` is required")
            g = self._check("guidance", guidance.reshape(-1))
            if g.numel() not in (1, B):
                raise ValueError("guidance must have 1 or batch elements")
            g = g.expand(B).repeat(K).contiguous()
        lib = L.lib()
        need = int(lib.rt_model_modulation_table_bytes(self._handle, K, B)) + 256
        buf = getattr(self, "_modtab_buf", None)
        if buf is None or buf.numel() < need:
            self.select_modulation(None)
            self._modtab_buf = buf = torch.empty(need, dtype=torch.uint8, device=self._device)   # kept across images
        base = (buf.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(self._device):
            L.check(lib.rt_model_build_modulation_table(self._handle, L.ptr(ts), L.ptr(g) if g is not None else None,
                                                        L.ptr(pooled), K, B, base, buf.numel() - (base - buf.data_ptr()),
                                                        L.stream_ptr()))
        return K

    def select_modulation(self, step: Optional[int]) -> None:
        """Use row ``step`` of the table for the following forwards (their ``timestep`` / ``guidance`` /
        ``pooled_projections`` are then not read); ``None``: compute per forward again."""
        if self._handle is not None:
            L.check(L.lib().rt_model_select_modulation(self._handle, -1 if step is None else int(step)))

    @staticmethod
    def _tensor_key(t):
        return None if t is None else (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.dtype, t.device)

    def __del__(self):
        try:
            if self._handle is not None and L._lib is not None:
                L._lib.rt_model_destroy(self._handle)
        except Exception:
            pass

    # ---- argument marshalling -----------------------------------------------------------------------
    def _check(self, name: str, t: torch.Tensor, last: Optional[int] = None, dtype: Optional[torch.dtype] = None):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"`{name}` must be a torch.Tensor")
        if not t.is_cuda:
            raise ValueError(f"`{name}` must be a CUDA tensor (reptext_b200 has no CPU path)")
        if last is not None and t.shape[-1] != last:
            raise ValueError(f"`{name}` has {t.shape[-1]} features, the model expects {last}")
        return t.to(dtype or self._dtype).contiguous()

    def _forward_args(self, hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids,
                      guidance, keep: list, sp=None, sp_rank: Optional[int] = None) -> L.ForwardArgs:
        """``sp``: a :class:`reptext_b200.parallel.SequenceParallelGroup` (or ``LockstepGroup``); every tensor then
        holds this rank's token shard and the group's peer-mapped buffer is the workspace."""
        c = self._cfg
        if self._handle is None:
            raise RuntimeError("model has no weights: call load_state_dict() or use random_init()")
        hs = self._check("hidden_states", hidden_states, c["in_channels"])
        use_inv = self._inv_on and sp_rank is None     # (lock-step ranks share this model object)
        if use_inv:
            src = (encoder_hidden_states, pooled_projections, guidance, txt_ids, img_ids, timestep.numel())
            key = tuple(self._tensor_key(t) if isinstance(t, torch.Tensor) else t for t in src)
            if key == self._inv_key:
                # same storage, same version: the converted tensors (and the native cache behind them) are current
                enc, pooled, g, ii, ti = self._inv_conv
                return self._fill_args(hs, enc, pooled, self._check("timestep", timestep.reshape(-1)), g, ii, ti, keep,
                                       sp, sp_rank)
        enc = self._check("encoder_hidden_states", encoder_hidden_states, c["joint_attention_dim"])
        pooled = self._check("pooled_projections", pooled_projections, c["pooled_projection_dim"])
        if hs.dim() != 3 or enc.dim() != 3 or pooled.dim() != 2:
            raise ValueError("hidden_states / encoder_hidden_states must be 3-D and pooled_projections 2-D")
        if txt_ids.ndim == 3:   # deprecated batched ids (controlnet_flux.py:302-315)
            txt_ids = txt_ids[0]
        if img_ids.ndim == 3:
            img_ids = img_ids[0]
        B = enc.shape[0]
        if pooled.shape[0] != B:
            raise ValueError("pooled_projections and encoder_hidden_states disagree on the batch size")
        if hs.shape[0] not in (1, B):
            raise ValueError("hidden_states batch must be 1 or match encoder_hidden_states")
        N, T = hs.shape[1], enc.shape[1]
        if img_ids.shape[0] != N or txt_ids.shape[0] != T:
            raise ValueError("img_ids / txt_ids do not match the token counts")
        ts = self._check("timestep", timestep.reshape(-1))
        if ts.numel() not in (1, B):
            raise ValueError("timestep must have 1 or batch elements")
        g = None
        if c.get("guidance_embeds", False):
            if guidance is None:
                raise ValueError("this model was built with guidance_embeds=True: `guidance` is required")
            g = self._check("guidance", guidance.reshape(-1))
            if g.numel() != ts.numel():
                g = g.expand(ts.numel()).contiguous() if g.numel() == 1 else g
            if g.numel() != ts.numel():
                raise ValueError("guidance and timestep disagree on the batch size")
        ii = self._check("img_ids", img_ids, 3, torch.float32)
        ti = self._check("txt_ids", txt_ids, 3, torch.float32)
        if use_inv:
            # new inputs: invalidate the native cache, remember the key and pin the tensors it stands for
            L.check(L.lib().rt_model_set_step_invariant_cache(self._handle, 1))
            self._inv_key, self._inv_conv, self._inv_src = key, (enc, pooled, g, ii, ti), src
        return self._fill_args(hs, enc, pooled, ts, g, ii, ti, keep, sp, sp_rank)

    def _fill_args(self, hs, enc, pooled, ts, g, ii, ti, keep: list, sp, sp_rank) -> L.ForwardArgs:
        c = self._cfg
        B, N, T = enc.shape[0], hs.shape[1], enc.shape[1]
        if hs.shape[0] not in (1, B) or ii.shape[0] != N or ts.numel() not in (1, B):
            raise ValueError("hidden_states / img_ids / timestep do not match the cached embeddings' shapes")
        nbytes = L.lib().rt_model_workspace_bytes(self._handle, B, N, T)
        a = L.ForwardArgs()
        if sp is None:
            ws = _workspace(hs.device, nbytes)
            base = (ws.data_ptr() + 255) // 256 * 256
            a.workspace, a.workspace_bytes = base, ws.numel() - (base - ws.data_ptr())
        else:
            if self._dtype != torch.bfloat16 or c["attention_head_dim"] != 128:
                raise ValueError("sequence-parallel mode needs a bf16 model with attention_head_dim 128")
            if c["num_attention_heads"] % sp.world:
                raise ValueError(f"{c['num_attention_heads']} heads cannot be sharded over {sp.world} ranks")
            ws = None
            sp.ensure_workspace(nbytes)          # collective on a multi-process group; a no-op once large enough
            grp = sp.struct(sp_rank)
            a.workspace, a.workspace_bytes = grp.peer_workspace[grp.rank], sp.workspace_bytes
            a.sp = C.pointer(grp)
            keep.append(grp)
        a.batch, a.lat_batch, a.t_batch, a.n_img, a.n_txt = B, hs.shape[0], ts.numel(), N, T
        a.hidden_states, a.encoder_hidden_states, a.pooled_projections = L.ptr(hs), L.ptr(enc), L.ptr(pooled)
        a.timestep, a.guidance, a.img_ids, a.txt_ids = L.ptr(ts), L.ptr(g), L.ptr(ii), L.ptr(ti)
        a.stream = L.stream_ptr()
        keep.extend([hs, enc, pooled, ts, g, ii, ti, ws])
        return a


class FluxControlNetModel(_RuntimeModel):
    """Drop-in for ``RepText/controlnet_flux.py:41`` (inference path)."""

    _kind = "controlnet"
    _class_name = "FluxControlNetModel"
    # constructor defaults of RepText/controlnet_flux.py:44-60 (a config.json only stores what was passed)
    _config_defaults = dict(patch_size=1, in_channels=64, num_layers=19, num_single_layers=38, attention_head_dim=128,
                            num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768,
                            guidance_embeds=False, axes_dims_rope=(16, 56, 56), num_mode=None,
                            extra_conditioning_channels=0, extra_condition_channels=0)

    def __init__(self, config: dict, state_dict=None, dtype=torch.bfloat16, device="cuda"):
        if config.get("num_mode") is not None:
            raise NotImplementedError("ControlNet-Union (num_mode) is outside the RepText hot path")
        self._live = (-1, -1)
        super().__init__(config, state_dict, dtype, device)
        self.union = False

    def load_state_dict(self, sd, strict: bool = True) -> None:
        super().load_state_dict(sd, strict)
        self._apply_live()

    # ---- samples nobody consumes are not computed -----------------------------------------------------
    @staticmethod
    def consumed_samples(n_samples: int, consumer_layers: int) -> int:
        """How many of ``n_samples`` ControlNet samples ``FluxTransformer2DModel.forward`` reads: sample
        ``i // ceil(L / n)`` after block ``i`` (diffusers 0.36): 6 samples, 19 blocks -> 5."""
        if n_samples == 0 or consumer_layers <= 0:
            return 0
        interval = -(-consumer_layers // n_samples)
        return min(n_samples, (consumer_layers - 1) // interval + 1)

    def set_consumer(self, num_layers: Optional[int], num_single_layers: Optional[int] = None) -> None:
        """Declare the transformer whose ``controlnet_block_samples`` / ``controlnet_single_block_samples`` this model
        feeds (``None, None`` = unknown consumer: compute every sample, the default).  Blocks that only produce samples
        the consumer never reads are then skipped (``rt_controlnet_set_live``); those samples come back as ZEROS.  The
        RepText pipelines call this with their transformer's layer counts (``skip_unconsumed_controlnet_blocks``);
        with FLUX.1-dev + RepText that removes the sixth ControlNet block and zero-linear - 1.7 % of a step - and the
        latents are bit-identical because sample 5 is never added anywhere."""
        c = self._cfg
        if num_layers is None:
            self._live = (-1, -1)
        else:
            live_d = self.consumed_samples(c["num_layers"], num_layers)
            live_s = self.consumed_samples(c["num_single_layers"], num_single_layers or 0)
            if live_s > 0:               # the single blocks run after ALL double blocks
                live_d = c["num_layers"]
            self._live = (live_d, live_s)
        self._apply_live()

    def _apply_live(self) -> None:
        if self._handle is not None:
            L.check(L.lib().rt_controlnet_set_live(self._handle, self._live[0], self._live[1]))

    @property
    def inner_dim(self) -> int:
        return self._cfg["num_attention_heads"] * self._cfg["attention_head_dim"]

    @torch.no_grad()
    def forward(
        self,
        hidden_states: torch.Tensor,
        controlnet_cond: torch.Tensor,
        controlnet_mode: torch.Tensor = None,
        conditioning_scale: float = 1.0,
        encoder_hidden_states: torch.Tensor = None,
        pooled_projections: torch.Tensor = None,
        timestep: torch.Tensor = None,
        img_ids: torch.Tensor = None,
        txt_ids: torch.Tensor = None,
        guidance: torch.Tensor = None,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        return_dict: bool = True,
        *,
        regional_mask: Optional[torch.Tensor] = None,
        accumulate_into: Optional[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]] = None,
        sp=None,
    ):
        """Same arguments as the reference.  Two keyword-only extensions let the pipelines fuse their
        per-step host work into the zero-linear epilogue: ``regional_mask`` ([1, N, 1]; the multiply at
        ``pipeline_flux_controlnet.py:1060-1069``) and ``accumulate_into`` (the stacked outputs of a previous
        text line; the sum at ``:1072-1087``).  ``sp``: sequence-parallel group (every tensor is then this rank's
        token shard; BASELINE.json configs[4])."""
        with torch.cuda.device(self._device):  # the library launches on the current device's stream
            call, keep, (blocks, singles) = self._marshal(
                hidden_states, controlnet_cond, controlnet_mode, conditioning_scale, encoder_hidden_states,
                pooled_projections, timestep, img_ids, txt_ids, guidance, joint_attention_kwargs, regional_mask,
                accumulate_into, sp, None)
            L.check(L.lib().rt_controlnet_forward(self._handle, C.byref(call.a), call.controlnet_cond,
                                                  call.cond_batch, call.conditioning_scale, call.mask,
                                                  call.accumulate, call.block_samples, call.single_block_samples))
        return self._wrap(blocks, singles, return_dict)

    def forward_lockstep(self, group, per_rank: List[dict], return_dict: bool = False):
        """Every rank's forward driven in phase order by this process on one GPU (``LockstepGroup``): the same
        peer-store indexing as the multi-process mode without needing several GPUs.  ``per_rank[i]`` holds rank i's
        keyword arguments of :meth:`forward`."""
        calls = (L.ControlNetCall * group.world)()
        keep, outs = [], []
        for r, kw in enumerate(per_rank):
            kw = dict(kw)
            call, k, bufs = self._marshal(
                kw.pop("hidden_states"), kw.pop("controlnet_cond"), kw.pop("controlnet_mode", None),
                kw.pop("conditioning_scale", 1.0), kw.pop("encoder_hidden_states"), kw.pop("pooled_projections"),
                kw.pop("timestep"), kw.pop("img_ids"), kw.pop("txt_ids"), kw.pop("guidance", None),
                kw.pop("joint_attention_kwargs", None), kw.pop("regional_mask", None),
                kw.pop("accumulate_into", None), group, r)
            if kw:
                raise TypeError(f"unexpected arguments {sorted(kw)}")
            calls[r] = call
            keep.append(k)
            outs.append(bufs)
        L.check(L.lib().rt_controlnet_forward_lockstep(self._handle, group.world, calls))
        return [self._wrap(b, s_, return_dict) for b, s_ in outs]

    def _marshal(self, hidden_states, controlnet_cond, controlnet_mode, conditioning_scale, encoder_hidden_states,
                 pooled_projections, timestep, img_ids, txt_ids, guidance, joint_attention_kwargs, regional_mask,
                 accumulate_into, sp, sp_rank):
        if controlnet_mode is not None:
            raise ValueError("`controlnet_mode` is only valid for ControlNet-Union models (num_mode is None here)")
        if joint_attention_kwargs:
            extra = set(joint_attention_kwargs) - {"scale"}
            if extra:
                raise NotImplementedError(f"joint_attention_kwargs {sorted(extra)} are not supported")
        c = self._cfg
        keep: list = []
        a = self._forward_args(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids,
                               guidance if c.get("guidance_embeds") else None, keep, sp, sp_rank)
        cond = self._check("controlnet_cond", controlnet_cond, c["in_channels"] + c.get("extra_condition_channels", 0))
        if cond.dim() != 3 or cond.shape[1] != a.n_img or cond.shape[0] not in (1, a.batch):
            raise ValueError("controlnet_cond must be [1 or batch, n_img, in_channels + extra_condition_channels]")
        B, N, D = a.batch, a.n_img, self.inner_dim
        nl, ns = c["num_layers"], c["num_single_layers"]
        acc_b, acc_s = accumulate_into if accumulate_into is not None else (None, None)

        def out_buf(n, acc, live):
            if n == 0:
                return None
            if acc is not None:
                if tuple(acc.shape) != (n, B, N, D) or acc.dtype != self._dtype or not acc.is_contiguous():
                    raise ValueError("accumulate_into must be the stacked [layers, B, N, D] output of a previous call")
                return acc
            out = torch.empty(n, B, N, D, dtype=self._dtype, device=cond.device)
            if 0 <= live < n:            # samples no block writes (set_consumer): defined, and zero
                out[live:].zero_()
            return out

        live_d, live_s = self._live
        blocks = out_buf(nl, acc_b, live_d if live_s <= 0 else -1)
        singles = out_buf(ns, acc_s, live_s)
        mask = None
        if regional_mask is not None:
            mask = self._check("regional_mask", regional_mask.reshape(-1))
            if mask.numel() != N:
                raise ValueError("regional_mask must have one value per image token")
        call = L.ControlNetCall()
        call.a = a
        call.controlnet_cond, call.cond_batch = L.ptr(cond), cond.shape[0]
        call.conditioning_scale, call.mask = float(conditioning_scale), L.ptr(mask)
        call.accumulate = int(accumulate_into is not None)
        call.block_samples, call.single_block_samples = L.ptr(blocks), L.ptr(singles)
        keep.extend([cond, mask])
        return call, keep, (blocks, singles)

    @staticmethod
    def _wrap(blocks, singles, return_dict):
        bl = list(blocks.unbind(0)) if blocks is not None else None
        sl = list(singles.unbind(0)) if singles is not None else None
        if bl is not None:
            bl[0]._rt_stacked = blocks          # lets the pipeline pass the stack back as accumulate_into
        if sl is not None:
            sl[0]._rt_stacked = singles
        if not return_dict:
            return (bl, sl)
        return FluxControlNetOutput(controlnet_block_samples=bl, controlnet_single_block_samples=sl)

    __call__ = forward


class FluxTransformer2DModel(_RuntimeModel):
    """Drop-in for diffusers' ``FluxTransformer2DModel`` as the reference pipelines call it."""

    _kind = "transformer"
    _class_name = "FluxTransformer2DModel"
    # constructor defaults of diffusers' FluxTransformer2DModel (FLUX.1-dev's config.json has no axes_dims_rope)
    _config_defaults = dict(patch_size=1, in_channels=64, out_channels=None, num_layers=19, num_single_layers=38,
                            attention_head_dim=128, num_attention_heads=24, joint_attention_dim=4096,
                            pooled_projection_dim=768, guidance_embeds=False, axes_dims_rope=(16, 56, 56))

    @torch.no_grad()
    def forward(
        self,
        hidden_states: torch.Tensor,
        encoder_hidden_states: torch.Tensor = None,
        pooled_projections: torch.Tensor = None,
        timestep: torch.Tensor = None,
        img_ids: torch.Tensor = None,
        txt_ids: torch.Tensor = None,
        guidance: torch.Tensor = None,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        controlnet_block_samples=None,
        controlnet_single_block_samples=None,
        return_dict: bool = True,
        controlnet_blocks_repeat: bool = False,
        *,
        sp=None,
    ):
        with torch.cuda.device(self._device):
            call, keep, out = self._marshal(hidden_states, encoder_hidden_states, pooled_projections, timestep,
                                            img_ids, txt_ids, guidance, joint_attention_kwargs,
                                            controlnet_block_samples, controlnet_single_block_samples,
                                            controlnet_blocks_repeat, sp, None)
            L.check(L.lib().rt_transformer_forward(self._handle, C.byref(call.a), call.controlnet_block_samples,
                                                   call.n_block_samples, call.controlnet_single_block_samples,
                                                   call.n_single_block_samples, call.out))
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)

    def forward_lockstep(self, group, per_rank: List[dict]) -> List[torch.Tensor]:
        """See :meth:`FluxControlNetModel.forward_lockstep`.  Returns every rank's ``[B, N_local, C]`` output."""
        calls = (L.TransformerCall * group.world)()
        keep, outs = [], []
        for r, kw in enumerate(per_rank):
            kw = dict(kw)
            call, k, out = self._marshal(
                kw.pop("hidden_states"), kw.pop("encoder_hidden_states"), kw.pop("pooled_projections"),
                kw.pop("timestep"), kw.pop("img_ids"), kw.pop("txt_ids"), kw.pop("guidance", None),
                kw.pop("joint_attention_kwargs", None), kw.pop("controlnet_block_samples", None),
                kw.pop("controlnet_single_block_samples", None), False, group, r)
            if kw:
                raise TypeError(f"unexpected arguments {sorted(kw)}")
            calls[r] = call
            keep.append(k)
            outs.append(out)
        L.check(L.lib().rt_transformer_forward_lockstep(self._handle, group.world, calls))
        return outs

    def _marshal(self, hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
                 joint_attention_kwargs, controlnet_block_samples, controlnet_single_block_samples,
                 controlnet_blocks_repeat, sp, sp_rank):
        if controlnet_blocks_repeat:
            raise NotImplementedError("controlnet_blocks_repeat is not used by the RepText pipelines")
        if joint_attention_kwargs:
            extra = set(joint_attention_kwargs) - {"scale"}
            if extra:
                raise NotImplementedError(f"joint_attention_kwargs {sorted(extra)} are not supported")
        c = self._cfg
        keep: list = []
        a = self._forward_args(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids,
                               guidance if c.get("guidance_embeds") else None, keep, sp, sp_rank)
        B, N, D = a.batch, a.n_img, c["num_attention_heads"] * c["attention_head_dim"]

        def ptr_array(samples, what):
            if samples is None or len(samples) == 0:
                return None, 0
            arr = (C.c_void_p * len(samples))()
            for i, s in enumerate(samples):
                s = self._check(what, s, D)
                if tuple(s.shape) != (B, N, D):
                    raise ValueError(f"{what}[{i}] must be [batch, n_img, {D}]")
                keep.append(s)
                arr[i] = L.ptr(s)
            return arr, len(samples)

        bp, nb = ptr_array(controlnet_block_samples, "controlnet_block_samples")
        spp, nsg = ptr_array(controlnet_single_block_samples, "controlnet_single_block_samples")
        co = (c.get("out_channels") or c["in_channels"]) * c.get("patch_size", 1) ** 2
        out = torch.empty(B, N, co, dtype=self._dtype, device=self._device)
        call = L.TransformerCall()
        call.a = a
        call.controlnet_block_samples, call.n_block_samples = bp, nb
        call.controlnet_single_block_samples, call.n_single_block_samples = spp, nsg
        call.out = L.ptr(out)
        keep.extend([bp, spp])
        return call, keep, out

    __call__ = forward
