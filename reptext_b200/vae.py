"""``AutoencoderKL`` as the RepText pipelines call it, on the C-ABI runtime.

Reference call sites: ``RepText/pipeline_flux_controlnet.py:705-708`` (``vae.encode(image).latent_dist.sample()``, then
``(z - shift_factor) * scaling_factor``), ``:1136-1140`` (``vae.decode(latents / scaling_factor + shift_factor,
return_dict=False)[0]``), ``:220-222`` (``vae.config.block_out_channels``), ``pipeline_flux_controlnet_inpaint.py:761-826``
(masked-image encode).  The arithmetic is diffusers' ``AutoencoderKL`` with the FLUX.1-dev VAE config (restated in
``oracle/vae_oracle.py``, which the parity tests use as the checker).

Layout: activations are NHWC bf16, shaped ``[B, H * W, C]``, so that
* a 3x3 / stride-1 convolution is ONE implicit-GEMM launch of the tcgen05 GEMM (``rt_gemm`` with ``conv_h / conv_w /
  conv_c``: the 9 taps are 9 shifted 4-D TMA boxes of the image, TMA's out-of-bounds zero fill is the padding) with the
  bias and - for ``conv2`` of a ResnetBlock2D - the skip connection fused into the epilogue,
* 1x1 shortcuts and the attention projections are plain GEMMs over pixels,
* the three stride-2 downsampling convolutions and the 3-channel ``conv_in`` (shapes TMA boxes cannot address) go
  through an im2col gather + plain GEMM,
* the single-head mid-block attention (head_dim = 512, which the flash kernel of the denoiser does not take) is three
  GEMMs around a row softmax: ``S = (q k^T) / sqrt(C)``, ``P = softmax(S)``, ``o = P v``; ``v`` is produced already
  transposed (``v^T = W_v t^T``) and its bias is folded into ``to_out``'s (softmax rows sum to one),
* GroupNorm (+ SiLU) is a statistics pass + an apply pass (HBM-bound).

There is no CPU path: every tensor must live on a CUDA device and the shared library must be present.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib as L
from . import ops
from .models import FrozenConfig

FLUX_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, scaling_factor=0.3611, shift_factor=0.1159,
                       use_quant_conv=False, use_post_quant_conv=False, force_upcast=True)


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


class DiagonalGaussianDistribution:
    """diffusers' posterior object over NHWC moments: ``sample(generator)`` / ``mode()`` return NCHW latents."""

    def __init__(self, moments_nhwc: torch.Tensor, hw: Tuple[int, int], latent_channels: int, dtype: torch.dtype):
        self._m, self._hw, self._lc, self._dtype = moments_nhwc, hw, latent_channels, dtype

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        from .pipeline_utils import randn_tensor
        shape = (self._m.shape[0], self._lc, self._hw[0], self._hw[1])
        noise = randn_tensor(shape, generator=generator, device=self._m.device, dtype=self._dtype)
        return ops.vae_posterior_sample(self._m, self._hw, self._lc, noise, self._dtype)

    def sample_with_noise(self, noise: torch.Tensor) -> torch.Tensor:
        return ops.vae_posterior_sample(self._m, self._hw, self._lc, noise.to(self._m.device), self._dtype)

    def mode(self) -> torch.Tensor:
        return ops.vae_posterior_sample(self._m, self._hw, self._lc, None, self._dtype)

    @property
    def mean(self) -> torch.Tensor:
        return self.mode()

    @property
    def parameters(self) -> torch.Tensor:
        """The moments as diffusers holds them: NCHW [B, 2 * latent, h, w]."""
        return ops.nhwc_to_nchw(self._m, self._hw, 2 * self._lc, self._dtype)


def vae_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """Every parameter of a FLUX-style ``AutoencoderKL`` (diffusers names) with its shape: what a checkpoint of
    ``black-forest-labs/FLUX.1-dev/vae`` holds, used here to draw random weights of that architecture."""
    boc, lpb, lat = tuple(cfg["block_out_channels"]), cfg["layers_per_block"], cfg["latent_channels"]
    out: Dict[str, Tuple[int, ...]] = {}

    def wb(name, *shape):
        out[name + ".weight"] = tuple(shape)
        out[name + ".bias"] = (shape[0],)

    def resnet(p, cin, cout):
        wb(p + "norm1", cin)
        wb(p + "conv1", cout, cin, 3, 3)
        wb(p + "norm2", cout)
        wb(p + "conv2", cout, cout, 3, 3)
        if cin != cout:
            wb(p + "conv_shortcut", cout, cin, 1, 1)

    def mid(p, c):
        resnet(p + "mid_block.resnets.0.", c, c)
        wb(p + "mid_block.attentions.0.group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            wb(p + "mid_block.attentions.0." + n, c, c)
        resnet(p + "mid_block.resnets.1.", c, c)

    wb("encoder.conv_in", boc[0], cfg["in_channels"], 3, 3)
    c = boc[0]
    for i, co in enumerate(boc):
        for j in range(lpb):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", c, co)
            c = co
        if i != len(boc) - 1:
            wb(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3, 3)
    mid("encoder.", c)
    wb("encoder.conv_norm_out", c)
    wb("encoder.conv_out", 2 * lat, c, 3, 3)
    rev = boc[::-1]
    wb("decoder.conv_in", rev[0], lat, 3, 3)
    mid("decoder.", rev[0])
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(lpb + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", c, co)
            c = co
        if i != len(rev) - 1:
            wb(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3, 3)
    wb("decoder.conv_norm_out", c)
    wb("decoder.conv_out", cfg["out_channels"], c, 3, 3)
    return out


def random_vae_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights (fan-in scaled, so activations stay O(1) through the stack); CPU fp32."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in vae_param_shapes(cfg).items():
        if k.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * fan_in ** -0.5
        sd[k] = t
    return sd


class AutoencoderKL:
    """Drop-in for ``diffusers.AutoencoderKL`` on the calls the RepText pipelines make.

    ``state_dict`` uses diffusers' parameter names (see ``oracle/vae_oracle.py: param_shapes``).  ``conv_impl``:
    ``"implicit"`` (product: TMA implicit GEMM) or ``"im2col"`` (gather + plain GEMM for every 3x3 convolution; A/B).
    """

    def __init__(self, config: Optional[dict], state_dict: Dict[str, torch.Tensor], dtype=torch.bfloat16,
                 device="cuda", conv_impl: str = "implicit"):
        if dtype != torch.bfloat16:
            raise ValueError("the VAE path computes in bfloat16 (fp32 accumulation)")
        cfg = dict(FLUX_VAE_CONFIG)
        cfg.update(config or {})
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        for c in cfg["block_out_channels"]:
            if c % 64 or c > 512:
                raise ValueError("block_out_channels must be multiples of 64, at most 512")
        self.config = FrozenConfig(**cfg)
        self.dtype, self.device = dtype, torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("reptext_b200 needs a CUDA device (there is no CPU path)")
        if conv_impl not in ("implicit", "im2col"):
            raise ValueError("conv_impl must be 'implicit' or 'im2col'")
        self.conv_impl = conv_impl
        L.lib()
        self._w: Dict[str, torch.Tensor] = {}
        self._prepare(state_dict)

    # ------------------------------------------------------------------ weights
    def _prepare(self, sd: Dict[str, torch.Tensor]) -> None:
        """diffusers' tensors -> GEMM operands: 3x3 kernels tap-major ``[Cout_pad, 9 * Cin_pad]`` (``ops.pack_conv3x3_weight``),
        1x1 kernels ``[Cout, Cin]``, biases padded with the output channels; norms and linears as they are."""
        dev, dt = self.device, self.dtype
        for name, t in sd.items():
            t = t.detach().to(dev, torch.float32)
            if name.endswith(".weight") and t.dim() == 4:
                co, ci, kh, kw = t.shape
                base = name[:-len(".weight")]
                n_pad = _pad64(co)
                if (kh, kw) == (3, 3):
                    if name == "encoder.conv_in.weight":      # 3 input channels: im2col rows of 27 (-> 32) values
                        w = torch.zeros(n_pad, 32, device=dev)
                        w[:co, :9 * ci] = t.permute(0, 2, 3, 1).reshape(co, 9 * ci)
                    else:
                        w = ops.pack_conv3x3_weight(t, _pad64(ci), n_pad)
                elif (kh, kw) == (1, 1):
                    w = torch.zeros(n_pad, ci, device=dev)
                    w[:co] = t[:, :, 0, 0]
                else:
                    raise ValueError(f"{name}: unsupported kernel size {kh}x{kw}")
                self._w[name] = w.to(dt).contiguous()
                b = torch.zeros(n_pad, device=dev)
                if base + ".bias" in sd:
                    b[:co] = sd[base + ".bias"].detach().to(dev, torch.float32)
                self._w[base + ".bias"] = b.to(dt).contiguous()
            elif name.endswith(".bias") and name[:-len(".bias")] + ".weight" in sd and \
                    sd[name[:-len(".bias")] + ".weight"].dim() == 4:
                continue                                        # handled with its convolution
            else:
                self._w[name] = t.to(dt).contiguous()
        for side in ("encoder.", "decoder."):
            a = side + "mid_block.attentions.0."
            if a + "to_v.bias" in sd:
                # softmax rows sum to one: P (V + 1 b_v^T) = P V + 1 b_v^T, so b_v goes through to_out once
                wo = sd[a + "to_out.0.weight"].detach().to(dev, torch.float32)
                bv = sd[a + "to_v.bias"].detach().to(dev, torch.float32)
                bo = sd[a + "to_out.0.bias"].detach().to(dev, torch.float32)
                self._w[a + "to_out.0.bias_folded"] = (wo @ bv + bo).to(dt).contiguous()

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, torch_dtype=torch.bfloat16,
                        device="cuda", variant: Optional[str] = None, conv_impl: str = "implicit", **unused) -> "AutoencoderKL":
        """``<dir>[/subfolder]/config.json`` + ``diffusion_pytorch_model.safetensors`` of a diffusers ``AutoencoderKL``
        (e.g. ``black-forest-labs/FLUX.1-dev``'s ``vae/``) from a LOCAL directory (:mod:`reptext_b200.checkpoint`).  The
        architecture must be the FLUX one this path implements: every parameter of :func:`vae_param_shapes` present with
        its shape, nothing else (no quant / post-quant convolutions)."""
        from . import checkpoint as ck
        d = ck.resolve_dir(pretrained_model_name_or_path, subfolder)
        raw, stored_cls = ck.read_config(d)
        if stored_cls not in (None, "AutoencoderKL"):
            raise ValueError(f"{d!r} holds a {stored_cls}, not an AutoencoderKL")
        cfg = dict(FLUX_VAE_CONFIG)
        cfg.update(raw)
        if cfg.get("use_quant_conv") or cfg.get("use_post_quant_conv"):
            raise ValueError("AutoencoderKL with quant_conv / post_quant_conv is not the FLUX VAE this path implements")
        sd = ck.load_state_dict(d, (ck.DIFFUSERS_STEM,), variant)
        want = vae_param_shapes(cfg)
        missing = [k for k in want if k not in sd]
        unexpected = [k for k in sd if k not in want]
        wrong = [k for k in want if k in sd and tuple(sd[k].shape) != want[k]]
        if missing or unexpected or wrong:
            raise RuntimeError(f"{d!r} is not a FLUX-architecture AutoencoderKL: missing {missing[:3]}, unexpected "
                               f"{unexpected[:3]}, wrong shape {wrong[:3]}")
        return cls(cfg, sd, dtype=torch_dtype, device=device, conv_impl=conv_impl)

    def save_pretrained(self, save_directory) -> None:
        """Not available: the constructor re-packs the convolution kernels for the implicit GEMM and drops the originals;
        keep the directory the model was loaded from."""
        raise NotImplementedError(self.save_pretrained.__doc__)

    @classmethod
    def random_init(cls, config: Optional[dict] = None, seed: int = 0, dtype=torch.bfloat16, device="cuda",
                    conv_impl: str = "implicit") -> "AutoencoderKL":
        """Random weights of the configured architecture (there is no network for checkpoints; BASELINE.json)."""
        cfg = dict(FLUX_VAE_CONFIG)
        cfg.update(config or {})
        return cls(cfg, random_vae_state_dict(cfg, seed), dtype=dtype, device=device, conv_impl=conv_impl)

    def _p(self, name: str) -> torch.Tensor:
        try:
            return self._w[name]
        except KeyError:
            raise ValueError(f"AutoencoderKL state dict has no '{name}'") from None

    def to(self, *a, **k):
        return self

    def enable_slicing(self):     # one image at 1024^2 needs < 4 GB of activations; nothing to slice on 180 GB
        pass

    def enable_tiling(self):
        pass

    # ------------------------------------------------------------------ blocks (x: [B, H * W, C] bf16, NHWC)
    @staticmethod
    def _check_hw(hw: Tuple[int, int]) -> None:
        h, w = hw
        ok = w >= 8 and (w % 128 == 0 or 128 % w == 0) and (w >= 128 or h % (128 // w) == 0)
        if not ok:
            raise ValueError(f"VAE feature map {h}x{w}: the width must be >= 8 and divide (or be a multiple of) 128, "
                             f"and the height a multiple of 128 / width")

    @staticmethod
    def _hw_ok(hw: Tuple[int, int]) -> bool:
        try:
            AutoencoderKL._check_hw(hw)
            return True
        except ValueError:
            return False

    def _conv3(self, x, hw, name, residual_into=None):
        """3x3 / stride 1 / pad 1.  Feature maps whose 128-pixel tiles are not whole rows or whole-row groups (e.g. 192
        wide at 1536^2) take the gather + GEMM form - the same arithmetic, one more pass over memory."""
        w, b = self._p(name + ".weight"), self._p(name + ".bias")
        if self.conv_impl == "implicit" and self._hw_ok(hw):
            return ops.conv3x3(x, hw, w, b, residual_into=residual_into)
        col = ops.im2col3x3_nhwc(x, hw, x.shape[2], hw, 1, 1, Kp=9 * x.shape[2])
        return self._linear(col, w, b, residual_into=residual_into)

    def _linear(self, x, w, b, residual_into=None, out=None):
        B = x.shape[0]
        if residual_into is not None:
            seg = ops.Segment(W=w, bias=b, out=residual_into, mode=L.EPI_GATE_RESID)
            out = residual_into
        else:
            if out is None:
                out = torch.empty(B, x.shape[1], w.shape[0], dtype=x.dtype, device=x.device)
            seg = ops.Segment(W=w, bias=b, out=out, mode=L.EPI_BIAS)
        ops.gemm([ops.Problem(A=x, segs=[seg])], B, x.dtype)
        return out

    def _gn(self, x, name, silu):
        return ops.groupnorm_nhwc(x, self.config.norm_num_groups, self._p(name + ".weight"), self._p(name + ".bias"),
                                  eps=1e-6, silu=silu)

    def _resnet(self, x, hw, p):
        """ResnetBlock2D without a time embedding: x + conv2(silu(norm2(conv1(silu(norm1(x)))))), 1x1 shortcut when the
        channel count changes.  The sum happens in conv2's epilogue (in place on the skip tensor)."""
        h = self._conv3(self._gn(x, p + "norm1", True), hw, p + "conv1")
        h = self._gn(h, p + "norm2", True)
        if (p + "conv_shortcut.weight") in self._w:
            x = self._linear(x, self._p(p + "conv_shortcut.weight"), self._p(p + "conv_shortcut.bias"))
        return self._conv3(h, hw, p + "conv2", residual_into=x)

    def _attention(self, x, p):
        """The mid block's single-head attention over all H * W positions (GroupNorm first, residual connection)."""
        B, S, Cc = x.shape
        t = self._gn(x, p + "group_norm", False)
        q = torch.empty(B, S, Cc, dtype=x.dtype, device=x.device)
        k = torch.empty_like(q)
        ops.gemm([ops.Problem(A=t, segs=[
            ops.Segment(W=self._p(p + "to_q.weight"), bias=self._p(p + "to_q.bias"), out=q),
            ops.Segment(W=self._p(p + "to_k.weight"), bias=self._p(p + "to_k.bias"), out=k)])], B, x.dtype)
        wv = self._p(p + "to_v.weight")[None]                                         # [1, C, C]
        vt = torch.empty(1, Cc, S, dtype=x.dtype, device=x.device)
        # The logits stay in fp32 from the accumulator to the softmax, like the reference's SDPA (a bf16 rounding of a
        # logit of magnitude 16 would already be a 3 % error of its weight); queries go through in chunks so that the fp32
        # score block stays <= 1 GB whatever the image size.
        rows = max(256, min(S, (1 << 28) // S // 256 * 256))
        sc = torch.empty(1, rows, S, dtype=torch.float32, device=x.device)
        pr = torch.empty(1, rows, S, dtype=x.dtype, device=x.device)
        o = torch.empty(B, S, Cc, dtype=x.dtype, device=x.device)
        for b in range(B):
            # v^T [C, S] = W_v [C, C] . t^T  (bias folded into to_out)
            ops.gemm([ops.Problem(A=wv, segs=[ops.Segment(W=t[b], out=vt)])], 1, x.dtype)
            for r0 in range(0, S, rows):
                n = min(rows, S - r0)
                # S = q k^T / sqrt(C), fp32 out
                ops.gemm([ops.Problem(A=q[b:b + 1], a_row0=r0, m_rows=n, scale=Cc ** -0.5,
                                      segs=[ops.Segment(W=k[b], out=sc, mode=L.EPI_SCALE_MASK, out_f32=True)])], 1, x.dtype)
                ops.softmax_rows_f32(sc[0, :n], pr[0, :n])
                ops.gemm([ops.Problem(A=pr, m_rows=n, out_row0=r0, segs=[ops.Segment(W=vt[0], out=o[b:b + 1])])], 1, x.dtype)
        return self._linear(o, self._p(p + "to_out.0.weight"), self._p(p + "to_out.0.bias_folded"), residual_into=x)

    def _mid(self, x, hw, side):
        x = self._resnet(x, hw, side + "mid_block.resnets.0.")
        x = self._attention(x, side + "mid_block.attentions.0.")
        return self._resnet(x, hw, side + "mid_block.resnets.1.")

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """[B, 3, H, W] in [-1, 1] -> posterior over [B, latent, H / 8, W / 8]."""
        cfg = self.config
        if x.dim() != 4 or x.shape[1] != cfg.in_channels:
            raise ValueError(f"encode expects [B, {cfg.in_channels}, H, W], got {tuple(x.shape)}")
        nb = len(cfg.block_out_channels)
        B, Cin, H, W = x.shape
        if H % (1 << (nb - 1)) or W % (1 << (nb - 1)):
            raise ValueError(f"image size {H}x{W} must be a multiple of {1 << (nb - 1)}")
        x = x.to(self.device)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        hw = (H, W)
        h = ops.nchw_to_nhwc(x, 8)
        col = ops.im2col3x3_nhwc(h, hw, Cin, hw, 1, 1, Kp=32)
        h = self._linear(col, self._p("encoder.conv_in.weight"), self._p("encoder.conv_in.bias"))
        del col
        for i in range(nb):
            for j in range(cfg.layers_per_block):
                h = self._resnet(h, hw, f"encoder.down_blocks.{i}.resnets.{j}.")
            if i != nb - 1:
                # Downsample2D(padding=0): zero row / column appended at the bottom / right, 3x3 stride-2 convolution
                n = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                hw2 = (hw[0] // 2, hw[1] // 2)
                col = ops.im2col3x3_nhwc(h, hw, h.shape[2], hw2, 2, 0, Kp=9 * h.shape[2])
                h = self._linear(col, self._p(n + ".weight"), self._p(n + ".bias"))
                del col
                hw = hw2
        h = self._mid(h, hw, "encoder.")
        h = self._gn(h, "encoder.conv_norm_out", True)
        mom = self._conv3(h, hw, "encoder.conv_out")            # [B, hw, pad64(2 * latent)]
        post = DiagonalGaussianDistribution(mom, hw, cfg.latent_channels, self.dtype)
        return FrozenConfig(latent_dist=post) if return_dict else (post,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        """[B, latent, h, w] -> [B, 3, 8h, 8w]."""
        cfg = self.config
        if z.dim() != 4 or z.shape[1] != cfg.latent_channels:
            raise ValueError(f"decode expects [B, {cfg.latent_channels}, h, w], got {tuple(z.shape)}")
        nb = len(cfg.block_out_channels)
        z = z.to(self.device)
        if z.dtype not in (torch.float32, torch.bfloat16):
            z = z.float()
        hw = (z.shape[2], z.shape[3])
        h = ops.nchw_to_nhwc(z, _pad64(cfg.latent_channels))
        h = self._conv3(h, hw, "decoder.conv_in")
        h = self._mid(h, hw, "decoder.")
        for i in range(nb):
            for j in range(cfg.layers_per_block + 1):
                h = self._resnet(h, hw, f"decoder.up_blocks.{i}.resnets.{j}.")
            if i != nb - 1:                                      # Upsample2D: nearest x2, then a 3x3 convolution
                h = ops.upsample_nearest2x_nhwc(h, hw)
                hw = (2 * hw[0], 2 * hw[1])
                h = self._conv3(h, hw, f"decoder.up_blocks.{i}.upsamplers.0.conv")
        h = self._gn(h, "decoder.conv_norm_out", True)
        h = self._conv3(h, hw, "decoder.conv_out")              # [B, HW, 64], first out_channels valid
        img = ops.nhwc_to_nchw(h, hw, cfg.out_channels, self.dtype)
        return FrozenConfig(sample=img) if return_dict else (img,)

    def __call__(self, sample: torch.Tensor, sample_posterior: bool = False, generator=None):
        post = self.encode(sample).latent_dist
        z = post.sample(generator) if sample_posterior else post.mode()
        return self.decode(z)
