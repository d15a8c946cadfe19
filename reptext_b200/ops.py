"""Operator-level entry points of the runtime (thin wrappers over the C-ABI; device tensors in / out).

These are what the parity tests and ``bench.py``'s roofline leg call; the model-level path
(``reptext_b200.models``) drives the same kernels from C++ without coming back to Python per op.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import _lib as L

IMPL_AUTO, IMPL_SIMT, IMPL_TC1, IMPL_TC2 = 0, 1, 2, 3


@dataclass
class Segment:
    W: torch.Tensor                      # [n, K]
    out: torch.Tensor                    # [batch, rows, out_ld]
    bias: Optional[torch.Tensor] = None
    mode: int = L.EPI_BIAS
    out_col0: int = 0
    norm_w: Optional[torch.Tensor] = None
    scatter: bool = False                # sequence-parallel: head blocks go to gemm(..., sp_out=[...]) buffers
    out_f32: bool = False                # `out` is float32: the accumulator leaves without a bf16 rounding


@dataclass
class Problem:
    A: torch.Tensor                      # [batch or 1, a_rows_total, a_ld]
    segs: List[Segment] = field(default_factory=list)
    a_row0: int = 0
    m_rows: Optional[int] = None
    out_row0: int = 0
    K: Optional[int] = None
    gate: Optional[torch.Tensor] = None  # [batch, gate_ld] fp32
    extra: Optional[torch.Tensor] = None  # [batch, rows, extra_ld]
    extra_row0: int = 0
    scale: float = 1.0
    mask: Optional[torch.Tensor] = None
    accumulate: bool = False
    broadcast_a: bool = False
    conv_hw: Optional[Sequence[int]] = None  # (H, W): A is an NHWC image [batch, H * W, C]; 3x3 / stride 1 / pad 1
    conv_c: Optional[int] = None             # channels used (default A.shape[2])


def _fill_problem(dst: L.GemmProblem, p: Problem, keep: list) -> None:
    A = p.A
    assert A.dim() == 3 and A.stride(2) == 1
    dst.A = L.ptr(A)
    dst.a_batch_stride = 0 if p.broadcast_a else A.stride(0)
    dst.a_ld = A.stride(1)
    dst.a_row0 = p.a_row0
    dst.a_rows_total = A.shape[1]
    dst.m_rows = p.m_rows if p.m_rows is not None else A.shape[1] - p.a_row0
    dst.out_row0 = p.out_row0
    dst.K = p.K if p.K is not None else A.shape[2]
    dst.nseg = len(p.segs)
    n0 = 0
    for i, s in enumerate(p.segs):
        sg = dst.seg[i]
        assert s.W.is_contiguous() and s.W.shape[1] == dst.K
        sg.W = L.ptr(s.W)
        sg.bias = L.ptr(s.bias)
        sg.n_begin = n0
        n0 += s.W.shape[0]
        sg.n_end = n0
        sg.mode = s.mode
        sg.out = L.ptr(s.out)
        sg.out_batch_stride = s.out.stride(0)
        sg.out_ld = s.out.stride(1)
        sg.out_col0 = s.out_col0
        sg.norm_w = L.ptr(s.norm_w)
        sg.scatter = int(s.scatter)
        sg.out_f32 = int(s.out_f32)
        assert s.scatter or s.out.dtype == (torch.float32 if s.out_f32 else A.dtype)
    dst.gate = L.ptr(p.gate)
    dst.gate_ld = p.gate.stride(0) if p.gate is not None else 0
    dst.extra = L.ptr(p.extra)
    if p.extra is not None:
        dst.extra_batch_stride = p.extra.stride(0)
        dst.extra_ld = p.extra.stride(1)
    dst.extra_row0 = p.extra_row0
    dst.scale = p.scale
    dst.mask = L.ptr(p.mask)
    dst.accumulate = int(p.accumulate)
    if p.conv_hw is not None:
        dst.conv_h, dst.conv_w = int(p.conv_hw[0]), int(p.conv_hw[1])
        dst.conv_c = int(p.conv_c if p.conv_c is not None else A.shape[2])
        assert dst.conv_h * dst.conv_w == A.shape[1] and p.a_row0 == 0
    keep.append(p)


def gemm(problems: Sequence[Problem], batch: int, dtype: torch.dtype, rope: Optional[torch.Tensor] = None,
         head_dim: int = 0, impl: int = IMPL_AUTO, sp_out: Optional[Sequence[torch.Tensor]] = None, sp_cols: int = 0,
         sp_row0: int = 0) -> None:
    """One launch of the (grouped) projection GEMM with fused epilogues; see ``rt_gemm``.  ``sp_out`` / ``sp_cols`` /
    ``sp_row0``: destinations of ``Segment(scatter=True)`` segments (their ``out`` only supplies the strides)."""
    g = L.GemmLaunch()
    g.dtype = L.dtype_code(dtype)
    g.batch = batch
    g.nprob = len(problems)
    keep: list = []
    for i, p in enumerate(problems):
        _fill_problem(g.prob[i], p, keep)
    g.rope = L.ptr(rope)
    g.head_dim = head_dim
    g.sp_cols, g.sp_row0 = sp_cols, sp_row0
    for i, t in enumerate(sp_out or []):
        g.sp_out[i] = L.ptr(t)
    L.check(L.lib().rt_gemm(C.byref(g), impl, L.stream_ptr()))


def linear(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, mode: int = L.EPI_BIAS,
           impl: int = IMPL_AUTO) -> torch.Tensor:
    """``F.linear`` (+ optional GELU-tanh) on a [batch, rows, K] tensor."""
    out = torch.empty(x.shape[0], x.shape[1], W.shape[0], dtype=x.dtype, device=x.device)
    gemm([Problem(A=x, segs=[Segment(W=W, bias=bias, out=out, mode=mode)])], x.shape[0], x.dtype, impl=impl)
    return out


def attention(qkv: torch.Tensor, heads: int, hd: int, q_col0: int, k_col0: int, v_col0: int,
              out: Optional[torch.Tensor] = None, out_col0: int = 0, impl: int = IMPL_AUTO,
              sp_out: Optional[Sequence[torch.Tensor]] = None, sp_rows: int = 0) -> torch.Tensor:
    """Joint non-causal attention over [batch, S, ld] with q/k/v at column offsets (head-major).  ``sp_out`` /
    ``sp_rows``: sequence-parallel row scatter (output row r goes to ``sp_out[r // sp_rows]``; ``out`` only supplies
    the strides)."""
    B, S, _ = qkv.shape
    if out is None:
        out = torch.empty(B, S, heads * hd, dtype=qkv.dtype, device=qkv.device)
    a = L.AttentionArgs()
    a.dtype = L.dtype_code(qkv.dtype)
    a.qkv = L.ptr(qkv)
    a.batch_stride = qkv.stride(0)
    a.ld = qkv.stride(1)
    a.q_col0, a.k_col0, a.v_col0 = q_col0, k_col0, v_col0
    a.out = L.ptr(out)
    a.out_batch_stride = out.stride(0)
    a.out_ld = out.stride(1)
    a.out_col0 = out_col0
    a.batch, a.S, a.heads, a.hd = B, S, heads, hd
    a.sp_rows = sp_rows
    for i, t in enumerate(sp_out or []):
        a.sp_out[i] = L.ptr(t)
    L.check(L.lib().rt_attention(C.byref(a), impl, L.stream_ptr()))
    return out


def layernorm_modulate(x: torch.Tensor, groups, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """groups: [(row_begin, row_end, shift[batch, ld] fp32, scale[batch, ld] fp32)] (1 or 2 groups)."""
    if out is None:
        out = torch.empty_like(x)
    arr = (L.LnModGroup * len(groups))()
    for i, (r0, r1, sh, sc) in enumerate(groups):
        assert sh.dtype == torch.float32 and sc.dtype == torch.float32 and sh.stride(0) == sc.stride(0)
        arr[i].row_begin, arr[i].row_end = r0, r1
        arr[i].shift, arr[i].scale, arr[i].ld = L.ptr(sh), L.ptr(sc), sh.stride(0)
    L.check(L.lib().rt_layernorm_modulate(L.dtype_code(x.dtype), L.ptr(x), x.stride(0), x.stride(1), L.ptr(out),
                                          out.stride(0), out.stride(1), x.shape[0], x.shape[2], len(groups), arr,
                                          L.stream_ptr()))
    return out


def rope_table(ids: torch.Tensor, axes_dims: Sequence[int]) -> torch.Tensor:
    """FluxPosEmbed: ids [S, 3] fp32 -> [S, sum(axes)/2, 2] fp32 (cos, sin) per rotation pair."""
    ids = ids.to(torch.float32).contiguous()
    S = ids.shape[0]
    out = torch.empty(S, sum(axes_dims) // 2, 2, dtype=torch.float32, device=ids.device)
    ax = (C.c_int * 3)(*axes_dims)
    L.check(L.lib().rt_rope_table(L.ptr(ids), S, ax, L.ptr(out), L.stream_ptr()))
    return out


def qknorm_rope_(buf: torch.Tensor, col0: int, heads: int, hd: int, norm_w: torch.Tensor,
                 rope: Optional[torch.Tensor], row0: int = 0, rows: Optional[int] = None, rope_row0: int = 0) -> None:
    rows = buf.shape[1] - row0 if rows is None else rows
    L.check(L.lib().rt_qknorm_rope(L.dtype_code(buf.dtype), L.ptr(buf), buf.stride(0), buf.stride(1), col0,
                                   buf.shape[0], row0, rows, heads, hd, L.ptr(norm_w), L.ptr(rope), rope_row0,
                                   L.stream_ptr()))


def euler_step(model_output: torch.Tensor, sample: torch.Tensor, sigma: float, sigma_next: float,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FlowMatchEulerDiscreteScheduler.step arithmetic (RepText/pipeline_flux_controlnet.py:1109)."""
    assert model_output.is_contiguous() and sample.is_contiguous() and model_output.dtype == sample.dtype
    if out is None:
        out = torch.empty_like(model_output)
    L.check(L.lib().rt_euler_step(L.dtype_code(sample.dtype), L.ptr(model_output), L.ptr(sample), L.ptr(out),
                                  sample.numel(), float(sigma), float(sigma_next), L.stream_ptr()))
    return out


def cfg_combine(v2: torch.Tensor, true_guidance_scale: float, zero_pred: bool) -> torch.Tensor:
    """pipeline_flux_controlnet_inpaint.py:1264-1270 on v2 = cat([uncond, text])."""
    assert v2.is_contiguous() and v2.shape[0] % 2 == 0
    out = torch.empty((v2.shape[0] // 2,) + tuple(v2.shape[1:]), dtype=v2.dtype, device=v2.device)
    L.check(L.lib().rt_cfg_combine(L.dtype_code(v2.dtype), L.ptr(v2), L.ptr(out), out.numel(),
                                   float(true_guidance_scale), int(zero_pred), L.stream_ptr()))
    return out


def cfg_euler_step(v2: torch.Tensor, sample: torch.Tensor, true_guidance_scale: float, zero_pred: bool,
                   sigma: float, sigma_next: float) -> torch.Tensor:
    assert v2.is_contiguous() and sample.is_contiguous() and v2.numel() == 2 * sample.numel()
    out = torch.empty_like(sample)
    L.check(L.lib().rt_cfg_euler_step(L.dtype_code(sample.dtype), L.ptr(v2), L.ptr(sample), L.ptr(out),
                                      sample.numel(), float(true_guidance_scale), int(zero_pred), float(sigma),
                                      float(sigma_next), L.stream_ptr()))
    return out


def mask_scale_add(x: torch.Tensor, mask: Optional[torch.Tensor], acc_in: Optional[torch.Tensor],
                   scale: float = 1.0) -> torch.Tensor:
    """mask[r] * scale * x[b, r, :] (+ acc_in): the pipelines' regional-mask multiply and multi-line sum."""
    assert x.is_contiguous() and x.dim() == 3
    out = torch.empty_like(x)
    m = mask.reshape(-1).contiguous() if mask is not None else None
    L.check(L.lib().rt_mask_scale_add(L.dtype_code(x.dtype), L.ptr(x), L.ptr(m), L.ptr(acc_in), L.ptr(out),
                                      x.shape[0], x.shape[1], x.shape[2], float(scale), L.stream_ptr()))
    return out


def glyph_init_blend(noise: torch.Tensor, glyph_latents: torch.Tensor, mask_u8: torch.Tensor,
                     w_glyph: float = 0.10, w_noise: float = 1.00) -> torch.Tensor:
    assert noise.is_contiguous() and glyph_latents.is_contiguous() and mask_u8.is_contiguous()
    assert mask_u8.dtype == torch.uint8 and mask_u8.numel() == noise.numel()
    out = torch.empty_like(noise)
    L.check(L.lib().rt_glyph_init_blend(L.dtype_code(noise.dtype), L.ptr(noise), L.ptr(glyph_latents),
                                        L.ptr(mask_u8), L.ptr(out), noise.numel(), float(w_glyph), float(w_noise),
                                        L.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------------
# VAE operators; activations are NHWC bf16 tensors shaped [batch, H * W, C]
# ------------------------------------------------------------------------------------------------------
def conv3x3(x: torch.Tensor, hw: Sequence[int], Wk: torch.Tensor, bias: Optional[torch.Tensor],
            out: Optional[torch.Tensor] = None, residual_into: Optional[torch.Tensor] = None,
            impl: int = IMPL_AUTO) -> torch.Tensor:
    """3x3 / stride 1 / padding 1 convolution as ONE implicit-GEMM launch.  ``Wk``: [Cout, 9 * C] tap-major weights
    (``pack_conv3x3_weight``), C = x.shape[2] a multiple of 64.  ``residual_into``: accumulate ``conv + bias`` into
    that tensor in place (the ResnetBlock2D skip connection) instead of writing ``out``."""
    B, HW, Cc = x.shape
    assert Cc % 64 == 0 and Wk.shape[1] == 9 * Cc and HW == hw[0] * hw[1]
    if residual_into is not None:
        seg = Segment(W=Wk, bias=bias, out=residual_into, mode=L.EPI_GATE_RESID)
        out = residual_into
    else:
        if out is None:
            out = torch.empty(B, HW, Wk.shape[0], dtype=x.dtype, device=x.device)
        seg = Segment(W=Wk, bias=bias, out=out, mode=L.EPI_BIAS)
    gemm([Problem(A=x, segs=[seg], K=9 * Cc, conv_hw=hw)], B, x.dtype, impl=impl if impl != IMPL_AUTO else IMPL_AUTO)
    return out


def pack_conv3x3_weight(w: torch.Tensor, c_pad: Optional[int] = None, n_pad: Optional[int] = None) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [n_pad, 9 * c_pad], K index (ky * 3 + kx) * c_pad + channel, zero padded."""
    co, ci = w.shape[:2]
    c_pad = c_pad or (ci + 63) // 64 * 64
    n_pad = n_pad or (co + 63) // 64 * 64
    t = torch.zeros(n_pad, 3, 3, c_pad, dtype=w.dtype, device=w.device)
    t[:co, :, :, :ci] = w.permute(0, 2, 3, 1)
    return t.reshape(n_pad, 9 * c_pad).contiguous()


def groupnorm_nhwc(x: torch.Tensor, groups: int, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6,
                   silu: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, HW, Cc = x.shape
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    if out is None:
        out = torch.empty_like(x)
    ws = torch.empty(B * groups * 3, dtype=torch.float64, device=x.device)
    L.check(L.lib().rt_groupnorm_nhwc(L.ptr(x), L.ptr(out), B, HW, Cc, groups, L.ptr(gamma), L.ptr(beta), float(eps),
                                      int(silu), L.ptr(ws), L.stream_ptr()))
    return out


def upsample_nearest2x_nhwc(x: torch.Tensor, hw: Sequence[int]) -> torch.Tensor:
    B, HW, Cc = x.shape
    assert x.is_contiguous() and HW == hw[0] * hw[1]
    out = torch.empty(B, 4 * HW, Cc, dtype=x.dtype, device=x.device)
    L.check(L.lib().rt_upsample_nearest2x_nhwc(L.ptr(x), L.ptr(out), B, hw[0], hw[1], Cc, L.stream_ptr()))
    return out


def softmax_rows_f32(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """Softmax over the last dimension of a float32 [rows, cols] matrix, written as bf16 to ``out`` [rows, cols] (row
    strides free): the logits never see a bf16 rounding."""
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16 and x.stride(-1) == 1 and out.stride(-1) == 1
    assert x.dim() == 2 and out.shape == x.shape
    L.check(L.lib().rt_softmax_rows_f32(L.ptr(x), x.shape[0], x.shape[1], x.stride(0), L.ptr(out), out.stride(0),
                                        L.stream_ptr()))
    return out


def softmax_rows_(x: torch.Tensor) -> torch.Tensor:
    """In-place softmax over the last dimension of a [rows, cols] (or [B, rows, cols] contiguous) bf16 matrix."""
    assert x.stride(-1) == 1 and x.dtype == torch.bfloat16
    rows = x.numel() // x.shape[-1]
    assert x.is_contiguous()
    L.check(L.lib().rt_softmax_rows(L.ptr(x), rows, x.shape[-1], x.stride(-2), L.stream_ptr()))
    return x


def im2col3x3_nhwc(x: torch.Tensor, hw: Sequence[int], C_used: int, out_hw: Sequence[int], stride: int, pad_lo: int,
                   Kp: Optional[int] = None) -> torch.Tensor:
    B, HW, ld = x.shape
    assert x.is_contiguous() and HW == hw[0] * hw[1]
    Kp = Kp or (9 * C_used + 7) // 8 * 8
    out = torch.empty(B, out_hw[0] * out_hw[1], Kp, dtype=x.dtype, device=x.device)
    L.check(L.lib().rt_im2col3x3_nhwc(L.ptr(x), L.ptr(out), B, hw[0], hw[1], C_used, ld, out_hw[0], out_hw[1], stride,
                                      pad_lo, Kp, L.stream_ptr()))
    return out


def nchw_to_nhwc(x: torch.Tensor, c_pad: int) -> torch.Tensor:
    B, Cc, H, W = x.shape
    x = x.contiguous()
    out = torch.empty(B, H * W, c_pad, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().rt_nchw_to_nhwc(L.dtype_code(x.dtype), L.ptr(x), L.ptr(out), B, Cc, H * W, c_pad, L.stream_ptr()))
    return out


def nhwc_to_nchw(x: torch.Tensor, hw: Sequence[int], C_used: int, dtype: torch.dtype) -> torch.Tensor:
    B, HW, ld = x.shape
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    out = torch.empty(B, C_used, hw[0], hw[1], dtype=dtype, device=x.device)
    L.check(L.lib().rt_nhwc_to_nchw(L.ptr(x), ld, L.ptr(out), L.dtype_code(dtype), B, C_used, HW, L.stream_ptr()))
    return out


def vae_posterior_sample(moments: torch.Tensor, hw: Sequence[int], latent_channels: int,
                         noise: Optional[torch.Tensor], dtype: torch.dtype) -> torch.Tensor:
    B, HW, ld = moments.shape
    assert moments.is_contiguous() and moments.dtype == torch.bfloat16
    out = torch.empty(B, latent_channels, hw[0], hw[1], dtype=dtype, device=moments.device)
    if noise is not None:
        noise = noise.to(dtype).contiguous()
        assert noise.shape == out.shape
    L.check(L.lib().rt_vae_posterior_sample(L.ptr(moments), ld, latent_channels, B, HW, L.ptr(noise), L.ptr(out),
                                            L.dtype_code(dtype), L.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------------
# Prompt-encoder operators; bf16 tensors shaped [batch, S, D]
# ------------------------------------------------------------------------------------------------------
def norm_rows(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float,
              subtract_mean: bool) -> torch.Tensor:
    """T5LayerNorm (``subtract_mean=False``, no bias) or nn.LayerNorm over the last dimension."""
    assert x.stride(-1) == 1 and x.is_contiguous() and x.dtype == torch.bfloat16
    D = x.shape[-1]
    out = torch.empty_like(x)
    L.check(L.lib().rt_norm_rows(L.ptr(x), D, L.ptr(out), D, x.numel() // D, D, L.ptr(weight), L.ptr(bias), float(eps),
                                 int(subtract_mean), L.stream_ptr()))
    return out


def text_attention(qkv: torch.Tensor, heads: int, scale: float, rel_bias: Optional[torch.Tensor] = None,
                   causal: bool = False) -> torch.Tensor:
    """qkv [B, S, 3 * heads * 64] (q | k | v, head-major) -> [B, S, heads * 64]."""
    B, S, W3 = qkv.shape
    D = heads * 64
    assert W3 == 3 * D and qkv.is_contiguous() and qkv.dtype == torch.bfloat16
    if rel_bias is not None:
        assert rel_bias.dtype == torch.float32 and rel_bias.is_contiguous() and tuple(rel_bias.shape) == (heads, 2 * S - 1)
    out = torch.empty(B, S, D, dtype=qkv.dtype, device=qkv.device)
    L.check(L.lib().rt_text_attention(L.ptr(qkv), qkv.stride(0), qkv.stride(1), 0, D, 2 * D, L.ptr(out), out.stride(0),
                                      out.stride(1), 0, B, S, heads, 64, float(scale), L.ptr(rel_bias), int(causal),
                                      L.stream_ptr()))
    return out


def glu_act(x: torch.Tensor, F_out: int, kind: int) -> torch.Tensor:
    """kind 0: x[..., :F] * x[..., F:2F]; kind 1: quick_gelu(x[..., :F])."""
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    rows = x.numel() // x.shape[-1]
    out = torch.empty(*x.shape[:-1], F_out, dtype=x.dtype, device=x.device)
    L.check(L.lib().rt_glu_act(kind, L.ptr(x), x.shape[-1], L.ptr(out), F_out, rows, F_out, L.stream_ptr()))
    return out


def embedding(table: torch.Tensor, ids: torch.Tensor, pos_table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """table[ids] (+ pos_table[position]) for ids [B, S] (int64); raises IndexError on ids outside the table."""
    assert table.is_contiguous() and table.dtype == torch.bfloat16 and ids.dim() == 2
    ids = ids.to(device=table.device, dtype=torch.int64).contiguous()
    B, S = ids.shape
    out = torch.empty(B, S, table.shape[1], dtype=table.dtype, device=table.device)
    bad = torch.zeros(1, dtype=torch.int32, device=table.device)
    L.check(L.lib().rt_embedding(L.ptr(table), table.shape[0], table.shape[1], L.ptr(ids), B * S, L.ptr(pos_table), S,
                                 L.ptr(out), L.ptr(bad), L.stream_ptr()))
    if int(bad.item()):
        raise IndexError("token id outside the embedding table")
    return out
