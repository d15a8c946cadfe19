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
