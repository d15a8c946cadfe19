"""ctypes binding of ``csrc/librt_reptext.so`` (C-ABI declared in ``include/reptext_rt.h``).

There is NO CPU fallback: if the shared library is missing, or a call fails, this module raises.
Tensors cross the boundary as raw device pointers + sizes; torch is used for device memory only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# RT_LIB: A/B aid - load another BUILD of this same library (e.g. the previous commit's) instead of the in-tree one
LIB_PATH = os.environ.get("RT_LIB") or os.path.join(HERE, "csrc", "librt_reptext.so")

RT_F32, RT_BF16 = 0, 1
RT_TRANSFORMER, RT_CONTROLNET = 0, 1
EPI_BIAS, EPI_GELU, EPI_QKNORM_ROPE, EPI_GATE_RESID, EPI_SCALE_MASK = range(5)
RT_ERR_INVALID = -1

EXPORTS = [
    "rt_last_error", "rt_abi_version", "rt_struct_size", "rt_launch_count", "rt_set_option", "rt_get_option",
    "rt_profile_reset", "rt_profile_read",
    "rt_model_create", "rt_model_set_weight", "rt_model_finalize", "rt_model_destroy", "rt_model_workspace_bytes",
    "rt_controlnet_forward", "rt_transformer_forward", "rt_controlnet_set_live", "rt_model_set_step_invariant_cache",
    "rt_model_modulation_table_bytes", "rt_model_build_modulation_table", "rt_model_select_modulation",
    "rt_controlnet_forward_lockstep", "rt_transformer_forward_lockstep",
    "rt_ipc_alloc", "rt_ipc_open", "rt_ipc_close", "rt_ipc_free", "rt_sp_barrier", "rt_sp_status", "rt_sp_reset",
    "rt_euler_step", "rt_cfg_combine", "rt_cfg_euler_step", "rt_mask_scale_add", "rt_glyph_init_blend",
    "rt_gemm", "rt_attention", "rt_layernorm_modulate", "rt_rope_table", "rt_qknorm_rope",
    "rt_groupnorm_nhwc", "rt_upsample_nearest2x_nhwc", "rt_softmax_rows", "rt_softmax_rows_f32", "rt_im2col3x3_nhwc",
    "rt_nchw_to_nhwc", "rt_nhwc_to_nchw", "rt_vae_posterior_sample",
    "rt_norm_rows", "rt_text_attention", "rt_glu_act", "rt_embedding",
]


class ModelConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("dtype", C.c_int), ("in_channels", C.c_int), ("cond_channels", C.c_int),
        ("out_channels", C.c_int), ("num_layers", C.c_int), ("num_single_layers", C.c_int),
        ("num_attention_heads", C.c_int), ("attention_head_dim", C.c_int), ("joint_attention_dim", C.c_int),
        ("pooled_projection_dim", C.c_int), ("guidance_embeds", C.c_int), ("axes_dims_rope", C.c_int * 3),
    ]


SP_MAX_RANKS = 8
SP_FLAG_WORDS = 16


class SpGroup(C.Structure):
    """``rt_sp_group``: one sample's tokens sharded over ``world`` GPUs (sequence-parallel attention)."""
    _fields_ = [
        ("world", C.c_int), ("rank", C.c_int), ("peer_workspace", C.c_void_p * SP_MAX_RANKS),
        ("peer_flags", C.c_void_p * SP_MAX_RANKS), ("lockstep", C.c_int),
    ]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("lat_batch", C.c_int), ("t_batch", C.c_int), ("n_img", C.c_int), ("n_txt", C.c_int),
        ("hidden_states", C.c_void_p), ("encoder_hidden_states", C.c_void_p), ("pooled_projections", C.c_void_p),
        ("timestep", C.c_void_p), ("guidance", C.c_void_p), ("img_ids", C.c_void_p), ("txt_ids", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
        ("sp", C.POINTER(SpGroup)),
    ]


class ControlNetCall(C.Structure):
    _fields_ = [
        ("a", ForwardArgs), ("controlnet_cond", C.c_void_p), ("cond_batch", C.c_int),
        ("conditioning_scale", C.c_float), ("mask", C.c_void_p), ("accumulate", C.c_int),
        ("block_samples", C.c_void_p), ("single_block_samples", C.c_void_p),
    ]


class TransformerCall(C.Structure):
    _fields_ = [
        ("a", ForwardArgs), ("controlnet_block_samples", C.POINTER(C.c_void_p)), ("n_block_samples", C.c_int),
        ("controlnet_single_block_samples", C.POINTER(C.c_void_p)), ("n_single_block_samples", C.c_int),
        ("out", C.c_void_p),
    ]


class GemmSegment(C.Structure):
    _fields_ = [
        ("W", C.c_void_p), ("bias", C.c_void_p), ("n_begin", C.c_int), ("n_end", C.c_int), ("mode", C.c_int),
        ("out", C.c_void_p), ("out_batch_stride", C.c_int64), ("out_ld", C.c_int), ("out_col0", C.c_int),
        ("norm_w", C.c_void_p), ("scatter", C.c_int), ("out_f32", C.c_int),
    ]


class GemmProblem(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_batch_stride", C.c_int64), ("a_ld", C.c_int), ("a_row0", C.c_int),
        ("a_rows_total", C.c_int), ("m_rows", C.c_int), ("out_row0", C.c_int), ("K", C.c_int), ("nseg", C.c_int),
        ("seg", GemmSegment * 4), ("gate", C.c_void_p), ("gate_ld", C.c_int), ("extra", C.c_void_p),
        ("extra_batch_stride", C.c_int64), ("extra_ld", C.c_int), ("extra_row0", C.c_int), ("scale", C.c_float),
        ("mask", C.c_void_p), ("accumulate", C.c_int), ("conv_h", C.c_int), ("conv_w", C.c_int), ("conv_c", C.c_int),
    ]


class GemmLaunch(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("batch", C.c_int), ("nprob", C.c_int), ("prob", GemmProblem * 2),
        ("rope", C.c_void_p), ("head_dim", C.c_int), ("sp_cols", C.c_int), ("sp_row0", C.c_int),
        ("sp_out", C.c_void_p * SP_MAX_RANKS),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("qkv", C.c_void_p), ("batch_stride", C.c_int64), ("ld", C.c_int), ("q_col0", C.c_int),
        ("k_col0", C.c_int), ("v_col0", C.c_int), ("out", C.c_void_p), ("out_batch_stride", C.c_int64),
        ("out_ld", C.c_int), ("out_col0", C.c_int), ("batch", C.c_int), ("S", C.c_int), ("heads", C.c_int),
        ("hd", C.c_int), ("sp_rows", C.c_int), ("sp_txt_rows", C.c_int), ("sp_out", C.c_void_p * SP_MAX_RANKS),
    ]


class LnModGroup(C.Structure):
    _fields_ = [("row_begin", C.c_int), ("row_end", C.c_int), ("shift", C.c_void_p), ("scale", C.c_void_p),
                ("ld", C.c_int)]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the CUDA runtime library.  Raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m reptext_b200.build` "
            "(reptext_b200 has no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    L.rt_last_error.restype = C.c_char_p
    L.rt_launch_count.restype = C.c_longlong
    L.rt_model_workspace_bytes.restype = C.c_int64
    L.rt_model_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.rt_set_option.argtypes = [C.c_char_p, C.c_int]
    L.rt_get_option.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.rt_profile_read.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.rt_model_create.argtypes = [C.POINTER(ModelConfig), C.POINTER(C.c_void_p)]
    L.rt_model_set_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    L.rt_model_finalize.argtypes = [C.c_void_p, C.c_void_p]
    L.rt_model_destroy.argtypes = [C.c_void_p]
    L.rt_controlnet_forward.argtypes = [C.c_void_p, C.POINTER(ForwardArgs), C.c_void_p, C.c_int, C.c_float,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.rt_transformer_forward.argtypes = [C.c_void_p, C.POINTER(ForwardArgs), C.POINTER(C.c_void_p), C.c_int,
                                         C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    L.rt_controlnet_set_live.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.rt_model_set_step_invariant_cache.argtypes = [C.c_void_p, C.c_int]
    L.rt_model_modulation_table_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.rt_model_modulation_table_bytes.restype = C.c_int64
    L.rt_model_build_modulation_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                  C.c_int64, C.c_void_p]
    L.rt_model_select_modulation.argtypes = [C.c_void_p, C.c_int]
    L.rt_controlnet_forward_lockstep.argtypes = [C.c_void_p, C.c_int, C.POINTER(ControlNetCall)]
    L.rt_transformer_forward_lockstep.argtypes = [C.c_void_p, C.c_int, C.POINTER(TransformerCall)]
    L.rt_ipc_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]
    L.rt_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.rt_ipc_close.argtypes = [C.c_void_p]
    L.rt_ipc_free.argtypes = [C.c_void_p]
    L.rt_sp_barrier.argtypes = [C.POINTER(SpGroup), C.c_void_p]
    L.rt_sp_status.argtypes = [C.POINTER(SpGroup), C.c_void_p, C.POINTER(C.c_int)]
    L.rt_euler_step.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                C.c_void_p]
    L.rt_cfg_combine.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_void_p]
    L.rt_cfg_euler_step.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int,
                                    C.c_float, C.c_float, C.c_void_p]
    L.rt_mask_scale_add.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_float, C.c_void_p]
    L.rt_glyph_init_blend.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_float, C.c_float, C.c_void_p]
    L.rt_gemm.argtypes = [C.POINTER(GemmLaunch), C.c_int, C.c_void_p]
    L.rt_attention.argtypes = [C.POINTER(AttentionArgs), C.c_int, C.c_void_p]
    L.rt_layernorm_modulate.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.POINTER(LnModGroup), C.c_void_p]
    L.rt_groupnorm_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    L.rt_upsample_nearest2x_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.rt_softmax_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p]
    L.rt_softmax_rows_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    L.rt_im2col3x3_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.rt_nchw_to_nhwc.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]
    L.rt_nhwc_to_nchw.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p]
    L.rt_vae_posterior_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_void_p]
    L.rt_norm_rows.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_float, C.c_int, C.c_void_p]
    L.rt_text_attention.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int,
                                    C.c_void_p]
    L.rt_glu_act.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
    L.rt_embedding.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p]
    L.rt_rope_table.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    L.rt_qknorm_rope.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    for i, st in enumerate((ModelConfig, ForwardArgs, SpGroup, ControlNetCall, TransformerCall, GemmSegment,
                            GemmProblem, GemmLaunch, AttentionArgs, LnModGroup)):
        if L.rt_struct_size(i) != C.sizeof(st):
            raise RuntimeError(f"ctypes mirror of {st.__name__} is {C.sizeof(st)} bytes, the library says "
                               f"{L.rt_struct_size(i)}: _lib.py and include/reptext_rt.h are out of step")
    _lib = L
    # A/B aid: RT_OPTIONS="attn_variant=20,gemm_cta_group=1" applies rt_set_option at load time
    for kv in filter(None, os.environ.get("RT_OPTIONS", "").split(",")):
        k, v = kv.split("=")
        check(L.rt_set_option(k.strip().encode(), int(v)))
    return L


def check(status: int) -> None:
    """Translate a status code into the exception the reference would raise (ValueError for bad
    arguments, e.g. RepText/controlnet_flux.py:297; RuntimeError otherwise)."""
    if status == 0:
        return
    msg = lib().rt_last_error().decode("utf-8", "replace")
    if status == RT_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(f"librt_reptext error {status}: {msg}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return RT_BF16
    if dt == torch.float32:
        return RT_F32
    raise ValueError(f"unsupported dtype {dt}: the runtime computes in float32 or bfloat16")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("reptext_b200 needs CUDA tensors (there is no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # the library launches on the CURRENT device's stream; a tensor elsewhere would be read from the wrong GPU
        raise ValueError(f"tensor on {t.device} but the current device is cuda:{torch.cuda.current_device()}: "
                         f"enter torch.cuda.device({t.device.index}) around the call")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().rt_launch_count())


def set_option(name: str, value: int) -> None:
    check(lib().rt_set_option(name.encode(), int(value)))


PROF_CLASSES = ["gemm_tcgen05", "gemm_simt", "attention_tcgen05", "attention_simt", "layernorm_modulate", "gemv",
                "elementwise", "attention_text_mma"]


def profile_reset() -> None:
    check(lib().rt_profile_reset())


def profile_read() -> dict:
    """{class name: (total ms, total algorithmic work, launches)} since the last reset (option "profile")."""
    out = {}
    for i, name in enumerate(PROF_CLASSES):
        ms, work, n = C.c_double(), C.c_double(), C.c_longlong()
        check(lib().rt_profile_read(i, C.byref(ms), C.byref(work), C.byref(n)))
        if n.value:
            out[name] = (ms.value, work.value, n.value)
    return out
