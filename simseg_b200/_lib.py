"""ctypes binding of libsimseg_b200.so (the C ABI declared in include/simseg_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a Python
exception is raised.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsimseg_b200.so")
# development knob for A/B timing of two builds (tools/): another build of the same library; symbols it lacks stay unbound
_ALT_LIB = os.environ.get("SIMSEG_B200_LIB")
if _ALT_LIB:
    LIB_PATH = _ALT_LIB

F32, BF16 = 0, 1
EPI_NONE, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_ROWSCALE = range(5)
PREC_FP32, PREC_TF32, PREC_SPLIT_BF16 = 0, 1, 2     # PREC_SPLIT_BF16: host-side selector of the fused tcgen05 entry points

vp, i32, i64, f32, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32


class GemmArgs(C.Structure):
    _fields_ = [("a", vp), ("b", vp), ("d", vp), ("M", i64), ("N", i64), ("K", i64),
                ("lda", i64), ("ldb", i64), ("ldd", i64), ("a_major", i32), ("b_major", i32),
                ("in_dtype", i32), ("out_dtype", i32), ("epilogue", i32), ("accumulate", i32),
                ("bias", vp), ("residual", vp), ("ld_res", i64), ("res_dtype", i32),
                ("aux", vp), ("ld_aux", i64), ("row_scale", vp), ("col_sum", vp),
                ("tile_n", i32), ("reserved", i32), ("aux2", vp), ("ld_aux2", i64)]


# name -> (restype, argtypes); every symbol include/simseg_b200.h declares
PROTOTYPES = {
    "simseg_version": (i32, []),
    "simseg_last_error": (C.c_char_p, []),
    "simseg_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "simseg_ctx_destroy": (i32, [vp]),
    "simseg_ctx_launch_count": (i64, [vp, i32]),
    "simseg_gemm": (i32, [vp, C.POINTER(GemmArgs), vp]),
    "simseg_cast_bf16": (i32, [vp, vp, vp, vp, i64, i64, vp]),
    "simseg_cast_bf16_multi": (i32, [vp, vp, i32, i64, vp]),
    "simseg_colsum": (i32, [vp, vp, i32, i64, i64, i64, vp, i32, vp]),
    "simseg_gelu_fwd": (i32, [vp, vp, vp, i64, vp]),
    "simseg_layernorm_fwd": (i32, [vp, vp, i32, vp, vp, f32, i64, i32, vp, vp, vp, vp, vp]),
    "simseg_add_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, f32, i64, i32, vp, vp, vp, vp, vp, vp]),
    "simseg_layernorm_bwd": (i32, [vp, vp, i32, vp, vp, i32, vp, vp, vp, i64, i32, vp, i32, vp, vp, vp, vp, vp]),
    "simseg_attention_fwd": (i32, [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp, f32, vp, vp, vp]),
    "simseg_attention_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp, f32, vp, vp, vp, vp]),
    "simseg_layernorm_fwd_dropout": (i32, [vp, vp, vp, vp, vp, f32, i64, i32, vp, vp, vp, vp, vp, f32, vp, u32, vp]),
    "simseg_layernorm_bwd_dropout": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, i64, i32, vp, i32, vp, vp, vp, vp, i32, f32, vp, u32, vp]),
    "simseg_attn_dropout_mask_words": (i64, [i32, i32, i32]),
    "simseg_attn_dropout_mask": (i32, [vp, i32, i32, i32, f32, vp, u32, vp, i64, vp]),
    "simseg_attention_fwd_dropout": (i32, [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp, f32, vp, vp, vp, f32, vp]),
    "simseg_attention_bwd_dropout": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp, f32, vp, vp, vp, vp, f32, vp]),
    "simseg_im2col16": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "simseg_vit_tokens_fwd": (i32, [vp, vp, i32, vp, vp, i32, i32, i32, vp, vp]),
    "simseg_vit_tokens_bwd": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "simseg_bert_embed_fwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    "simseg_bert_embed_bwd": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "simseg_topk_pool_l2norm_fwd": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, f32, vp, vp, vp, vp]),
    "simseg_topk_pool_l2norm_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, vp, vp]),
    "simseg_proj_topk_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, f32, vp, vp, vp, vp]),
    "simseg_proj_topk_bwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, vp, vp, vp, vp]),
    "simseg_infonce_fwd": (i32, [vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp]),
    "simseg_infonce_bwd": (i32, [vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, f32, vp, vp, vp, vp, vp]),
    "simseg_debug_trace_enable": (i32, [i32]),
    "simseg_debug_trace_read": (i32, [vp, i32]),
    "simseg_infonce_fused_workspace_bytes": (i64, [i32, i32, i32]),
    "simseg_infonce_fused_fwd": (i32, [vp, vp, vp, i32, i32, i32, vp, i32, vp, i64, vp, vp, vp, vp]),
    "simseg_infonce_fused_bwd": (i32, [vp, i32, i32, i32, vp, i32, vp, f32, vp, i64, vp, vp, vp, vp]),
    "simseg_retrieval_fused_workspace_bytes": (i64, [i32, i32, i32]),
    "simseg_retrieval_rank_fused": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, i64, vp, vp]),
    "simseg_allpairs_sim_split": (i32, [vp, vp, vp, i32, i32, i32, vp, i64, vp, vp]),
    "simseg_patch_text_sim_workspace_bytes": (i64, [i64]),
    "simseg_patch_text_sim": (i32, [vp, vp, i32, i64, i32, vp, i32, i32, vp, vp, vp, i64, vp]),
    "simseg_allpairs_sim": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    "simseg_retrieval_rank": (i32, [vp, vp, i32, i32, vp, vp, vp, vp]),
    "simseg_seg_class_embed": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "simseg_seg_select": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "simseg_seg_upsample_norm": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "simseg_pos_embed_bicubic": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
}

_lib = None


class SimsegError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built — no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SimsegError(f"{LIB_PATH} not found: run `python -m simseg_b200.build` "
                              "(or __graft_entry__.build()); there is no fallback path")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            if _ALT_LIB and not hasattr(lib, name):
                continue
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().simseg_last_error().decode("utf-8", "replace")
        raise SimsegError(f"{what} failed (rc={rc}): {msg}")
