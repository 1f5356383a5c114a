"""ctypes binding of librg_b200.so (include/rg_b200.h).  No fallback: if the library cannot be
loaded, or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librg_b200.so")

PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
OP_NONE, OP_RESIDUAL, OP_GELU, OP_SILU = 0, 1, 2, 4


class RgConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "latent_dim", "num_heads", "ffn_dim", "time_embed_dim", "num_layers", "n_tokens",
        "n_chunks", "text_dim", "num_speakers", "precision")]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/rg_b200.h declaration by declaration
SIGNATURES = {
    "rg_last_error": (C.c_char_p, []),
    "rg_abi_version": (_I, []),
    "rg_launch_count": (_L, []),
    "rg_create": (_I, [C.POINTER(RgConfig), _I, C.POINTER(C.c_char_p), C.POINTER(_P), C.POINTER(_L),
                       C.POINTER(_P)]),
    "rg_destroy": (_I, [_P]),
    "rg_set_lanes": (_I, [_P, _I]),
    "rg_set_graphs": (_I, [_P, _I]),
    "rg_set_schedule": (_I, [_P, _I, C.POINTER(C.c_int32), C.POINTER(_F), _P]),
    "rg_encode_conditions": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "rg_state_floats_per_clip": (_L, [_P]),
    "rg_precompute_clip_state": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "rg_denoise": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "rg_denoise_groups": (_I, [_P, _P, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P, _P, _P, _P]),
    "rg_run_levels": (_I, [_P, _I, _P, _I, _I, _P, _P, _P, C.POINTER(C.c_int32), _F, _I, _P, _P, _P, _P, _P, _P]),
    "rg_ddim_update": (_I, [_P, _P, _P, _I, _I, _P, _L, _P]),
    "rg_blend_in_seq": (_I, [_P, _P, _P, _P, _I, _P, _L, _P]),
    "rg_guidance_steps": (_I, [_P, _P, _P, _L, _I, _F, _L, _P]),
    "rg_mix_branches": (_I, [_P, _P, _I, _P, _P, _P, _P]),
    "rg_op_linear": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "rg_op_linear_tc": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rg_op_split_bf16": (_I, [_P, _P, _I, _I, _I, _P]),
    "rg_op_linear_tc_w16": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rg_op_mha": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _L, _P]),
    "rg_probe_gemm_tc": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _L, C.POINTER(C.c_float), _P]),
    "rg_probe_gemm_only": (_I, [_P, _I]),
    "rg_probe_l2_read": (_I, [_P, _L, _I, C.POINTER(C.c_float), _P]),
    "rg_set_gemm_kernel": (_I, [_I, _I, _I, _I]),
    "rg_probe_gemm_trace": (_I, [_I, _I, _I, _I, _I, C.POINTER(_L), _L, _P]),
    "rg_op_layernorm": (_I, [_P, _P, _P, _P, _I, _P]),
    "rg_op_silu": (_I, [_P, _P, _L, _P]),
    "rg_op_stylization_rows": (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _P]),
    "rg_op_self_attention": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _P]),
    "rg_op_cross_attention": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _P]),
    "rg_op_self_attention_core": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "rg_op_cross_attention_core": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "rg_op_self_attention_tc": (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _P]),
    "rg_op_cross_attention_tc": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _P]),
    "rg_op_kv_state": (_I, [_P, _I, _I, _P, _P]),
    "rg_text_similarity": (_I, [_P, _P, _L, _I, _I, _P, _I, _P, _L, _P, _P]),
    "rg_knn_topk": (_I, [_P, _L, _I, _P, _I, _I, _L, _P, _P, _P]),
    "rg_probe_knn_scan": (_I, [_P, _L, _I, _P, _I, _I, _I, _P, _L, C.POINTER(C.c_float), _P]),
    "rg_knn_merge": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "rg_knn_merge_packed": (_I, [_P, _L, _I, _I, _I, _P, _P, _P]),
    "rg_knn_index_create": (_I, [_P, _L, _I, C.POINTER(_P), _P]),
    "rg_knn_index_destroy": (_I, [_P]),
    "rg_knn_topk_tc": (_I, [_P, _P, _P, _I, _I, _L, _P, _P, C.POINTER(C.c_int32), _P]),
    "rg_probe_knn_tc": (_I, [_P, _P, _I, _I, _P, _L, C.POINTER(C.c_float), _P]),
}

_lib = None


def load():
    """Load the shared library (building it is __graft_entry__.build()'s job, not ours)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -m rag_gesture_b200.build` (needs nvcc). "
            "This package has no CPU or PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype, fn.argtypes = res, args
    if lib.rg_abi_version() != 1:
        raise RuntimeError("librg_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("rg_b200: " + load().rg_last_error().decode(errors="replace"))


def ptr(t):
    """Device (or host) address of a contiguous tensor, NULL for None."""
    if t is None:
        return None
    assert t.is_contiguous(), "rg_b200 expects contiguous tensors"
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rg_b200: the hot path runs on CUDA tensors only (no CPU fallback)")
