"""Op-level wrappers over the C ABI (the `rg_op_*` entry points of include/rg_b200.h).

Used by the per-module mirrors in mogen_api.py (API-faithful path: one call per reference module)
and by the unit parity tests.  Inputs must be CUDA fp32 tensors; nothing here computes in PyTorch.
"""
import torch

from . import _lib

D = 512


def _prep(*ts):
    _lib.require_cuda(*ts)
    return [None if t is None else t.float().contiguous() for t in ts]


def linear(x, weight, bias=None, residual=None, epilogue=_lib.OP_NONE):
    """epilogue(x @ weight.T + bias [+ residual]); x [..., K] -> [..., N]."""
    x, weight, bias, residual = _prep(x, weight, bias, residual)
    N, K = weight.shape
    M = x.numel() // K
    out = torch.empty(*x.shape[:-1], N, device=x.device)
    if residual is not None:
        epilogue = _lib.OP_RESIDUAL
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rg_op_linear(_lib.ptr(x), K, _lib.ptr(weight), _lib.ptr(bias),
                                            _lib.ptr(residual), _lib.ptr(out), M, N, K, epilogue,
                                            _lib.stream_ptr()))
    return out


def linear_tc(x, weight, bias=None, residual=None, epilogue=_lib.OP_NONE, split=False, want_bf16=False):
    """Tensor-core (tcgen05) version of `linear`: bf16 operands (split=False) or bf16x3 (split=True).
    Returns fp32 [.., N] and, if want_bf16, also the bf16 output planes as a bf16 tensor."""
    x, weight, bias, residual = _prep(x, weight, bias, residual)
    N, K = weight.shape
    M = x.numel() // K
    out = torch.empty(*x.shape[:-1], N, device=x.device)
    o16 = torch.empty(M, N * (2 if split else 1), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    if residual is not None:
        epilogue = _lib.OP_RESIDUAL
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rg_op_linear_tc(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(residual),
                                               _lib.ptr(out), _lib.ptr(o16), M, N, K, epilogue, int(split),
                                               _lib.stream_ptr()))
    return (out, o16) if want_bf16 else out


def split_bf16(weight, split):
    """Weight [N, K] -> its bf16 planes ([N, K], or [N, 2K] = hi | lo when split) for linear_tc(w16=...)."""
    (weight,) = _prep(weight)
    N, K = weight.shape
    out = torch.empty(N, K * (2 if split else 1), device=weight.device, dtype=torch.bfloat16)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().rg_op_split_bf16(_lib.ptr(weight), _lib.ptr(out), N, K, int(split), _lib.stream_ptr()))
    return out


def linear_tc_w16(x, w16, n_out, bias=None, epilogue=_lib.OP_NONE, split=False):
    """linear_tc with the weight already converted by split_bf16 (the codec keeps its weights that way)."""
    x, bias = _prep(x, bias)
    K = x.shape[-1]
    M = x.numel() // K
    assert w16.dtype == torch.bfloat16 and tuple(w16.shape) == (n_out, K * (2 if split else 1)) and w16.is_contiguous()
    out = torch.empty(*x.shape[:-1], n_out, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rg_op_linear_tc_w16(_lib.ptr(x), _lib.ptr(w16), _lib.ptr(bias), None, _lib.ptr(out), None,
                                                   M, n_out, K, epilogue, int(split), _lib.stream_ptr()))
    return out


def mha(q, k, v, heads, keep=None):
    """softmax(q k^T / sqrt(dh)) v per head.  q [N, Sq, D], k, v [N, Sk, D] fp32 CUDA tensors whose last dim is
    contiguous (column slices of a fused projection are taken as they are, no copy); keep [N, Sk] bool or None."""
    _lib.require_cuda(q, k, v)
    N, Sq, D = q.shape
    Sk = k.shape[1]

    def rows(t, S):
        if t.dtype != torch.float32 or t.stride(2) != 1 or t.stride(0) != S * t.stride(1) or t.stride(1) % 4 \
                or t.data_ptr() % 16:
            t = t.float().contiguous()
        return t, t.stride(1)
    (q, ldq), (k, ldk), (v, ldv) = rows(q, Sq), rows(k, Sk), rows(v, Sk)
    if keep is not None:
        keep = keep.to(torch.bool).contiguous()
        assert tuple(keep.shape) == (N, Sk)
    out = torch.empty(N, Sq, D, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().rg_op_mha(q.data_ptr(), k.data_ptr(), v.data_ptr(), _lib.ptr(keep), _lib.ptr(out), N, Sq, Sk,
                                         heads, D // heads, ldq, ldk, ldv, _lib.stream_ptr()))
    return out


def set_gemm_kernel(mode=0, min_rows=0, persist_tiles=0, pair128_min_rows=0):
    """rg_set_gemm_kernel: 0 automatic, 1 the 128x128 one-tile-per-CTA kernel, 2 the 2-CTA 256x256 kernel whenever
    the shape allows (N % 256 == 0; persist_tiles: pair-tile count from which its persistent variant runs), 3 the
    pair128 kernel (a CTA pair shares each weight tile).  Thresholds <= 0 are left as they are."""
    _lib.check(_lib.load().rg_set_gemm_kernel(int(mode), int(min_rows), int(persist_tiles), int(pair128_min_rows)))


def layernorm(x, gamma=None, beta=None):
    x, gamma, beta = _prep(x, gamma, beta)
    assert x.shape[-1] == D, "rg_b200 LayerNorm kernels are built for 512-wide rows"
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rg_op_layernorm(_lib.ptr(x), _lib.ptr(gamma), _lib.ptr(beta),
                                               _lib.ptr(out), x.numel() // D, _lib.stream_ptr()))
    return out


def silu(x):
    (x,) = _prep(x)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rg_op_silu(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr()))
    return out


def stylization_rows(y, gamma, beta, ss, rows_per_clip):
    """silu(LN(y) * (1 + scale) + shift); ss [B,1024] (per clip) or [1024] (shared)."""
    y, gamma, beta, ss = _prep(y, gamma, beta, ss)
    out = torch.empty_like(y)
    with torch.cuda.device(y.device):
        _lib.check(_lib.load().rg_op_stylization_rows(
            _lib.ptr(y), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(ss), int(ss.dim() > 1),
            rows_per_clip, _lib.ptr(out), y.numel() // D, _lib.stream_ptr()))
    return out


def self_attention(qkv, src_mask, gamma=None, beta=None, ss=None, x_res=None):
    """qkv [B,T,1536], src_mask [B,T] -> [B,T,512]; stylised rows if ss is given else x_res + Y."""
    qkv, src_mask, gamma, beta, ss, x_res = _prep(qkv, src_mask, gamma, beta, ss, x_res)
    B, T = qkv.shape[0], qkv.shape[1]
    out = torch.empty(B, T, D, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().rg_op_self_attention(
            _lib.ptr(qkv), _lib.ptr(src_mask), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(ss),
            int(ss is not None and ss.dim() > 1), _lib.ptr(x_res), _lib.ptr(out), B, T,
            int(ss is not None), _lib.stream_ptr()))
    return out


def cross_attention(q, state, query_mask, gamma, beta, ss):
    """q [B,T,512] pre-softmax, state [B,16,32,32], query_mask [B,T] or None -> stylised rows."""
    q, state, query_mask, gamma, beta, ss = _prep(q, state, query_mask, gamma, beta, ss)
    B, T = q.shape[0], q.shape[1]
    out = torch.empty(B, T, D, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().rg_op_cross_attention(
            _lib.ptr(q), _lib.ptr(state), _lib.ptr(query_mask), _lib.ptr(gamma), _lib.ptr(beta),
            _lib.ptr(ss), int(ss.dim() > 1), _lib.ptr(out), B, T, _lib.stream_ptr()))
    return out


def self_attention_core(qkv, src_mask, mode=0):
    """qkv [B,T,1536], src_mask [B,T] -> Y [B,T,512] before the Stylization prologue.
    mode 0: fp32 FMA, 1: TF32 mma.sync, 2: 3xTF32 (hi/lo split)."""
    qkv, src_mask = _prep(qkv, src_mask)
    B, T = qkv.shape[0], qkv.shape[1]
    out = torch.empty(B, T, D, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().rg_op_self_attention_core(_lib.ptr(qkv), _lib.ptr(src_mask), _lib.ptr(out), B, T,
                                                         int(mode), _lib.stream_ptr()))
    return out


def cross_attention_core(q3, state, query_mask, mode=0):
    """q3 [B,T,1536] (three conditions), state [B,3,16,32,32], query_mask [3,B,T] or None -> Y [B,T,1536]."""
    q3, state, query_mask = _prep(q3, state, query_mask)
    B, T = q3.shape[0], q3.shape[1]
    out = torch.empty(B, T, 3 * D, device=q3.device)
    with torch.cuda.device(q3.device):
        _lib.check(_lib.load().rg_op_cross_attention_core(_lib.ptr(q3), _lib.ptr(state), _lib.ptr(query_mask),
                                                          _lib.ptr(out), B, T, int(mode), _lib.stream_ptr()))
    return out


def self_attention_tc(qkv, src_mask, gamma, beta, ss, split=False):
    """Fused mma.sync core + Stylization prologue: qkv [B,T,1536] -> stylised rows [B,T,512] (fp32)."""
    qkv, src_mask, gamma, beta, ss = _prep(qkv, src_mask, gamma, beta, ss)
    B, T = qkv.shape[0], qkv.shape[1]
    out = torch.empty(B, T, D, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().rg_op_self_attention_tc(
            _lib.ptr(qkv), _lib.ptr(src_mask), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(ss), int(ss.dim() > 1),
            _lib.ptr(out), B, T, int(split), _lib.stream_ptr()))
    return out


def cross_attention_tc(q3, state, query_mask, gamma3, beta3, ss3, split=False):
    """Fused core + Stylization for the three conditions: q3 [B,T,1536], state [B,3,16,32,32], query_mask
    [3,B,T] or None, gamma3/beta3 [3,512], ss3 [3,1024] or [B,3,1024] -> [B,T,1536] (fp32)."""
    q3, state, query_mask, gamma3, beta3, ss3 = _prep(q3, state, query_mask, gamma3, beta3, ss3)
    B, T = q3.shape[0], q3.shape[1]
    out = torch.empty(B, T, 3 * D, device=q3.device)
    with torch.cuda.device(q3.device):
        _lib.check(_lib.load().rg_op_cross_attention_tc(
            _lib.ptr(q3), _lib.ptr(state), _lib.ptr(query_mask), _lib.ptr(gamma3), _lib.ptr(beta3), _lib.ptr(ss3),
            int(ss3.dim() > 2), _lib.ptr(out), B, T, int(split), _lib.stream_ptr()))
    return out


def kv_state(kv, B, n_tokens):
    """kv [B*N,1024] = [key | value] projections -> state [B,16,32,32]."""
    (kv,) = _prep(kv)
    state = torch.empty(B, 16, 32, 32, device=kv.device)
    with torch.cuda.device(kv.device):
        _lib.check(_lib.load().rg_op_kv_state(_lib.ptr(kv), n_tokens, B, _lib.ptr(state),
                                              _lib.stream_ptr()))
    return state
