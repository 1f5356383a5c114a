"""ORACLE (test infrastructure, not product): CPU restatement of the reference denoiser.

Plain torch fp32 on the CPU, op for op in the order the reference executes them -- including the
work the product hoists (cross-attention K/V projections and the StylizationBlock `emb` GEMVs are
recomputed on every call, exactly as efficient_attention.py / stylization_block.py do), so that
timing this file is timing the reference's algorithm.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

Parity pin: tests/golden/make_golden.py runs the UNMODIFIED reference modules (imported from
/root/reference through tests/golden/refshim.py) on the same seeded inputs and commits their
outputs under tests/golden/; tests/test_oracle_golden.py holds this file to those vectors.

`sd` is a plain dict of tensors keyed like the reference state dict of ReGestureTransformer
(SURVEY 8b), without the leading "model.".
"""
import math

import torch
import torch.nn.functional as F

CONDS = ("xf_text", "xf_audio", "xf_spk")


def timestep_embedding(timesteps, dim, max_period=10000):
    """diffusion_transformer.py:27-46."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def stylization_block(sd, p, h, emb):
    """stylization_block.py:29-40 (dropout p=0)."""
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb)).unsqueeze(1)
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = _ln(sd, p + ".norm", h) * (1 + scale) + shift
    return _lin(sd, p + ".out_layers.2", F.silu(h))


def efficient_self_attention(sd, p, x, src_mask, emb, H):
    """efficient_attention.py:23-45.  src_mask [B,T,1]."""
    B, T, D = x.shape
    query = _lin(sd, p + ".query", _ln(sd, p + ".norm", x))
    key = _lin(sd, p + ".key", _ln(sd, p + ".norm", x)) + (1 - src_mask) * -1000000
    query = F.softmax(query.view(B, T, H, -1), dim=-1)
    key = F.softmax(key.view(B, T, H, -1), dim=1)
    value = (_lin(sd, p + ".value", _ln(sd, p + ".norm", x)) * src_mask).view(B, T, H, -1)
    attention = torch.einsum("bnhd,bnhl->bhdl", key, value)
    y = torch.einsum("bnhd,bhdl->bnhl", query, attention).reshape(B, T, D)
    return x + stylization_block(sd, p + ".proj_out", y, emb)


def efficient_cross_attention(sd, p, x, xf, emb, query_mask, cond_type, H, taps=None):
    """efficient_attention.py:62-102.  cond_type [B,1,1] or None; query_mask [B,T] or None."""
    B, T, D = x.shape
    N = xf.shape[1]
    query = _lin(sd, p + ".query", _ln(sd, p + ".norm", x))
    key = _lin(sd, p + ".key", _ln(sd, p + ".text_norm", xf))
    query = F.softmax(query.view(B, T, H, -1), dim=-1)
    if cond_type is None:
        key = F.softmax(key.view(B, N, H, -1), dim=1)
        value = _lin(sd, p + ".value", _ln(sd, p + ".text_norm", xf)).view(B, N, H, -1)
    else:
        c = ((cond_type % 10) > 0).float().view(B, 1, 1).repeat(1, N, 1)
        key = key + (1 - c) * -1000000
        key = F.softmax(key.view(B, N, H, -1), dim=1)
        value = _lin(sd, p + ".value", _ln(sd, p + ".text_norm", xf) * c).view(B, N, H, -1)
    attention = torch.einsum("bnhd,bnhl->bhdl", key, value)
    y = torch.einsum("bnhd,bhdl->bnhl", query, attention)
    if taps is not None:
        taps.setdefault("ca_y_absmax", []).append(float(y.abs().max()))
        taps.setdefault("ca_state", []).append(attention)
    if query_mask is not None:
        y = y + (1 - query_mask).view(B, T, 1, 1) * -1000000
    y = y.reshape(B, T, D)
    return x + stylization_block(sd, p + ".proj_out", y, emb)


def ffn(sd, p, x, emb):
    """diffusion_transformer.py:84-87 (exact-erf GELU, dropout p=0)."""
    y = _lin(sd, p + ".linear2", F.gelu(_lin(sd, p + ".linear1", x)))
    return x + stylization_block(sd, p + ".proj_out", y, emb)


def decoder_layer(sd, p, x, xf, emb, src_mask, query_mask, cond_type, H, taps=None):
    """diffusion_transformer.py:105-127: SA, three CAs on the same x, ca_mix, FFN."""
    x = efficient_self_attention(sd, p + ".sa_block", x, src_mask, emb, H)
    outs = []
    for cond, xf_cond in xf.items():
        qm = query_mask[cond] if query_mask is not None else None
        outs.append(efficient_cross_attention(sd, f"{p}.ca_blocks.{cond}", x, xf_cond, emb, qm,
                                              cond_type, H, taps))
    x = _lin(sd, p + ".ca_mix", torch.cat(outs, dim=-1))
    return ffn(sd, p + ".ffn", x, emb)


def encode_conditions(sd, word, audio, speaker_ids):
    """raggesture.py:978-987 -> diffusion_transformer.py:544-606 with the shipped encoders
    (pretrained_model=None, num_layers=0): two Linear(768->512) and an Embedding lookup."""
    return {
        "xf_text": _lin(sd, "text_pre_proj", word),
        "xf_audio": _lin(sd, "audio_pre_proj", audio),
        "xf_spk": F.embedding(speaker_ids, sd["speaker_embedding.weight"]),
    }


def embed_latents(sd, motion):
    """diffusion_transformer.py:646-659: joint_embed + per-part sine pos (0 on separators) +
    learned global pos."""
    T = motion.shape[1]
    h = _lin(sd, "joint_embed", motion)
    n = (T - 3) // 4
    pos = sd["sequence_embedding.pe"].permute(1, 0, 2)[:, :n, :]
    sep = torch.zeros_like(pos[:, :1, :])
    h = h + torch.cat([pos, sep, pos, sep, pos, sep, pos], dim=1)
    return h + sd["global_positional_embedding.pe"].permute(1, 0, 2)[:, :T, :]


def denoiser_forward(sd, motion, timesteps, motion_mask, xf_out, query_mask, num_heads=16,
                     num_layers=8, taps=None):
    """DiffusionTransformer.forward (diffusion_transformer.py:620-668) + ReGestureTransformer
    .forward_test single-branch (raggesture.py:1041-1086, scale_func_cfg=None, no clf guidance).

    motion [B,T,D] fp32, timesteps [B] int64 on the ORIGINAL 0..999 scale, motion_mask [B,T],
    xf_out dict of 3 x [B,N,D], query_mask dict of 3 x [B,T] -> x0 prediction [B,T,D]."""
    B, T, D = motion.shape
    src_mask = motion_mask.clone().unsqueeze(-1)
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", timestep_embedding(timesteps, D))))
    h = embed_latents(sd, motion)
    cond_type = torch.zeros(B, 1, 1) + 1
    for l in range(num_layers):
        h = decoder_layer(sd, f"temporal_decoder_blocks.{l}", h, xf_out, emb, src_mask, query_mask,
                          cond_type, num_heads, taps)
        if taps is not None:
            taps.setdefault("h", []).append(h)
    return _lin(sd, "out", h)


def scale_func_retr(scale_func_cfg, timestep, rng):
    """raggesture.py:925-954: the four mixing coefficients; above t = 100 one of two sets is drawn with
    rng.randint(0, 1) (the reference uses Python's global `random`)."""
    w = (1 - (1000 - timestep) / 1000) * scale_func_cfg["coarse_scale"] + 1
    if timestep > 100:
        if rng.randint(0, 1) == 0:
            return dict(both_coef=w, text_coef=0, retr_coef=1 - w, none_coef=0)
        return dict(both_coef=0, text_coef=w, retr_coef=0, none_coef=1 - w)
    both, text, retr = scale_func_cfg["both_coef"], scale_func_cfg["text_coef"], scale_func_cfg["retr_coef"]
    return dict(both_coef=both, text_coef=text, retr_coef=retr, none_coef=1 - both - text - retr)


def joint_scale_mask(per_joint_scale, T=43):
    """raggesture.py:910-921: one scale per token row (separator rows keep 1)."""
    n = (T - 3) // 4
    m = torch.ones(T)
    m[0:n] = per_joint_scale["upper"]
    m[n + 1:2 * n + 1] = per_joint_scale["hands"]
    m[2 * n + 2:3 * n + 2] = per_joint_scale["face"]
    m[3 * n + 3:T] = per_joint_scale["lowertransl"]
    return m


def denoiser_forward_two_branch(sd, motion, timesteps, motion_mask, xf_out, query_mask, scale_func_cfg,
                                per_joint_scale, rng, num_heads=16, num_layers=8):
    """forward_test with scale_func_cfg set (raggesture.py:1041-1111): the batch is evaluated twice -- cond_type 1
    (text branch) and cond_type 0 ("none" branch: keys - 1e6, values of the zeroed condition) -- and mixed with
    scale_func_retr's coefficients and the per-row joint scale, in the reference's operation order."""
    B, T, D = motion.shape
    src_mask = motion_mask.clone().unsqueeze(-1).repeat(2, 1, 1)
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", timestep_embedding(timesteps, D)))).repeat(2, 1)
    h = embed_latents(sd, motion).repeat(2, 1, 1)
    xf = {k: v.repeat(2, 1, 1) for k, v in xf_out.items()}
    qm = {k: v.repeat(2, 1) for k, v in query_mask.items()} if query_mask is not None else None
    cond_type = torch.cat([torch.zeros(B, 1, 1) + 1, torch.zeros(B, 1, 1)], 0)
    for l in range(num_layers):
        h = decoder_layer(sd, f"temporal_decoder_blocks.{l}", h, xf, emb, src_mask, qm, cond_type, num_heads)
    out = _lin(sd, "out", h)
    c = scale_func_retr(scale_func_cfg, int(timesteps[0]), rng)
    out_text, out_none = out[:B].contiguous(), out[B:].contiguous()
    js = joint_scale_mask(per_joint_scale, T).unsqueeze(0).unsqueeze(-1).expand(B, -1, D)
    return (out_text * c["both_coef"] * js + out_text * c["text_coef"] * js
            + out_none * c["retr_coef"] * (1 / js) + out_none * c["none_coef"] * (1 / js))


class OracleDenoiser:
    """Callable with the reference denoiser's call signature (gaussian_diffusion.py:529-534), so
    oracle/diffusion.py drives it the way SpacedDiffusion drives ReGestureTransformer."""

    def __init__(self, sd, num_heads=16, num_layers=8):
        self.sd, self.num_heads, self.num_layers = sd, num_heads, num_layers

    def get_precompute_condition(self, text=None, audio=None, speaker_ids=None, xf_out=None, **kw):
        if xf_out is None:
            xf_out = encode_conditions(self.sd, text, audio, speaker_ids)
        return {"xf_out": xf_out, "re_dict": kw.get("re_dict")}

    def __call__(self, x, timesteps, motion_mask=None, xf_out=None, query_mask=None, **kw):
        return denoiser_forward(self.sd, x, timesteps, motion_mask, xf_out, query_mask,
                                self.num_heads, self.num_layers)
