"""The latent codec of the path: four per-body-part TransformerVAEs behind
`GestureRepEncoder.encode / decode` (mogen/models/transformers/diffusion_transformer.py:130-330,
gesture_vae.py:24-239, the DETR-style blocks of mogen/models/utils/detr_utils.py:27-210,335-479).

Frozen, inference only.  Module and parameter names follow the reference so its VAE checkpoints load
(`test_vae_state_dict_keys_match_reference`); the computation is organised batch-first: every 15-frame chunk of
every clip (encode) / every clip (decode) is one row of a single fused-attention call
(`scaled_dot_product_attention`), with no per-exemplar loop.  This is host-framework code (PyTorch on the
caller's device), not part of the C ABI: the codec runs twice per batch and once per exemplar batch, outside
the 50-level loops.

Behaviour kept from the reference, on purpose:
  * `encode` samples (rsample) even in eval mode (gesture_vae.py:173-193): one standard-normal draw of shape
    [B*n_chunks, 1, D] per body part, in the order upper, hands, face, lower+trans, on the input's device;
  * in the "all_encoder" decoder the positional term handed to every layer is `xseq + pe`, not `pe`
    (gesture_vae.py:213-216 calls the embedding module on the sequence itself);
  * `encode` rebases `motion_transl[..., 0]` and `[..., 2]` to the first frame IN PLACE (diffusion_transformer.py:231-232).
"""
import math
import os
from argparse import Namespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .longform import axis_angle_to_matrix, matrix_to_axis_angle, matrix_to_rotation_6d, rotation_6d_to_matrix


def _activation(name):
    if name == "relu":
        return F.relu
    if name == "gelu":
        return F.gelu
    raise ValueError(f"transformer_activation must be relu or gelu, not {name!r}")


def _lin(owner, x, W, b, act=None):
    """y = act(x W^T + b).  On the device, with a GEMM tier set on `owner` (GestureRepEncoder.set_gemm_tier) and a
    shape the tcgen05 kernel takes (N % 128 == 0, K % 64 == 0: every projection of the blocks; not the feature
    embedding / final layer, whose nfeats side is ~100 wide), the library's tensor-core GEMM with its fused
    bias / GELU epilogue: "bf16x3" (hi|lo operand split, three passes, fp32-class results) or "bf16".  Otherwise
    cuBLAS fp32 through F.linear, which on this part runs on the FP32 pipes at a tenth of the rate."""
    tier = getattr(owner, "_rg_tier", None)
    if tier is not None and x.is_cuda and x.dtype == torch.float32 and W.shape[0] % 128 == 0 and W.shape[1] % 64 == 0:
        from . import _lib, ops
        split = tier == "bf16x3"
        # the weight's bf16 planes are converted once and kept on the module that owns the projection, keyed by the
        # row block's address and checked against the parameter's in-place version counter
        cache = owner.__dict__.setdefault("_rg_w16", {})
        key, stamp = (W.data_ptr(), W.shape[0], split), W._version
        hit = cache.get(key)
        if hit is None or hit[0] != stamp:
            if len(cache) >= 16:                           # parameters moved (.to(), new storage): drop the old planes
                cache.clear()
            hit = cache[key] = (stamp, ops.split_bf16(W, split))
            torch.cuda.current_stream(x.device).synchronize()      # once per weight: other streams (the pipeline's
                                                                   # worker / main thread) read the planes later
        y = ops.linear_tc_w16(x, hit[1], W.shape[0], b, epilogue=_lib.OP_GELU if act is F.gelu else _lib.OP_NONE, split=split)
        return y if act is None or act is F.gelu else act(y)
    y = F.linear(x, W, b)
    return y if act is None else act(y)


class _LearnedPositions(nn.Module):
    """`pe` [max_len, 1, D] as in detr_utils.py:60-79; applied batch-first here."""

    def __init__(self, dim, max_len=1024):
        super().__init__()
        self.pe = nn.Parameter(torch.empty(max_len, 1, dim))
        nn.init.xavier_uniform_(self.pe)

    def forward(self, x):                      # x [N, S, D]
        return x + self.pe[:x.shape[1], 0]


class _SinePositions(nn.Module):
    """Fixed sin/cos table under the same buffer name (detr_utils.py:27-57)."""

    def __init__(self, dim, max_len=1024):
        super().__init__()
        pos = torch.arange(max_len, dtype=torch.float32)[:, None]
        freq = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
        pe = torch.zeros(max_len, 1, dim)
        pe[:, 0, 0::2] = torch.sin(pos * freq)
        pe[:, 0, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", pe)

    def forward(self, x):
        return x + self.pe[:x.shape[1], 0]


def _mha(container, q_in, k_in, v_in, keep):
    """Multi-head attention with the packed parameters of an nn.MultiheadAttention `container`
    (in_proj_weight [3D, D], in_proj_bias, out_proj).  q_in [N, Sq, D]; k_in, v_in [N, Sk, D];
    keep [N, Sk] bool (True = attend) or None."""
    D, H = container.embed_dim, container.num_heads
    W, b = container.in_proj_weight, container.in_proj_bias
    N, Sq, Sk = q_in.shape[0], q_in.shape[1], k_in.shape[1]
    # projections that share their input are one GEMM over the stacked weight rows
    if q_in is k_in and k_in is v_in:
        q, k, v = _lin(container, q_in, W, b).split(D, dim=-1)
    elif q_in is k_in:
        q, k = _lin(container, q_in, W[:2 * D], b[:2 * D]).split(D, dim=-1)
        v = _lin(container, v_in, W[2 * D:], b[2 * D:])
    elif k_in is v_in:
        q = _lin(container, q_in, W[:D], b[:D])
        k, v = _lin(container, k_in, W[D:], b[D:]).split(D, dim=-1)
    else:
        q, k, v = (_lin(container, t, W[i * D:(i + 1) * D], b[i * D:(i + 1) * D]) for i, t in enumerate((q_in, k_in, v_in)))
    if getattr(container, "_rg_tier", None) is not None and q.is_cuda and D // H in (16, 32, 64, 128):
        from . import ops
        o = ops.mha(q, k, v, H, keep)                      # strided column blocks of the fused projection, no copies
    else:
        q = q.view(N, Sq, H, D // H).transpose(1, 2)
        k = k.view(N, Sk, H, D // H).transpose(1, 2)
        v = v.view(N, Sk, H, D // H).transpose(1, 2)
        mask = None if keep is None else keep[:, None, None, :]
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2).reshape(N, Sq, D)
    return _lin(container, o, container.out_proj.weight, container.out_proj.bias)


class _SelfBlock(nn.Module):
    """detr_utils.py:335-393 (post-norm and pre-norm forms); dropout is identity at inference."""

    def __init__(self, dim, heads, ff, act, pre_norm):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(dim, heads)       # parameter container only
        self.linear1, self.linear2 = nn.Linear(dim, ff), nn.Linear(ff, dim)
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.act, self.pre_norm = _activation(act), pre_norm

    def forward(self, x, keep=None, pos=None, **_):
        if self.pre_norm:
            y = self.norm1(x)
            qk = y if pos is None else y + pos
            x = x + _mha(self.self_attn, qk, qk, y, keep)
            return x + self._ffn(self.norm2(x))
        qk = x if pos is None else x + pos
        x = self.norm1(x + _mha(self.self_attn, qk, qk, x, keep))
        return self.norm2(x + self._ffn(x))

    def _ffn(self, x):
        return _lin(self, _lin(self, x, self.linear1.weight, self.linear1.bias, self.act), self.linear2.weight,
                    self.linear2.bias)


class _CrossBlock(nn.Module):
    """detr_utils.py:396-479: self-attention over the queries, attention into the memory, FFN."""

    def __init__(self, dim, heads, ff, act, pre_norm):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(dim, heads)
        self.multihead_attn = nn.MultiheadAttention(dim, heads)
        self.linear1, self.linear2 = nn.Linear(dim, ff), nn.Linear(ff, dim)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.act, self.pre_norm = _activation(act), pre_norm

    def forward(self, x, keep=None, pos=None, memory=None, **_):
        if self.pre_norm:
            y = self.norm1(x)
            x = x + _mha(self.self_attn, y, y, y, keep)
            x = x + _mha(self.multihead_attn, self.norm2(x), memory, memory, None)
            return x + self._ffn(self.norm3(x))
        x = self.norm1(x + _mha(self.self_attn, x, x, x, keep))
        x = self.norm2(x + _mha(self.multihead_attn, x, memory, memory, None))
        return self.norm3(x + self._ffn(x))

    _ffn = _SelfBlock._ffn


class _SkipStack(nn.Module):
    """U-shaped stack (detr_utils.py:101-209): n input blocks, a middle block, n output blocks, each output
    block fed Linear(concat(x, matching input block's output))."""

    def __init__(self, make_block, num_layers, dim):
        super().__init__()
        n = (num_layers + (num_layers % 2 == 0) - 1) // 2
        self.input_blocks = nn.ModuleList(make_block() for _ in range(n))
        self.middle_block = make_block()
        self.output_blocks = nn.ModuleList(make_block() for _ in range(n))
        self.linear_blocks = nn.ModuleList(nn.Linear(2 * dim, dim) for _ in range(n))
        self.norm = nn.LayerNorm(dim)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, x, **kw):
        kept = []
        for blk in self.input_blocks:
            x = blk(x, **kw)
            kept.append(x)
        x = self.middle_block(x, **kw)
        for blk, lin in zip(self.output_blocks, self.linear_blocks):
            x = blk(_lin(self, torch.cat([x, kept.pop()], dim=-1), lin.weight, lin.bias), **kw)
        return self.norm(x)


def _frame_mask(lengths, n_frames, batch, device):
    if lengths is None:
        return torch.ones(batch, n_frames, dtype=torch.bool, device=device)
    lengths = torch.as_tensor(lengths, device=device)
    return torch.arange(n_frames, device=device)[None, :] < lengths[:, None]


class TransformerVAE(nn.Module):
    """gesture_vae.py:24-239.  `args`: Namespace / dict with latent_dim, frame_chunk_size, decoder_arch
    ("all_encoder" | "encoder_decoder"), position_embedding ("learned" | "sine"), num_frames, num_heads,
    ff_size, dropout, transformer_activation, transformer_normalize_before, num_layers, nfeats, vae_dist
    ("normal" | "multivariate_normal")."""

    def __init__(self, args, **_):
        super().__init__()
        a = args if isinstance(args, Namespace) else Namespace(**dict(args))
        D = self.latent_dim = a.latent_dim
        self.frame_chunk_size, self.arch, self.num_frames = a.frame_chunk_size, a.decoder_arch, a.num_frames
        self.dist_type = a.vae_dist
        if self.dist_type not in ("normal", "multivariate_normal"):
            raise ValueError("Not support distribution type!")
        pos = {"learned": _LearnedPositions, "sine": _SinePositions}[a.position_embedding]
        self.query_pos_encoder, self.query_pos_decoder, self.mem_pos_decoder = pos(D), pos(D), pos(D)
        act, pre = a.transformer_activation, a.transformer_normalize_before
        self.encoder = _SkipStack(lambda: _SelfBlock(D, a.num_heads, a.ff_size, act, pre), a.num_layers, D)
        if self.arch == "all_encoder":
            self.decoder = _SkipStack(lambda: _SelfBlock(D, a.num_heads * 8, a.ff_size, act, pre), a.num_layers, D)
        elif self.arch == "encoder_decoder":
            self.decoder = _SkipStack(lambda: _CrossBlock(D, a.num_heads * 4, a.ff_size, act, pre),
                                      (a.num_layers - 1) * 4 + 1, D)
        else:
            raise ValueError("Not support architecture!")
        self.global_motion_token = nn.Parameter(torch.randn(2, D))          # mu and logvar tokens
        self.skel_embedding = nn.Linear(a.nfeats, D)
        self.final_layer = nn.Linear(D, a.nfeats)

    # -- encoder: every chunk of every clip is one sequence of 2 + frame_chunk_size tokens --------------------
    def encode(self, features, lengths=None):
        """[B, F, nfeats] -> [B*n_chunks, 2, D] (mu token, logvar token)."""
        B, n_frames, _ = features.shape
        c = self.frame_chunk_size
        n_chunks = n_frames // c
        keep = _frame_mask(lengths, n_frames, B, features.device).reshape(B * n_chunks, n_frames // n_chunks)
        x = self.skel_embedding(features.reshape(B * n_chunks, n_frames // n_chunks, -1))
        tok = self.global_motion_token[None].expand(x.shape[0], -1, -1)
        seq = self.query_pos_encoder(torch.cat([tok, x], dim=1))
        keep = torch.cat([torch.ones(x.shape[0], 2, dtype=torch.bool, device=x.device), keep], dim=1)
        return self.encoder(seq, keep=keep)[:, :2]

    def reparameterize(self, latent, eps=None):
        mu, raw = latent[:, 0:1], latent[:, 1:]
        scale = raw.exp().pow(0.5) if self.dist_type == "normal" else F.softplus(raw) + 1e-8
        if eps is None:                         # what Normal / MultivariateNormal .rsample() draw
            eps = torch.empty(mu.shape, dtype=mu.dtype, device=mu.device).normal_(generator=getattr(self, "generator", None))
        return mu + scale * eps, Namespace(loc=mu, scale=scale)

    def encode_to_dist(self, features, lengths=None, eps=None):
        B, n_frames, _ = features.shape
        z, dist = self.reparameterize(self.encode(features, lengths), eps)
        return z.reshape(B, n_frames // self.frame_chunk_size, self.latent_dim), dist

    # -- decoder: one sequence per clip ----------------------------------------------------------------------------
    def decode(self, z, lengths=None):
        """[B, n_chunks, D] -> [B, num_frames (or max(lengths)), nfeats], zero beyond each length."""
        B, n_chunks, D = z.shape
        n_frames = self.num_frames if lengths is None else int(max(lengths))
        keep = _frame_mask(lengths, n_frames, B, z.device)
        queries = torch.zeros(B, n_frames, D, dtype=z.dtype, device=z.device)
        if self.arch == "all_encoder":
            seq = torch.cat([z, queries], dim=1)
            keep_all = torch.cat([torch.ones(B, n_chunks, dtype=torch.bool, device=z.device), keep], dim=1)
            out = self.decoder(seq, keep=keep_all, pos=self.query_pos_decoder(seq))[:, n_chunks:]
        else:
            out = self.decoder(self.query_pos_decoder(queries), keep=keep, memory=self.mem_pos_decoder(z))
        out = self.final_layer(out)
        return out * keep[..., None].to(out.dtype)

    def forward(self, features, lengths=None):
        z, dist = self.encode_to_dist(features, lengths)
        return {"rec_pose": self.decode(z, lengths), "poses_feat": z, "rec_dist": dist}


def _to_6d(aa):
    """[B, F, J*3] axis-angle -> [B, F, J*6] (first two rows of the rotation matrix)."""
    B, n, j3 = aa.shape
    return matrix_to_rotation_6d(axis_angle_to_matrix(aa.reshape(B, n, j3 // 3, 3))).reshape(B, n, j3 * 2)


def _to_aa(d6, joints):
    B, n, _ = d6.shape
    return matrix_to_axis_angle(rotation_6d_to_matrix(d6.reshape(B, n, joints, 6))).reshape(B, n, joints * 3)


class _CapturedPass:
    """One codec pass (fixed shapes) as a CUDA graph: static inputs -> static outputs.  The pass is ~700 eager
    launches (torch ops + library calls through ctypes) whose enqueue time, not their device time, bounds it once
    the GEMMs are on the tensor cores; replayed, it is one launch."""

    def __init__(self, fn, tensors):
        import threading
        self.lock = threading.Lock()
        self.static_in = [t.clone() for t in tensors]
        fn(*[t.clone() for t in self.static_in])          # eager once: weight-plane caches, cuBLAS workspaces
        torch.cuda.current_stream().synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # own capture stream (torch's default one is shared between threads: the pipeline's worker captures encode
        # passes while the main thread may capture a decode pass) and thread-local error mode (the other thread keeps
        # allocating and launching meanwhile)
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(tensors[0].device), capture_error_mode="thread_local"):
            out = fn(*self.static_in)
        self.static_out = list(out) if isinstance(out, (tuple, list)) else [out]

    def __call__(self, tensors):
        with self.lock:
            for s_, t in zip(self.static_in, tensors):
                s_.copy_(t)
            self.graph.replay()
            return [o.clone() for o in self.static_out]


class GestureRepEncoder(nn.Module):
    """diffusion_transformer.py:130-330: SMPL-X axis-angle body parts <-> [B, 4*n_chunks+3, D] latents with zero
    separator tokens (body_part_cat_axis="time") or concatenated along features (otherwise)."""
    PARTS = ("upper", "hands", "face", "lowertrans")      # order of the sampling draws in encode
    draws_on_device = True      # rsample noise comes from a CUDA generator: `generator` (None = the default one)
    generator = None
    gemm_tier = None            # set_gemm_tier()
    use_graphs = False          # enable_graphs(): encode / encode_many / decode passes replayed as CUDA graphs
    EXEMPLAR_BUCKET = 16        # encode_many pads its batch to a multiple of this under graphs (one graph per bucket)

    def __init__(self, vae_cfg, body_part_cat_axis="time", vaes=None):
        super().__init__()
        self.vae_cfg, self.body_part_cat_axis = vae_cfg, body_part_cat_axis
        self.frame_chunk_size, self.vae_latent_dim = vae_cfg["frame_chunk_size"], vae_cfg["latent_dim"]
        for part in ("upper", "face", "hands", "lowertrans"):          # construction order of the reference
            vae = vaes[part] if vaes is not None else self.load_vae(vae_cfg[f"{part}_cfg"])
            vae.eval()
            for p in vae.parameters():
                p.requires_grad = False
            setattr(self, f"{part}_vae", vae)
            setattr(self, f"{part}_latproj", nn.Identity())
        self.uj = self.lj = self.hj = self.fj = None
        self.tj = 3

    def enable_graphs(self, on=True):
        """Replay the codec's passes as CUDA graphs (CUDA tensors, inference only).  Results are those of the eager
        pass: the Gaussian draws are made outside the graph in the eager order, and a change of any parameter
        (storage or in-place version) or of the GEMM tier drops the captured passes."""
        self.use_graphs = bool(on)
        self.__dict__["_rg_graphs"] = {}
        return self

    def _run(self, kind, fn, tensors):
        if not (self.use_graphs and all(t.is_cuda for t in tensors)):
            return fn(*tensors)
        stamp = hash((getattr(self, "gemm_tier", None),) + tuple((p.data_ptr(), p._version) for p in self.parameters()))
        graphs = self.__dict__.setdefault("_rg_graphs", {})
        if graphs.get("stamp") != stamp:
            graphs.clear()
            graphs["stamp"] = stamp
        key = (kind,) + tuple((tuple(t.shape), t.dtype) for t in tensors)
        if key not in graphs:
            if len(graphs) > 12:                         # a few batch sizes + exemplar buckets; not a leak
                for k in [k for k in graphs if k != "stamp"][:4]:
                    del graphs[k]
            graphs[key] = _CapturedPass(fn, tensors)
        out = graphs[key](tensors)
        return out if len(out) > 1 else out[0]

    GEMM_TIERS = (None, "bf16x3", "bf16")

    def set_gemm_tier(self, tier):
        """None: every projection through F.linear (cuBLAS fp32).  "bf16x3" / "bf16": the 512/1024-wide projections
        of the four VAEs through the library's tcgen05 GEMM (see _lin); ReGestureTransformer sets "bf16x3" when its
        own precision tier is a tensor-core one, so the codec's results stay fp32-class in every tier."""
        assert tier in self.GEMM_TIERS, tier
        self.gemm_tier = tier
        for m in self.modules():
            m._rg_tier = tier
        return self

    @staticmethod
    def load_vae(cfg_path):
        """YAML hyper-parameters + the checkpoint named by its `test_ckpt`, looked up next to the YAML."""
        import yaml
        with open(cfg_path, "r", encoding="utf-8") as f:
            args = Namespace(**yaml.safe_load(f))
        model = TransformerVAE(args)
        GestureRepEncoder.load_checkpoints(model, os.path.join(os.path.dirname(cfg_path), os.path.basename(args.test_ckpt)))
        return model

    @staticmethod
    def load_checkpoints(model, path, load_name="model"):
        """`{'model_state': sd}`; a DataParallel 'module.' prefix on the first key means: try the stripped keys,
        fall back to the keys as stored (diffusion_transformer.py:169-187)."""
        sd = torch.load(path, map_location="cpu")["model_state"]
        first = next(iter(sd), "")
        if "module" in first:
            try:
                model.load_state_dict({k[7:]: v for k, v in sd.items()})
                return
            except RuntimeError:
                pass
        model.load_state_dict(sd)

    def _features(self, motion_upper, motion_lower, motion_face, motion_hands, motion_transl, motion_facial,
                  motion_contact):
        self.uj, self.lj = motion_upper.shape[-1] // 3, motion_lower.shape[-1] // 3
        self.hj, self.fj = motion_hands.shape[-1] // 3, motion_face.shape[-1] // 3
        motion_transl[:, :, 0] -= motion_transl[:, 0:1, 0].clone()       # in place, as the reference
        motion_transl[:, :, 2] -= motion_transl[:, 0:1, 2].clone()
        self.tj = motion_transl.shape[-1]
        return {"upper": _to_6d(motion_upper), "hands": _to_6d(motion_hands),
                "face": torch.cat([_to_6d(motion_face), motion_facial], dim=-1),
                "lowertrans": torch.cat([_to_6d(motion_lower), motion_transl, motion_contact], dim=-1)}

    def _assemble(self, z, motion_mask):
        m = motion_mask[:, ::self.frame_chunk_size]
        if self.body_part_cat_axis == "time":
            sep, msep = torch.zeros_like(z["upper"][:, :1]), torch.zeros_like(m[:, :1])
            motion = torch.cat([z["upper"], sep, z["hands"], sep, z["face"], sep, z["lowertrans"]], dim=1)
            return motion, torch.cat([m, msep, m, msep, m, msep, m], dim=1)
        sep = torch.zeros_like(z["upper"][:, :, :1])
        return torch.cat([z["upper"], sep, z["hands"], sep, z["face"], sep, z["lowertrans"]], dim=-1), m

    def _encode_core(self, motion_upper, motion_lower, motion_face, motion_hands, motion_transl, motion_facial,
                     motion_contact, motion_mask, eps):
        """eps [4, B*n_chunks, 1, D]: the rsample draw of each part (order PARTS)."""
        feats = self._features(motion_upper, motion_lower, motion_face, motion_hands, motion_transl, motion_facial,
                               motion_contact)
        z = {p: getattr(self, f"{p}_vae").encode_to_dist(feats[p], eps=eps[i])[0] for i, p in enumerate(self.PARTS)}
        motion, mask = self._assemble(z, motion_mask)
        return motion, mask, motion_transl               # the translation is re-based in place, as the reference does

    def _encode(self, inputs, eps):
        motion, mask, transl = self._run("encode", self._encode_core, list(inputs) + [eps])
        if transl.data_ptr() != inputs[4].data_ptr():    # replayed pass: hand the in-place edit back to the caller
            inputs[4].copy_(transl)
        return motion, mask

    def draw_encode_eps(self, motion_upper):
        """The Gaussian draws encode() makes for this batch ([4, B*n_chunks, 1, D], part order upper, hands, face,
        lowertrans as in the reference: one Normal(...).rsample() per part).  A caller that wants the draws at
        encode's place in the random stream but the encode itself later (MotionDiffusion.prepare) takes them here
        and passes them back as encode(..., eps=...)."""
        B, n = motion_upper.shape[0], motion_upper.shape[1] // self.frame_chunk_size
        gen = getattr(self, "generator", None)
        return torch.stack([torch.empty(B * n, 1, self.vae_latent_dim, device=motion_upper.device).normal_(generator=gen)
                            for _ in self.PARTS], 0)

    @torch.no_grad()
    def encode(self, motion_upper, motion_lower, motion_face, motion_hands, motion_transl, motion_facial,
               motion_contact, motion_mask, eps=None):
        if eps is None:
            eps = self.draw_encode_eps(motion_upper)
        return self._encode((motion_upper, motion_lower, motion_face, motion_hands, motion_transl, motion_facial,
                             motion_contact, motion_mask), eps)

    @torch.no_grad()
    def encode_many(self, motion_upper, motion_lower, motion_face, motion_hands, motion_transl, motion_facial,
                    motion_contact, motion_mask):
        """E exemplars in one pass through each VAE.  The Gaussian draws are made exemplar by exemplar, part by
        part, i.e. the sequence E separate encode() calls at B=1 consume (diffusion_architecture / raggesture.py:580
        encode one exemplar at a time)."""
        E, n = motion_upper.shape[0], motion_upper.shape[1] // self.frame_chunk_size
        dev, D = motion_upper.device, self.vae_latent_dim
        gen = getattr(self, "generator", None)
        eps = torch.stack([torch.stack([torch.empty(n, 1, D, device=dev).normal_(generator=gen) for _ in self.PARTS], 0)
                           for _ in range(E)], 0)                        # [E, 4, n, 1, D]
        eps = eps.transpose(0, 1).reshape(len(self.PARTS), E * n, 1, D)
        inputs = (motion_upper, motion_lower, motion_face, motion_hands, motion_transl, motion_facial, motion_contact,
                  motion_mask)
        pad = (-E) % self.EXEMPLAR_BUCKET if self.use_graphs and motion_upper.is_cuda else 0
        if pad == 0:
            return self._encode(inputs, eps)
        # one captured pass per bucket of exemplar counts: zero rows appended, their outputs dropped
        grow = lambda t, rows: torch.cat([t, t.new_zeros((rows,) + tuple(t.shape[1:]))], 0)
        padded = tuple(grow(t, pad) for t in inputs)
        eps_p = torch.cat([eps, eps.new_zeros(len(self.PARTS), pad * n, 1, D)], 1)
        motion, mask = self._encode(padded, eps_p)
        motion_transl.copy_(padded[4][:E])
        return motion[:E], mask[:E]

    @torch.no_grad()
    def decode(self, z_output):
        if self.uj is None:
            raise RuntimeError("GestureRepEncoder.decode before any encode: joint counts are recorded by encode "
                               "(diffusion_transformer.py:196-243), as in the reference")
        return tuple(self._run("decode", self._decode_core, [z_output]))

    def _decode_core(self, z_output):
        n = z_output.shape[1]
        if self.body_part_cat_axis == "time":
            n = (n - 3) // 4
            zu, zh = z_output[:, :n], z_output[:, n + 1:2 * n + 1]
            zf, zl = z_output[:, 2 * n + 2:3 * n + 2], z_output[:, 3 * n + 3:]
            assert zu.shape[1] == zh.shape[1] == zf.shape[1] == zl.shape[1]
        else:
            d = (z_output.shape[2] - 3) // 4
            zu, zh = z_output[..., :d], z_output[..., d + 1:2 * d + 1]
            zf, zl = z_output[..., 2 * d + 2:3 * d + 2], z_output[..., 3 * d + 3:]
            assert zu.shape[2] == self.vae_latent_dim == zh.shape[2] == zf.shape[2] == zl.shape[2]
        assert self.tj == 3
        upper = _to_aa(self.upper_vae.decode(zu), self.uj)
        hands = _to_aa(self.hands_vae.decode(zh), self.hj)
        face = self.face_vae.decode(zf)
        exps = face[:, :, self.fj * 6:].reshape(face.shape[0], face.shape[1], 100)
        facej = _to_aa(face[:, :, :self.fj * 6], self.fj)
        lt = self.lowertrans_vae.decode(zl)
        lower = _to_aa(lt[:, :, :self.lj * 6], self.lj)
        transl = lt[:, :, self.lj * 6:self.lj * 6 + self.tj]
        contact = lt[:, :, self.lj * 6 + self.tj:]
        return upper, lower, facej, hands, transl, exps, contact
