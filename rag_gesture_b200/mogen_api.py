"""Host-side mirror of the reference's registry classes on the hot path.

Same class names, constructor arguments, state-dict keys and call signatures as
  mogen/models/builder.py                       (registries, build_* helpers)
  mogen/models/utils/stylization_block.py       StylizationBlock
  mogen/models/attentions/efficient_attention.py EfficientSelfAttention / EfficientCrossAttention
  mogen/models/transformers/diffusion_transformer.py  FFN, DecoderLayer, DiffusionTransformer.forward
  mogen/models/transformers/raggesture.py       ReGestureTransformer
so that configs select them by `type=` and `load_checkpoint` fills them unchanged.  torch.nn classes
are used as PARAMETER CONTAINERS only (they give the exact key names); every forward() below is a
sequence of CUDA-kernel calls through the C ABI -- there is no PyTorch arithmetic and no CPU path.

Two execution paths, both on the GPU:
  * fused path  -- ReGestureTransformer.forward / the samplers in diffusion.py: one rg_denoise call
    per step over the packed weights, timestep table (K7) and per-clip cross-attention state (K6);
  * module path -- calling a DecoderLayer / attention / StylizationBlock / forward_test directly
    (arbitrary `emb`, per-module weights): composed from the rg_op_* kernels, recomputing K/V like
    the reference does.  Used for API fidelity and for unit parity tests.
"""
import copy

import torch
import torch.nn as nn

from . import _lib, ops
from . import config as CFG
from .engine import DenoiserEngine


# ---- registry (mogen/models/builder.py:5-36) -------------------------------------------------------
class Registry:
    """Minimal stand-in with mmcv.utils.Registry's surface used by the reference: register_module
    (decorator, force=), get, build(cfg with `type`).  When the real mogen/mmcv is importable,
    `register_into(mogen.models.builder.MODELS)` re-registers these classes under the same names."""

    def __init__(self, name):
        self.name, self._module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._module_dict and not force and self._module_dict[key] is not cls:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._module_dict[key] = cls
            return cls
        return _reg(module) if module is not None else _reg

    def get(self, key):
        return self._module_dict.get(key)

    def build(self, cfg, default_args=None):
        if cfg is None:
            return None
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        t = args.pop("type")
        cls = self.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry")
        return cls(**args)


MODELS = Registry("models")
LOSSES = ARCHITECTURES = SUBMODULES = ATTENTIONS = MODELS


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_architecture(cfg, **kwargs):
    name = cfg.pop("type")
    return MODELS.get(name)(**cfg, **kwargs)


def build_submodule(cfg, **kwargs):
    name = cfg.pop("type")
    return SUBMODULES.get(name)(**cfg, **kwargs)


def build_attention(cfg):
    return ATTENTIONS.build(cfg)


def register_into(registry, force=True):
    """Re-register the B200 classes into the reference's own registry (drop-in for visualize.py)."""
    for name, cls in MODELS._module_dict.items():
        registry.register_module(name=name, force=force, module=cls)


def _no_grad_needed(*ts):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts):
        raise NotImplementedError("rg_b200 kernels are inference-only (wrap the call in torch.no_grad())")


# ---- StylizationBlock (stylization_block.py:14-40) ---------------------------------------------------
class StylizationBlock(nn.Module):
    def __init__(self, latent_dim, time_embed_dim, dropout):
        super().__init__()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(time_embed_dim, 2 * latent_dim))
        self.norm = nn.LayerNorm(latent_dim)
        self.out_layers = nn.Sequential(nn.SiLU(), nn.Dropout(p=dropout), nn.Linear(latent_dim, latent_dim))
        for p in self.out_layers[2].parameters():      # zero_module(...) of the reference
            p.detach().zero_()

    def scale_shift(self, emb):
        """emb [B,E] -> [B, 2*latent] (scale | shift)."""
        return ops.linear(ops.silu(emb), self.emb_layers[1].weight, self.emb_layers[1].bias)

    def forward(self, h, emb, residual=None):
        B, T, D = h.shape
        a = ops.stylization_rows(h.reshape(B * T, D), self.norm.weight, self.norm.bias,
                                 self.scale_shift(emb), T)
        out = ops.linear(a, self.out_layers[2].weight, self.out_layers[2].bias,
                         residual=None if residual is None else residual.reshape(B * T, D))
        return out.view(B, T, D)


# ---- attentions (efficient_attention.py) ---------------------------------------------------------------
@ATTENTIONS.register_module()
class EfficientSelfAttention(nn.Module):
    def __init__(self, latent_dim, num_heads, dropout, time_embed_dim=None):
        super().__init__()
        self.num_heads = num_heads
        self.norm = nn.LayerNorm(latent_dim)
        self.query = nn.Linear(latent_dim, latent_dim)
        self.key = nn.Linear(latent_dim, latent_dim)
        self.value = nn.Linear(latent_dim, latent_dim)
        self.dropout = nn.Dropout(dropout)
        self.time_embed_dim = time_embed_dim
        if time_embed_dim is not None:
            self.proj_out = StylizationBlock(latent_dim, time_embed_dim, dropout)

    def forward(self, x, src_mask, emb=None, **kwargs):
        B, T, D = x.shape
        _no_grad_needed(x)
        w = torch.cat([self.query.weight, self.key.weight, self.value.weight], 0)
        b = torch.cat([self.query.bias, self.key.bias, self.value.bias], 0)
        qkv = ops.linear(ops.layernorm(x, self.norm.weight, self.norm.bias), w, b)
        mask = src_mask.reshape(B, T)
        if self.time_embed_dim is None:
            return ops.self_attention(qkv, mask, x_res=x)
        po = self.proj_out
        a = ops.self_attention(qkv, mask, po.norm.weight, po.norm.bias, po.scale_shift(emb))
        return ops.linear(a, po.out_layers[2].weight, po.out_layers[2].bias, residual=x)


@ATTENTIONS.register_module()
class EfficientCrossAttention(nn.Module):
    def __init__(self, latent_dim, text_latent_dim, num_heads, dropout, time_embed_dim):
        super().__init__()
        self.num_heads = num_heads
        self.norm = nn.LayerNorm(latent_dim)
        self.text_norm = nn.LayerNorm(text_latent_dim)
        self.query = nn.Linear(latent_dim, latent_dim)
        self.key = nn.Linear(text_latent_dim, latent_dim)
        self.value = nn.Linear(text_latent_dim, latent_dim)
        self.dropout = nn.Dropout(dropout)
        self.proj_out = StylizationBlock(latent_dim, time_embed_dim, dropout)

    def kv_state(self, xf):
        """softmax_tokens(key(text_norm(xf)))^T value(text_norm(xf)) -> [B,16,32,32]."""
        B, N, _ = xf.shape
        w = torch.cat([self.key.weight, self.value.weight], 0)
        b = torch.cat([self.key.bias, self.value.bias], 0)
        kv = ops.linear(ops.layernorm(xf, self.text_norm.weight, self.text_norm.bias), w, b)
        return ops.kv_state(kv.view(B * N, -1), B, N)

    def forward(self, x, xf, emb, query_mask, cond_type=None, **kwargs):
        _no_grad_needed(x, xf)
        if cond_type is not None and not bool(((cond_type % 10) > 0).all()):
            raise NotImplementedError("cond_type=0 (unconditional branch of the 2-branch mode) is out "
                                      "of scope (SURVEY 8f.4)")
        q = ops.linear(ops.layernorm(x, self.norm.weight, self.norm.bias), self.query.weight, self.query.bias)
        po = self.proj_out
        a = ops.cross_attention(q, self.kv_state(xf), query_mask, po.norm.weight, po.norm.bias,
                                po.scale_shift(emb))
        return ops.linear(a, po.out_layers[2].weight, po.out_layers[2].bias, residual=x)


# ---- FFN / DecoderLayer (diffusion_transformer.py:74-127) ------------------------------------------------
class FFN(nn.Module):
    def __init__(self, latent_dim, ffn_dim, dropout, time_embed_dim):
        super().__init__()
        self.linear1 = nn.Linear(latent_dim, ffn_dim)
        self.linear2 = nn.Linear(ffn_dim, latent_dim)
        for p in self.linear2.parameters():
            p.detach().zero_()
        self.activation = nn.GELU()
        self.dropout = nn.Dropout(dropout)
        self.proj_out = StylizationBlock(latent_dim, time_embed_dim, dropout)

    def forward(self, x, emb, **kwargs):
        _no_grad_needed(x)
        y = ops.linear(ops.linear(x, self.linear1.weight, self.linear1.bias, epilogue=_lib.OP_GELU),
                       self.linear2.weight, self.linear2.bias)
        return self.proj_out(y, emb, residual=x)


class DecoderLayer(nn.Module):
    def __init__(self, sa_block_cfg=None, ca_block_cfg=None, ffn_cfg=None):
        super().__init__()
        self.sa_block = build_attention(sa_block_cfg)
        self.ca_blocks = nn.ModuleDict({c: build_attention(ca_block_cfg) for c in CFG.CONDS})
        self.ca_mix = nn.Linear(ffn_cfg["latent_dim"] * 3, ffn_cfg["latent_dim"])
        self.ffn = FFN(**ffn_cfg)

    def forward(self, **kwargs):
        x = self.sa_block(**kwargs)
        kwargs["x"] = x
        xf = kwargs.pop("xf")
        qms = kwargs.pop("query_mask")
        outs = [self.ca_blocks[c](xf=xf_c, query_mask=None if qms is None else qms[c], **kwargs)
                for c, xf_c in xf.items()]
        kwargs["x"] = ops.linear(torch.cat(outs, dim=-1), self.ca_mix.weight, self.ca_mix.bias)
        return self.ffn(**kwargs)


class _PositionTable(nn.Module):
    """Holds `pe` under the reference's key (detr_utils.py:27-79); the add happens in the kernel."""

    def __init__(self, d_model, max_len, learned):
        super().__init__()
        if learned:
            self.pe = nn.Parameter(torch.randn(max_len, 1, d_model))
            nn.init.xavier_uniform_(self.pe)
        else:
            from .synthetic import sine_position_table
            self.register_buffer("pe", sine_position_table(max_len, d_model))


class PreparedBatch:
    """Per-batch, step-invariant inputs of the fused path."""

    def __init__(self, src_mask, query_mask, state):
        self.src_mask, self.query_mask, self.state = src_mask, query_mask, state


# ---- the denoiser (diffusion_transformer.py:334-668, raggesture.py:887-1113) --------------------------------
@SUBMODULES.register_module()
class ReGestureTransformer(nn.Module):
    def __init__(self, retrieval_cfg=None, scale_func_cfg=None, per_joint_scale=None,
                 retrieval_train=False, use_retrieval_for_test=False, database=None,
                 input_feats=CFG.INPUT_FEATS, max_seq_len=240, frame_chunk_size=16, latent_dim=512,
                 time_embed_dim=2048, num_layers=8, sa_block_cfg=None, ca_block_cfg=None, vae_cfg=None,
                 ffn_cfg=None, text_encoder=None, audio_encoder=None, speaker_embedding=None,
                 use_cache_for_text=False, init_cfg=None, body_part_cat_axis="time",
                 gesture_rep_encoder=None, precision=_lib.PREC_FP32):
        super().__init__()
        assert not retrieval_train
        # 2-branch mode (raggesture.py:908-921): scale_func_cfg switches it on; the reference builds
        # joint_scale_mask only when per_joint_scale is given and otherwise raises AttributeError in the first
        # forward_test (:1102) -- the shipped config does exactly that.  Mirrored: see two_branch_x0.
        if body_part_cat_axis != "time":
            raise NotImplementedError("Only time axis is supported for body part categorization")
        for enc, nm in ((text_encoder, "text"), (audio_encoder, "audio")):
            if enc.get("pretrained_model") is not None or enc.get("num_layers", 0) > 0 or enc.get("use_text_proj"):
                raise NotImplementedError(f"{nm}_encoder: only pretrained_model=None, num_layers=0 (shipped config)")
        self.input_feats, self.latent_dim, self.num_layers = input_feats, latent_dim, num_layers
        self.time_embed_dim, self.frame_chunk_size = time_embed_dim, frame_chunk_size
        self.body_part_cat_axis = body_part_cat_axis
        self.scale_func_cfg, self.per_joint_scale = scale_func_cfg, per_joint_scale
        if per_joint_scale is not None:
            T = 43                                      # hard-coded in the reference (:911)
            n = (T - 3) // 4
            self.joint_scale_mask = torch.ones(T)
            self.joint_scale_mask[0:n] = per_joint_scale["upper"]
            self.joint_scale_mask[n + 1:2 * n + 1] = per_joint_scale["hands"]
            self.joint_scale_mask[2 * n + 2:3 * n + 2] = per_joint_scale["face"]
            self.joint_scale_mask[3 * n + 3:T] = per_joint_scale["lowertransl"]
        self.precision = precision
        # latent codec (adjacent component): any object with encode/decode/vae_latent_dim
        if gesture_rep_encoder is None and vae_cfg is not None:
            from .codec import build_codec
            gesture_rep_encoder = build_codec(vae_cfg, body_part_cat_axis)
        self.gesture_rep_encoder = gesture_rep_encoder
        if precision != _lib.PREC_FP32 and hasattr(gesture_rep_encoder, "set_gemm_tier"):
            gesture_rep_encoder.set_gemm_tier("bf16x3")        # F1: the codec's GEMMs on the tensor cores too
        n_chunks = max_seq_len if gesture_rep_encoder is None else max_seq_len // frame_chunk_size
        self.max_seq_len = n_chunks
        if gesture_rep_encoder is None:
            raise NotImplementedError("rg_b200 denoiser works on the 4-part VAE latents (vae_cfg required)")
        self.sequence_embedding = _PositionTable(latent_dim, n_chunks, learned=False)
        self.global_positional_embedding = _PositionTable(latent_dim, n_chunks * 4 + 3, learned=True)
        self.text_pre_proj = nn.Linear(text_encoder["latent_dim"], latent_dim)
        self.audio_pre_proj = nn.Linear(audio_encoder["latent_dim"], latent_dim)
        self.num_speakers = speaker_embedding["num_speakers"]
        self.speaker_embedding = nn.Embedding(self.num_speakers, latent_dim)
        self.speaker_embedding.weight.data.normal_(mean=0, std=1)
        self.speaker_embedding.weight.data /= latent_dim
        self.joint_embed = nn.Linear(gesture_rep_encoder.vae_latent_dim, latent_dim)
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, time_embed_dim), nn.SiLU(),
                                        nn.Linear(time_embed_dim, time_embed_dim))
        self.temporal_decoder_blocks = nn.ModuleList(
            DecoderLayer(sa_block_cfg=sa_block_cfg, ca_block_cfg=ca_block_cfg, ffn_cfg=ffn_cfg)
            for _ in range(num_layers))
        self.out = nn.Linear(latent_dim, gesture_rep_encoder.vae_latent_dim)
        for p in self.out.parameters():
            p.detach().zero_()
        self._ffn_dim = ffn_cfg["ffn_dim"]
        self._num_heads = sa_block_cfg["num_heads"]
        self._text_dim = text_encoder["latent_dim"]
        if retrieval_cfg is not None and use_retrieval_for_test:
            from .retrieval import RetrievalDatabase
            self.database = RetrievalDatabase(**retrieval_cfg, dataset=database)
        else:
            self.database = None
        self._engine, self._engine_key, self._sched_key = None, None, None
        self._epoch = getattr(self, "_epoch", 0)
        self._state_cache = (None, None)

    # -- engine lifetime: rebuilt when weights move or change (load_state_dict, .to(), .cuda()) ------
    def _weights_key(self):
        """Fingerprint of the parameter set the packed device engine was built from: the epoch bumped by this
        module's load_state_dict / .to() / .cuda() hooks plus the storage address and in-place version counter of
        EVERY denoiser parameter and buffer.  A checkpoint loaded through a parent module
        (MotionDiffusion.load_state_dict, mmcv load_checkpoint: in-place copies into the same Parameter objects), a
        strict=False partial load or an in-place edit of any tensor therefore rebuilds the engine.  The tensor list
        is collected once per epoch (walking state_dict() costs ~2 ms; this runs on every rg_engine() call, i.e.
        several times per batch), the fingerprint itself is ~100 us.  Edits that bypass the version counter (through
        .data) or that REPLACE a Parameter object need `invalidate_engine()`."""
        cached = getattr(self, "_key_tensors", None)
        if cached is None or cached[0] != self._epoch:
            tensors = [t for name, t in self.state_dict(keep_vars=True).items()
                       if not name.startswith(("gesture_rep_encoder.", "database."))]
            cached = self._key_tensors = (self._epoch, tensors)
        ptr = ver = 0
        for t in cached[1]:
            ptr = (ptr * 1000003 + t.data_ptr()) & 0xFFFFFFFFFFFFFFFF
            ver += t._version
        return (self._epoch, ptr, ver)

    def invalidate_engine(self):
        self._epoch += 1

    def _apply(self, fn, *a, **k):
        self._epoch = getattr(self, "_epoch", 0) + 1
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._epoch = getattr(self, "_epoch", 0) + 1
        return super().load_state_dict(*a, **k)

    def rg_engine(self, diffusion=None):
        dev = self.out.weight.device
        if dev.type != "cuda":
            raise RuntimeError("rg_b200: move the model to a CUDA device first (no CPU fallback)")
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith(("gesture_rep_encoder.", "database."))}
            self._engine = DenoiserEngine(sd, latent_dim=self.latent_dim, num_heads=self._num_heads,
                                          ffn_dim=self._ffn_dim, time_embed_dim=self.time_embed_dim,
                                          num_layers=self.num_layers, n_chunks=self.max_seq_len,
                                          text_dim=self._text_dim, num_speakers=self.num_speakers,
                                          precision=self.precision, device=dev)
            self._engine_key, self._sched_key, self._state_cache = key, None, (None, None)
        if diffusion is not None:
            # keyed on the schedule's CONTENT (timestep map + alpha-bar table): id() can alias after garbage collection
            skey = getattr(diffusion, "_rg_sched_key", None)
            if skey is None or skey[0] != len(diffusion.timestep_map):
                skey = (len(diffusion.timestep_map), tuple(int(t) for t in diffusion.timestep_map),
                        diffusion.alphas_cumprod.tobytes())
                try:
                    diffusion._rg_sched_key = skey      # content hash computed once per schedule object
                except AttributeError:
                    pass
            if self._sched_key != skey:
                self._engine.set_schedule(diffusion.timestep_map, diffusion.coef_table())
                self._sched_key = skey
        return self._engine

    # -- conditions (raggesture.py:957-1013) -----------------------------------------------------------
    def encode_text(self, text, device):
        return self.rg_engine().encode_conditions(word=text.to(device))["xf_text"]

    def encode_audio(self, audio, device):
        return self.rg_engine().encode_conditions(audio=audio.to(device))["xf_audio"]

    def encode_spks(self, spk_ids, device):
        if self.num_speakers == 1:
            # diffusion_transformer.py:545-546, shape quirk included: zeros [B, B, latent] (B "speaker tokens").  All
            # tokens are equal, so the cross-attention state does not depend on their number and batching the
            # exemplars' inversions (one zeros [E, E, latent] instead of E times [1, 1, latent]) changes nothing.
            return torch.zeros((spk_ids.shape[0], spk_ids.shape[0], self.latent_dim), device=device)
        return self.rg_engine().encode_conditions(speaker_ids=spk_ids.to(device))["xf_spk"]

    def encode_all_conditions(self, text, audio, speaker_ids, device):
        """xf_text / xf_audio / xf_spk of a batch in one library call (raggesture.py:978-987)."""
        eng = self.rg_engine()
        if self.num_speakers == 1:
            xf = eng.encode_conditions(text.to(device), audio.to(device), None)
            xf["xf_spk"] = self.encode_spks(speaker_ids, device)
            return xf
        return eng.encode_conditions(text.to(device), audio.to(device), speaker_ids.to(device))

    def get_precompute_condition(self, text=None, raw_text=None, text_features=None, audio=None,
                                 raw_audio=None, discourse=None, prominence=None, speaker_ids=None,
                                 gesture_labels=None, text_times=None, motion_length=None, xf_out=None,
                                 re_dict=None, device=None, sample_idx=None, sample_name=None,
                                 retrieval_method="gesture_type", **kwargs):
        if xf_out is None:
            device = device if device is not None else self.out.weight.device
            xf_out = self.encode_all_conditions(text, audio, speaker_ids, device)
        output = {"xf_out": xf_out}
        if re_dict is None and self.database is not None:
            retr_conditions = dict(text=raw_text, audio=raw_audio, text_enc=text, text_features=text_features,
                                   audio_enc=audio, discourse=discourse, prominence=prominence,
                                   speaker_ids=speaker_ids, gesture_labels=gesture_labels, text_times=text_times)
            re_dict = self.database(retr_conditions, motion_length, device, idx=sample_name,
                                    retrieval_method=retrieval_method,
                                    gesture_rep_encoder=self.gesture_rep_encoder)
        output["re_dict"] = re_dict
        return output

    def post_process(self, motion):
        return motion

    # -- fused path -------------------------------------------------------------------------------------
    def prepare_batch(self, model_kwargs, B):
        """K6 state + packed masks for a batch.  The state is cached while the SAME xf_out tensor objects
        (unchanged `_version`) are passed again -- the sampler calls this once per loop, `forward` once per
        step.  The cache holds references to those tensors, so their storage cannot be recycled for a
        different batch behind an identical data_ptr."""
        eng = self.rg_engine()
        xf = model_kwargs["xf_out"]
        ts = tuple(xf[c] for c in CFG.CONDS)
        vs = tuple(t._version for t in ts)
        cached = self._state_cache[0]
        hit = cached is not None and all(a is b for a, b in zip(cached[0], ts)) and cached[1] == vs
        if not hit:
            self._state_cache = ((ts, vs), eng.precompute_state(xf))
        state = self._state_cache[1]
        dev = state.device
        mm = model_kwargs["motion_mask"]
        src_mask = mm.reshape(B, -1).to(device=dev, dtype=torch.float32).contiguous()
        qm = model_kwargs.get("query_mask", None)
        if qm is not None:
            qm = torch.stack([qm[c].to(device=dev, dtype=torch.float32) for c in CFG.CONDS], 0).contiguous()
        if self.two_branch:
            # forward_test repeats the batch (:1058-1072): text branch, then the "none" branch, whose cross-attention
            # sees keys - 1e6 and the values of a zeroed condition, i.e. the state A[d][l] = value.bias[l]
            state = torch.cat([state, self._none_state(dev).expand(B, -1, -1, -1, -1, -1)], 0).contiguous()
            src_mask = src_mask.repeat(2, 1)
            qm = None if qm is None else qm.repeat(1, 2, 1).contiguous()
        return PreparedBatch(src_mask, qm, state)

    # -- 2-branch mode (raggesture.py:925-954, 1041-1111) ------------------------------------------------------------
    @property
    def two_branch(self):
        return self.scale_func_cfg is not None

    def scale_func_retr(self, timestep):
        """The four mixing coefficients (:925-954); above t = 100 one of two sets is drawn with Python's global
        `random`, as in the reference."""
        import random
        cfg = self.scale_func_cfg
        w = (1 - (1000 - timestep) / 1000) * cfg["coarse_scale"] + 1
        if timestep > 100:
            if random.randint(0, 1) == 0:
                return {"both_coef": w, "text_coef": 0, "retr_coef": 1 - w, "none_coef": 0}
            return {"both_coef": 0, "text_coef": w, "retr_coef": 0, "none_coef": 1 - w}
        both, text, retr = cfg["both_coef"], cfg["text_coef"], cfg["retr_coef"]
        return {"both_coef": both, "text_coef": text, "retr_coef": retr, "none_coef": 1 - both - text - retr}

    def _none_state(self, dev):
        """[1, L, 3, H, 32, 32]: the K6 state of the "none" branch.  With cond_type 0 every value row is value.bias and
        the key softmax still sums to one over the tokens (efficient_attention.py:83-89), so A[d][l] = value.bias[l]
        for every d, whatever the condition is."""
        key = self._weights_key()
        cached = getattr(self, "_none_state_cache", None)
        if cached is None or cached[0] != key or cached[1].device != dev:
            H = self._num_heads
            rows = []
            for blk in self.temporal_decoder_blocks:
                per_cond = [blk.ca_blocks[c].value.bias.detach().float().view(H, 1, -1).expand(H, self.latent_dim // H, -1)
                            for c in CFG.CONDS]
                rows.append(torch.stack(per_cond, 0))
            cached = self._none_state_cache = (key, torch.stack(rows, 0).unsqueeze(0).to(dev).contiguous())
        return cached[1]

    def two_branch_x0(self, eng, prep, x, step_idx=-1, tau=0, coefs=None):
        """forward_test with scale_func_cfg: both branches in ONE evaluation of 2B clips, then rg_mix_branches.
        `coefs` [B,4] (both, text, retr, none per clip) overrides the per-call draw: the batched inversion loop draws
        them exemplar by exemplar in the order the reference's per-exemplar loops would."""
        if not hasattr(self, "joint_scale_mask"):
            raise AttributeError("'ReGestureTransformer' object has no attribute 'joint_scale_mask' (scale_func_cfg "
                                 "without per_joint_scale: the reference fails the same way, raggesture.py:1102)")
        B = x.shape[0]
        if coefs is None:
            t_orig = int(tau) if step_idx < 0 else int(eng.timestep_map[step_idx])
            c = self.scale_func_retr(t_orig)
            coefs = torch.tensor([[c["both_coef"], c["text_coef"], c["retr_coef"], c["none_coef"]]],
                                 dtype=torch.float32).expand(B, 4)
        coefs = coefs.to(device=x.device, dtype=torch.float32).contiguous()
        js = self.joint_scale_mask.to(device=x.device, dtype=torch.float32).contiguous()
        out2 = eng.denoise(x.repeat(2, 1, 1), prep.src_mask, prep.query_mask, prep.state, step_idx=step_idx, tau=tau)
        return eng.mix_branches(out2, coefs, js)

    def forward(self, motion, timesteps, motion_mask=None, **kwargs):
        """motion [B,T,D], timesteps [B] on the ORIGINAL 0..999 scale (all equal) -> x0 [B,T,D]."""
        if self.training:
            raise NotImplementedError("rg_b200 is inference-only; call model.eval() (training is out of scope)")
        _no_grad_needed(motion)
        B = motion.shape[0]
        tau = int(timesteps[0])
        if not bool((timesteps == tau).all()):
            raise NotImplementedError("all clips of a batch must share one timestep")
        if kwargs.get("do_clf_guidance", False):
            raise NotImplementedError("classifier-free guidance (2-branch mode) is out of scope")
        cond = self.get_precompute_condition(device=motion.device, **kwargs)
        prep = self.prepare_batch({"xf_out": cond["xf_out"], "motion_mask": motion_mask,
                                   "query_mask": copy.copy(kwargs.get("query_mask", None))}, B)
        if self.two_branch:
            return self.two_branch_x0(self.rg_engine(), prep, motion.float().contiguous(), step_idx=-1, tau=tau)
        return self.rg_engine().denoise(motion, prep.src_mask, prep.query_mask, prep.state, step_idx=-1, tau=tau)

    # -- module path (raggesture.py:1041-1113 single branch) ------------------------------------------------
    def forward_train(self, *a, **k):
        raise NotImplementedError("training is out of scope of rg_b200")

    def forward_test(self, h=None, src_mask=None, emb=None, xf_out=None, query_mask=None,
                     timesteps=None, do_clf_guidance=False, **kwargs):
        if do_clf_guidance:
            raise NotImplementedError("classifier-free guidance (2-branch mode) is out of scope")
        for module in self.temporal_decoder_blocks:
            h = module(x=h, xf=xf_out, emb=emb, src_mask=src_mask, query_mask=query_mask, cond_type=None)
        return ops.linear(h, self.out.weight, self.out.bias)
