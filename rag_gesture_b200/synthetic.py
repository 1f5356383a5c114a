"""Deterministic synthetic weights and inputs (there is no dataset or checkpoint on the boxes).

Everything here is a pure function of integer seeds drawn from torch CPU generators, so the
golden-fixture script (run next to the reference), the CPU oracle tests and the GPU parity tests
all see bit-identical tensors without shipping 564 MB of weights.

Weight scales (see DESIGN.md "synthetic weights"): the reference zero-initialises `out`,
`ffn.linear2` and every `StylizationBlock.out_layers.2` (diffusion_transformer.py:79,415;
stylization_block.py:26), which makes a fresh model the zero function, so those are drawn
non-zero.  Cross-attention `value` projections are drawn small so that |y| < 1/32 on the
query-masked rows {10,20,30}: there the reference adds -1e6 in fp32 (efficient_attention.py:98),
i.e. rounds y to a 1/16 grid, and its output is a discontinuous function of y unless the row
collapses to exactly -1e6 (SURVEY 7 "-1e6 additive masks").  `normal_scale=True` lifts that
restriction for the single-step row-group test.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import config as C


def denoiser_param_shapes():
    """state-dict keys and shapes of ReGestureTransformer minus the VAEs, in the
    reference's own registration order."""
    D, E, F, TD = C.LATENT_DIM, C.TIME_EMBED_DIM, C.FF_SIZE, C.TEXT_DIM
    s = OrderedDict()
    s["sequence_embedding.pe"] = (C.N_CHUNKS, 1, D)
    s["global_positional_embedding.pe"] = (C.N_TOKENS, 1, D)
    s["text_pre_proj.weight"] = (D, TD)
    s["text_pre_proj.bias"] = (D,)
    s["audio_pre_proj.weight"] = (D, TD)
    s["audio_pre_proj.bias"] = (D,)
    s["speaker_embedding.weight"] = (C.NUM_SPEAKERS, D)
    s["joint_embed.weight"] = (D, D)
    s["joint_embed.bias"] = (D,)
    s["time_embed.0.weight"] = (E, D)
    s["time_embed.0.bias"] = (E,)
    s["time_embed.2.weight"] = (E, E)
    s["time_embed.2.bias"] = (E,)

    def styl(p):
        s[p + ".emb_layers.1.weight"] = (2 * D, E)
        s[p + ".emb_layers.1.bias"] = (2 * D,)
        s[p + ".norm.weight"] = (D,)
        s[p + ".norm.bias"] = (D,)
        s[p + ".out_layers.2.weight"] = (D, D)
        s[p + ".out_layers.2.bias"] = (D,)

    for l in range(C.NUM_LAYERS):
        p = f"temporal_decoder_blocks.{l}"
        s[p + ".sa_block.norm.weight"] = (D,)
        s[p + ".sa_block.norm.bias"] = (D,)
        for n in ("query", "key", "value"):
            s[f"{p}.sa_block.{n}.weight"] = (D, D)
            s[f"{p}.sa_block.{n}.bias"] = (D,)
        styl(p + ".sa_block.proj_out")
        for c in C.CONDS:
            q = f"{p}.ca_blocks.{c}"
            for n in ("norm", "text_norm"):
                s[f"{q}.{n}.weight"] = (D,)
                s[f"{q}.{n}.bias"] = (D,)
            for n in ("query", "key", "value"):
                s[f"{q}.{n}.weight"] = (D, D)
                s[f"{q}.{n}.bias"] = (D,)
            styl(q + ".proj_out")
        s[p + ".ca_mix.weight"] = (D, 3 * D)
        s[p + ".ca_mix.bias"] = (D,)
        s[p + ".ffn.linear1.weight"] = (F, D)
        s[p + ".ffn.linear1.bias"] = (F,)
        s[p + ".ffn.linear2.weight"] = (D, F)
        s[p + ".ffn.linear2.bias"] = (D,)
        styl(p + ".ffn.proj_out")
    s["out.weight"] = (D, D)
    s["out.bias"] = (D,)
    return s


def sine_position_table(max_len, d_model):
    """PositionEmbeddingSine1D buffer (detr_utils.py:33-39), shape [max_len, 1, d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def synthetic_state_dict(seed=0, normal_scale=False):
    """Key-ordered deterministic weights for the denoiser (fp32, CPU)."""
    g = torch.Generator().manual_seed(int(seed))
    sd = OrderedDict()
    for k, shp in denoiser_param_shapes().items():
        if k == "sequence_embedding.pe":
            sd[k] = sine_position_table(shp[0], shp[2])
            continue
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        if k == "global_positional_embedding.pe":
            v = 0.05 * r
        elif k == "speaker_embedding.weight":
            v = r / C.LATENT_DIM                      # diffusion_transformer.py:540-541
        elif k.endswith("norm.weight"):
            v = 1.0 + 0.1 * r
        elif k.endswith("norm.bias"):
            v = 0.1 * r
        elif k.endswith(".bias"):
            v = 0.02 * r
            if ".ca_blocks." in k and k.endswith("value.bias") and not normal_scale:
                v = (0.0005 if ".xf_spk." in k else 0.002) * r
        else:                                         # Linear weight [out, in]
            gain = 1.0
            if k.endswith("emb_layers.1.weight") or k.endswith("out_layers.2.weight"):
                gain = 0.5
            if ".ca_blocks." in k and k.endswith("value.weight") and not normal_scale:
                # xf_spk repeats one embedding row 150 times, so y == that value row (no
                # averaging over tokens): it needs the smaller gain to stay below 1/32
                gain = 0.003 if ".xf_spk." in k else 0.02
            v = r * (gain / math.sqrt(shp[-1]))
        sd[k] = v.contiguous()
    return sd


VAE_PART_FEATS = {"upper": 13 * 6, "hands": 30 * 6, "face": 1 * 6 + 100, "lowertrans": 9 * 6 + 3 + 4}


def vae_args(part, latent_dim=64, num_heads=2, ff_size=128, num_layers=3, arch="all_encoder",
             position_embedding="learned", vae_dist="normal", pre_norm=False, activation="gelu"):
    """Hyper-parameter dict of one body-part TransformerVAE (the keys of the reference's VAE YAMLs,
    gesture_vae.py:31-97; the shipped YAMLs and checkpoints are not in the reference repo)."""
    return dict(latent_dim=latent_dim, frame_chunk_size=C.FRAME_CHUNK, decoder_arch=arch,
                position_embedding=position_embedding, num_frames=C.MAX_SEQ_LEN, num_heads=num_heads,
                ff_size=ff_size, dropout=0.1, transformer_activation=activation,
                transformer_normalize_before=pre_norm, num_layers=num_layers, nfeats=VAE_PART_FEATS[part],
                vae_dist=vae_dist, test_ckpt=f"ckpt/{part}_vae.bin")


def synthetic_vae_state_dict(shapes, seed):
    """Key-ordered deterministic weights for a TransformerVAE given {key: shape} of its state dict."""
    g = torch.Generator().manual_seed(int(seed))
    sd = OrderedDict()
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        parts = k.split(".")
        if k.endswith(".pe"):
            v = 0.1 * r
        elif len(parts) >= 2 and "norm" in parts[-2] and k.endswith("weight"):
            v = 1.0 + 0.1 * r
        elif k.endswith("bias"):
            v = 0.05 * r
        elif k == "global_motion_token":
            v = 0.5 * r
        else:
            v = r * (1.0 / math.sqrt(shp[-1]))
        sd[k] = v.contiguous()
    return sd


def write_vae_files(root, seed0=300, latent_dim=512, **vae_kw):
    """Synthetic YAML + checkpoint per body part, laid out as the reference's load_vae expects
    (diffusion_transformer.py:151-167): the shipped VAE YAMLs / checkpoints are not in the reference repo, so the
    bench's `e2e.vae_codec` figure and the codec tests build the four TransformerVAEs from these files.  Returns the
    `vae_cfg` dict for the denoiser config."""
    import os
    import yaml
    from .vae import TransformerVAE
    cfg = {"frame_chunk_size": C.FRAME_CHUNK, "latent_dim": latent_dim}
    for i, part in enumerate(("upper", "hands", "face", "lowertrans")):
        args = vae_args(part, latent_dim=latent_dim, **vae_kw)
        d = os.path.join(root, part)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "cfg.yaml"), "w") as f:
            yaml.safe_dump(args, f)
        shapes = {k: tuple(v.shape) for k, v in TransformerVAE(args).state_dict().items()}
        torch.save({"model_state": synthetic_vae_state_dict(shapes, seed0 + i)},
                   os.path.join(d, os.path.basename(args["test_ckpt"])))      # load_vae looks next to the YAML
        cfg[f"{part}_cfg"] = os.path.join(d, "cfg.yaml")
    return cfg


def synthetic_conditions(n_clips, seed=1234, first_clip=0):
    """Per-clip condition features of the len150@15fps shape: BERT-like
    `word` [B,150,768], wav2vec-like `audio` [B,499,768], `speaker_ids` [B,150] int64.
    Clip i depends only on (seed, first_clip+i), so shards of a batch agree with the whole."""
    word = torch.empty(n_clips, C.N_TEXT, C.TEXT_DIM)
    audio = torch.empty(n_clips, C.N_AUDIO, C.TEXT_DIM)
    spk = torch.empty(n_clips, C.N_SPK, dtype=torch.int64)
    for i in range(n_clips):
        g = torch.Generator().manual_seed(int(seed) * 1_000_003 + first_clip + i)
        word[i] = torch.randn(C.N_TEXT, C.TEXT_DIM, generator=g)
        audio[i] = torch.randn(C.N_AUDIO, C.TEXT_DIM, generator=g)
        spk[i] = int(torch.randint(0, C.NUM_SPEAKERS, (1,), generator=g))
    return dict(word=word, audio=audio, speaker_ids=spk)


def synthetic_latents(n_clips, seed=99, first_clip=0, scale=1.0):
    """[B,43,512] latents with exact zeros on the separator rows {10,21,32} (what
    GestureRepEncoder.encode emits, diffusion_transformer.py:241-250)."""
    x = torch.empty(n_clips, C.N_TOKENS, C.LATENT_DIM)
    for i in range(n_clips):
        g = torch.Generator().manual_seed(int(seed) * 1_000_003 + first_clip + i)
        x[i] = scale * torch.randn(C.N_TOKENS, C.LATENT_DIM, generator=g)
    x[:, C.SEPARATOR_ROWS] = 0
    return x


def motion_mask(n_clips):
    """src/motion mask after encode: 1 everywhere, 0 on separators {10,21,32}."""
    m = torch.ones(n_clips, C.N_TOKENS)
    m[:, C.SEPARATOR_ROWS] = 0
    return m


def query_masks(n_clips):
    """cross-attention query masks: 0 on rows {10,20,30} (diffusion_architecture.py:152-166)."""
    m = torch.ones(n_clips, C.N_TOKENS)
    m[:, C.QUERY_MASK_ZERO_ROWS] = 0
    return {c: m.clone() for c in C.CONDS}


class NoiseTape:
    """Pre-drawn Gaussian draws handed out in call order.

    The reference draws from the global torch generator in a fixed order.  CPU and
    CUDA generators produce different streams, so parity tests draw the tape once on the CPU and
    replay it on both sides."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(int(seed))
        self.n = 0

    def randn(self, shape, device="cpu"):
        self.n += 1
        return torch.randn(tuple(shape), generator=self.g, dtype=torch.float32).to(device)


# ---- synthetic annotated dataset (schema of beatx_dataset.py:1262-1295) -------------
CONNECTIVES = ["and", "but", "so", "because", "then", "when", "if", "also", "while", "although",
               "however", "since", "though", "or", "as well", "so that"]
SENSES = ["Expansion.Conjunction", "Comparison.Contrast", "Contingency.Cause", "Temporal.Asynchronous",
          "Temporal.Synchronous", "Contingency.Condition", "Comparison.Concession", "Expansion.Alternative"]
FILLERS = ["the", "gesture", "people", "really", "think", "time", "maybe", "little"]
# words of the semantic gesture labels: shared stems and multi-word entries so that the word-similarity
# fall-back of the gesture-type rules (rag/utils.py:239-272) sees exact, partial and no matches
GESTURE_WORDS = ["this", "that", "these", "there", "here", "big", "bigger", "biggest", "small", "smaller", "round",
                 "around", "up", "upward", "down", "downward", "open", "open hand", "both hands", "together", "apart",
                 "you", "your", "me", "myself", "forward", "back", "backward", "one", "two", "first of all", "all"]
GESTURE_TYPES = ["beat", "deictic", "iconic", "metaphoric"]


class SyntheticGestureDataset:
    """In-memory stand-in for BEATXDataset: `ds[i]` / `ds[sample_name]` -> per-sample dict with the
    keys the hot path reads (raggesture.py:556-570, 244-276).  Every sample is a pure function of
    (seed, index); tensors are generated on access, annotations are cheap and cached."""

    def __init__(self, n, seed=7, frames=150, per_file=30):
        self.n, self.seed, self.frames, self.per_file = n, seed, frames, per_file
        self._ann, self._samples = {}, {}
        self.names = [self._name(i) for i in range(n)]
        self._by_name = {nm: i for i, nm in enumerate(self.names)}

    def _name(self, i):
        f, j = divmod(i, self.per_file)
        spk = (f * 7 + 3) % C.NUM_SPEAKERS
        return f"{spk}_synth_0_{f}_{f}/{j * 15 if j % 2 == 0 else j * 15 + 4}"

    def __len__(self):
        return self.n

    def __iter__(self):
        return (self[i] for i in range(self.n))

    def annotations(self, i):
        """(speaker, discourse, prominence, gesture_labels, n_text_tokens) of sample i."""
        if i in self._ann:
            return self._ann[i]
        import random
        r = random.Random(self.seed * 1_000_003 + i)
        f = i // self.per_file
        spk = (f * 7 + 3) % C.NUM_SPEAKERS
        discourse, prominence = [], []
        for _ in range(r.choice([0, 1, 1, 2, 2, 3])):
            conn = r.choice(CONNECTIVES)
            sense = r.choice(SENSES[:4]) if r.random() < 0.7 else r.choice(SENSES)
            ts = sorted(round(r.uniform(0, 10), 3) for _ in range(4))
            cs = round(r.uniform(ts[0], ts[3]), 3)
            ce = round(min(10.0, cs + r.uniform(0.15, 0.8)), 3)
            discourse.append((conn, sense, "arg1", "arg2", ts[0], ts[3], cs, ce))
        discourse.sort(key=lambda d: d[6])
        for w in r.sample(FILLERS, 3):
            s = round(r.uniform(0, 9.5), 3)
            prominence.append((w, s, round(s + 0.3, 3), round(r.uniform(0, 3), 4)))
        for d in discourse:
            if r.random() < 0.65:                      # the rest map to None (integer-score tiers)
                for w in d[0].split():
                    prominence.append((w, d[6], d[7], round(r.uniform(0, 3), 4)))
        prominence.sort(key=lambda p: p[1])
        gestures = [{"name": r.choice(GESTURE_TYPES), "start": d[6],
                     "end": d[7], "word": d[0]} for d in discourse[:1]]
        # further semantic gesture labels from their own stream (the draws above stay what the committed goldens saw)
        r2 = random.Random(self.seed * 7_368_787 + 31 * i + 5)
        for _ in range(r2.choice([0, 1, 1, 2, 3])):
            gs = round(r2.uniform(0, 9.3), 3)
            gestures.append({"name": r2.choice(GESTURE_TYPES), "start": gs,
                             "end": round(min(10.0, gs + r2.uniform(0.2, 1.6)), 3), "word": r2.choice(GESTURE_WORDS)})
        gestures.sort(key=lambda g_: g_["start"])
        self._ann[i] = (spk, discourse, prominence, gestures, r.randint(8, 32))
        return self._ann[i]

    def text_feature(self, i):
        n_tok = self.annotations(i)[4]
        g = torch.Generator().manual_seed(self.seed * 7_000_003 + i)
        return torch.randn(n_tok, C.TEXT_DIM, generator=g)

    def __getitem__(self, key):
        """Samples are materialised once and then served from memory (the stand-in for the dataset's
        LMDB cache; generating 0.5 M Gaussians per access would dominate the host side)."""
        i = self._by_name[key] if isinstance(key, str) else int(key)
        if i not in self._samples:
            self._samples[i] = self._make(i)
        return dict(self._samples[i])

    def _make(self, i):
        spk, discourse, prominence, gestures, _ = self.annotations(i)
        g = torch.Generator().manual_seed(self.seed * 9_000_011 + i)
        F = self.frames
        rn = lambda *s: 0.1 * torch.randn(*s, generator=g)
        upper, lower, face, hands = rn(F, 39), rn(F, 27), rn(F, 3), rn(F, 90)
        trans, facial, contact = rn(F, 3), rn(F, 100), rn(F, 4)
        return dict(
            motion=torch.cat([upper, lower, face, hands, rn(F, 6)], -1), motion_upper=upper,
            motion_lower=lower, motion_face=face, motion_hands=hands, trans=trans, facial=facial,
            contact=contact, motion_mask=torch.ones(F), motion_length=F,
            word=torch.randn(C.N_TEXT, C.TEXT_DIM, generator=g),
            audio=torch.randn(C.N_AUDIO, C.TEXT_DIM, generator=g),
            speaker_id=torch.full((F,), spk, dtype=torch.int64), text_feature=self.text_feature(i),
            raw_word="", raw_audio=None, text_segments=[], discourse=list(discourse),
            prominence=list(prominence), gesture_labels=list(gestures), sample_name=self.names[i],
            sample_idx=i)


def collate(samples):
    """beatx_collate_fn (mogen/datasets/builder.py:59-91): tensors stacked, list fields kept as lists,
    speaker_id -> speaker_ids, text_feature -> text_features."""
    out = {}
    for k in samples[0]:
        vals = [s[k] for s in samples]
        if isinstance(vals[0], torch.Tensor) and k != "text_feature":
            out[k] = torch.stack(vals, 0)
        else:
            out[k] = vals
    out["speaker_ids"] = out.pop("speaker_id")
    out["text_features"] = out.pop("text_feature")
    out["motion_length"] = torch.tensor(out["motion_length"])
    return out
