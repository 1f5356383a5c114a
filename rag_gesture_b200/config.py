"""Hyper-parameters of the hot path, as shipped by the reference config.

Mirrors /root/reference configs/raggesture_beatx/basegesture_len150_beat.py:32-160 (the `model`
dict that `tools/visualize.py:138` hands to `build_architecture`).  Only `scale_func_cfg` deviates:
the shipped value makes `forward_test` raise AttributeError (raggesture.py:1102), so the
single-branch mode (`scale_func_cfg=None`) is the configuration every parity claim is made on.
"""
import copy

INPUT_FEATS = 189
MAX_SEQ_LEN = 150          # frames @ 15 fps
FRAME_CHUNK = 15
MOTION_FPS = 15
LATENT_DIM = 512
TIME_EMBED_DIM = 2048
TEXT_DIM = 768
FF_SIZE = 1024
NUM_HEADS = 16
NUM_LAYERS = 8
NUM_SPEAKERS = 25
N_CHUNKS = MAX_SEQ_LEN // FRAME_CHUNK          # 10 latent tokens per body part
N_TOKENS = 4 * N_CHUNKS + 3                    # 43 = 4 parts + 3 zero separators
N_TEXT = 150
N_AUDIO = 499
N_SPK = 150
DIFFUSION_STEPS = 1000
RESPACE = "15,15,8,6,6"
NUM_INFERENCE_STEPS = 50
CONDS = ("xf_text", "xf_audio", "xf_spk")

# token-row layout (diffusion_architecture.py:146-149) and the two mask row sets
UPPER_ROWS = list(range(0, N_CHUNKS))
HANDS_ROWS = list(range(N_CHUNKS + 1, 2 * N_CHUNKS + 1))
FACE_ROWS = list(range(2 * N_CHUNKS + 2, 3 * N_CHUNKS + 2))
LOWER_ROWS = list(range(3 * N_CHUNKS + 3, N_TOKENS))
SEPARATOR_ROWS = [N_CHUNKS, 2 * N_CHUNKS + 1, 3 * N_CHUNKS + 2]            # {10, 21, 32}
QUERY_MASK_ZERO_ROWS = [(N_TOKENS - 3) // 4, 2 * (N_TOKENS - 3) // 4, 3 * (N_TOKENS - 3) // 4]  # {10, 20, 30}


def denoiser_cfg(retrieval_cfg=None):
    """The `model=dict(type="ReGestureTransformer", ...)` block exactly as the config spells it.
    `database=` and `use_retrieval_for_test=` arrive as kwargs through build_architecture /
    build_submodule (diffusion_architecture.py:104, visualize.py:135-138), not from this dict."""
    d = LATENT_DIM
    return dict(
        type="ReGestureTransformer",
        input_feats=INPUT_FEATS,
        max_seq_len=MAX_SEQ_LEN,
        frame_chunk_size=FRAME_CHUNK,
        latent_dim=d,
        time_embed_dim=TIME_EMBED_DIM,
        num_layers=NUM_LAYERS,
        body_part_cat_axis="time",
        sa_block_cfg=dict(type="EfficientSelfAttention", latent_dim=d, num_heads=NUM_HEADS,
                          dropout=0, time_embed_dim=TIME_EMBED_DIM),
        ca_block_cfg=dict(type="EfficientCrossAttention", latent_dim=d, text_latent_dim=d,
                          num_heads=NUM_HEADS, dropout=0, time_embed_dim=TIME_EMBED_DIM),
        ffn_cfg=dict(latent_dim=d, ffn_dim=FF_SIZE, dropout=0, time_embed_dim=TIME_EMBED_DIM),
        vae_cfg=dict(latent_dim=d, frame_chunk_size=FRAME_CHUNK),
        text_encoder=dict(pretrained_model=None, latent_dim=TEXT_DIM, num_layers=0, ff_size=2048,
                          dropout=0, use_text_proj=False),
        audio_encoder=dict(pretrained_model=None, latent_dim=TEXT_DIM, num_layers=0, dropout=0.1),
        speaker_embedding=dict(num_speakers=NUM_SPEAKERS),
        retrieval_train=False,
        retrieval_cfg=copy.deepcopy(retrieval_cfg),
        scale_func_cfg=None,
    )


def retrieval_cfg():
    """retrieval_cfg of the shipped config (config:96-133), minus the on-disk LMDB options."""
    return dict(num_retrieval=1, topk=2, latent_dim=LATENT_DIM, text_latent_dim=TEXT_DIM,
                max_seq_len=MAX_SEQ_LEN, motion_fps=MOTION_FPS, motion_framechunksize=FRAME_CHUNK)


def diffusion_test_cfg():
    return dict(beta_scheduler="scaled_linear", diffusion_steps=DIFFUSION_STEPS,
                model_mean_type="start_x", model_var_type="fixed_large", respace=RESPACE,
                num_inference_timesteps=NUM_INFERENCE_STEPS, classifier_free_guidance_scale=0)


def diffusion_train_cfg():
    return dict(beta_scheduler="scaled_linear", diffusion_steps=DIFFUSION_STEPS,
                model_mean_type="start_x", model_var_type="fixed_large")


def model_cfg():
    """The `model = dict(type="MotionDiffusion", ...)` block of the shipped config."""
    return dict(
        type="MotionDiffusion",
        model=denoiser_cfg(retrieval_cfg=retrieval_cfg()),
        loss_recon=dict(type="MSELoss", loss_weight=1, reduction="none"),
        body_part_lossweights=dict(upper=1.0, hands=1.0, face=1.0, lowertransl=1.0),
        diffusion_train=diffusion_train_cfg(),
        diffusion_test=diffusion_test_cfg(),
        inference_type="ddim",
    )
