"""Latent codec seam (GestureRepEncoder, diffusion_transformer.py:130-330).

The 4 body-part TransformerVAEs (rag_gesture_b200/vae.py) are an ADJACENT component: frozen,
run twice per batch and once per exemplar batch; their hyper-parameter YAMLs + checkpoints are not in the
reference repo.  The hot path only needs an object with `.encode(...) -> (latents [B,43,512], mask [B,43])`,
`.decode(latents) -> 7 pose tensors`, `.vae_latent_dim`, `.body_part_cat_axis`, `.frame_chunk_size`.
`SyntheticGestureCodec` is such an object with fixed random linear maps per 15-frame chunk; it is
what the golden pipeline fixtures, the tests and bench.py use on BOTH sides (it is injected into the
unmodified reference in tests/golden/make_golden.py), and it draws from the global RNG in the same
order and shapes as the real VAEs' rsample.  Plain PyTorch, any device.
"""
import torch
import torch.nn as nn

# (name, feature width per frame) in the order GestureRepEncoder.encode samples them
PARTS = (("upper", 39), ("hands", 90), ("face", 3 + 100), ("lowertrans", 27 + 3 + 4))


class SyntheticGestureCodec(nn.Module):
    def __init__(self, vae_cfg, body_part_cat_axis="time", seed=7):
        super().__init__()
        self.vae_cfg = vae_cfg
        self.body_part_cat_axis = body_part_cat_axis
        self.frame_chunk_size = vae_cfg["frame_chunk_size"]
        self.vae_latent_dim = vae_cfg["latent_dim"]
        g = torch.Generator().manual_seed(seed)
        for name, w in PARTS:
            k = self.frame_chunk_size * w
            self.register_buffer(f"{name}_enc", torch.randn(k, self.vae_latent_dim, generator=g) / k ** 0.5)
            self.register_buffer(f"{name}_dec", torch.randn(self.vae_latent_dim, k, generator=g) / self.vae_latent_dim ** 0.5)

    def _enc(self, name, feats):
        B, n, w = feats.shape
        c = self.frame_chunk_size
        mu = feats.reshape(B, n // c, c * w) @ getattr(self, f"{name}_enc")
        eps = torch.randn(B * (n // c), 1, self.vae_latent_dim).to(mu.device)      # rsample draw, gesture_vae.py:190
        return mu + 0.05 * eps.reshape(B, n // c, self.vae_latent_dim)

    def encode(self, motion_upper, motion_lower, motion_face, motion_hands, motion_transl,
               motion_facial, motion_contact, motion_mask):
        z_upper = self._enc("upper", motion_upper)
        z_hands = self._enc("hands", motion_hands)
        z_face = self._enc("face", torch.cat([motion_face, motion_facial], dim=-1))
        z_lt = self._enc("lowertrans", torch.cat([motion_lower, motion_transl, motion_contact], dim=-1))
        sep = torch.zeros_like(z_upper[:, :1, :])
        motion = torch.cat([z_upper, sep, z_hands, sep, z_face, sep, z_lt], dim=1)
        m = motion_mask[:, ::self.frame_chunk_size]
        ms = torch.zeros_like(m[:, :1])
        return motion, torch.cat([m, ms, m, ms, m, ms, m], dim=1)

    def encode_many(self, motion_upper, motion_lower, motion_face, motion_hands, motion_transl,
                    motion_facial, motion_contact, motion_mask):
        """E exemplars in one pass; the Gaussian draws are made exemplar by exemplar, part by part, i.e.
        exactly the sequence E separate encode() calls at B=1 would consume."""
        E, F, D, c = motion_upper.shape[0], motion_upper.shape[1], self.vae_latent_dim, self.frame_chunk_size
        eps = torch.stack([torch.stack([torch.randn(F // c, 1, D) for _ in PARTS], 0) for _ in range(E)], 0)
        eps = eps.to(motion_upper.device)                       # [E, 4, n, 1, D]
        feats = (motion_upper, motion_hands, torch.cat([motion_face, motion_facial], dim=-1),
                 torch.cat([motion_lower, motion_transl, motion_contact], dim=-1))
        zs = []
        for p, ((name, w), f) in enumerate(zip(PARTS, feats)):
            mu = f.reshape(E, F // c, c * w) @ getattr(self, f"{name}_enc")
            zs.append(mu + 0.05 * eps[:, p, :, 0, :])
        sep = torch.zeros_like(zs[0][:, :1, :])
        motion = torch.cat([zs[0], sep, zs[1], sep, zs[2], sep, zs[3]], dim=1)
        m = motion_mask[:, ::c]
        ms = torch.zeros_like(m[:, :1])
        return motion, torch.cat([m, ms, m, ms, m, ms, m], dim=1)

    def _dec(self, name, z, w):
        B, n, _ = z.shape
        return (z @ getattr(self, f"{name}_dec")).reshape(B, n * self.frame_chunk_size, w)

    def decode(self, z_output):
        n = (z_output.shape[1] - 3) // 4
        up = self._dec("upper", z_output[:, :n], 39)
        hands = self._dec("hands", z_output[:, n + 1:2 * n + 1], 90)
        face = self._dec("face", z_output[:, 2 * n + 2:3 * n + 2], 103)
        lt = self._dec("lowertrans", z_output[:, 3 * n + 3:], 34)
        return up, lt[..., :27], face[..., :3], hands, lt[..., 27:30], face[..., 3:], lt[..., 30:]


def build_codec(vae_cfg, body_part_cat_axis="time"):
    """vae_cfg with `upper_cfg / hands_cfg / face_cfg / lowertrans_cfg` YAML paths (the reference's config,
    basegesture_len150_beat.py:78-81) -> the four TransformerVAEs (rag_gesture_b200.vae.GestureRepEncoder,
    checkpoints loaded); without them -> the synthetic stand-in used by the fixtures and bench.py."""
    if any(k in vae_cfg for k in ("upper_cfg", "hands_cfg", "face_cfg", "lowertrans_cfg")):
        from .vae import GestureRepEncoder
        return GestureRepEncoder(vae_cfg, body_part_cat_axis)
    return SyntheticGestureCodec(vae_cfg, body_part_cat_axis)
