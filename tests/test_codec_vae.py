"""The latent codec (rag_gesture_b200/vae.py: TransformerVAE + GestureRepEncoder) against outputs of the
UNMODIFIED reference classes (tests/golden/make_golden.py group "codec": gesture_vae.py / diffusion_transformer.py
run on the same synthetic weights, inputs regenerated from seeds).  CPU: the codec is PyTorch host code."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from rag_gesture_b200 import config as C
from rag_gesture_b200 import synthetic as S
from rag_gesture_b200.vae import GestureRepEncoder, TransformerVAE

_spec = importlib.util.spec_from_file_location("make_golden_codec", os.path.join(GOLDEN, "make_golden.py"))
TOL = 2e-5


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "codec_vae.npz"))


@pytest.fixture(scope="module")
def mg():
    """VAE_VARIANTS / codec_inputs / write_vae_files of the generator (refshim.load() is NOT called)."""
    m = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(m)
    return m


def _vae(mg, name):
    vae = TransformerVAE(S.vae_args("upper", **mg.VAE_VARIANTS[name])).eval()
    shapes = {k: tuple(v.shape) for k, v in vae.state_dict().items()}
    vae.load_state_dict(S.synthetic_vae_state_dict(shapes, 100))
    return vae, shapes


@pytest.mark.parametrize("name", ["a", "b"])
def test_vae_state_dict_keys_match_reference(gold, mg, name):
    _, shapes = _vae(mg, name)
    ref = {k: tuple(v) for k, v in json.loads(str(gold[f"{name}_keys"])).items()}
    assert shapes == ref                      # same keys AND shapes: the reference's VAE checkpoints load


@pytest.mark.parametrize("name", ["a", "b"])
def test_vae_encode_decode_vs_reference(gold, mg, name):
    """all_encoder/learned/normal/post-norm/gelu and encoder_decoder/sine/multivariate_normal/pre-norm/relu;
    full-length and ragged (lengths 150/120/45) batches; same global-RNG draw as Distribution.rsample()."""
    vae, _ = _vae(mg, name)
    x = 0.5 * torch.randn(3, C.MAX_SEQ_LEN, 78, generator=torch.Generator().manual_seed(5))
    assert mg.digest(x) == str(gold[f"{name}_in_digest"])
    with torch.no_grad():
        torch.manual_seed(21)
        z, dist = vae.encode_to_dist(x)
        rec = vae.decode(z)
        torch.manual_seed(22)
        z_r, _ = vae.encode_to_dist(x, [150, 120, 45])
        rec_r = vae.decode(z_r, [150, 120, 45])
    assert tuple(z.shape) == (3, 10, 64) and tuple(rec.shape) == (3, 150, 78) and tuple(dist.loc.shape) == (30, 1, 64)
    assert rel_l2(z, torch.from_numpy(gold[f"{name}_z"])) < TOL
    assert rel_l2(rec[:, ::5], torch.from_numpy(gold[f"{name}_rec"])) < TOL
    assert rel_l2(z_r, torch.from_numpy(gold[f"{name}_z_ragged"])) < TOL
    assert rel_l2(rec_r[:, ::5], torch.from_numpy(gold[f"{name}_rec_ragged"])) < TOL
    assert bool((rec_r[1, 120:] == 0).all()) and bool((rec_r[2, 45:] == 0).all())      # zero beyond each length


def test_gesture_rep_encoder_vs_reference(gold, mg, tmp_path):
    """YAML + checkpoint loading (incl. the 'module.' prefix), axis-angle -> 6D -> 4 VAEs -> separator layout,
    decode back to axis-angle, the in-place translation rebase, and encode_many == E encodes at B=1 (same draws)."""
    def shapes_of(args):
        return {k: tuple(v.shape) for k, v in TransformerVAE(args).state_dict().items()}
    cfg = mg.write_vae_files(str(tmp_path), "a", 200, shapes_of)
    from rag_gesture_b200.codec import build_codec
    enc = build_codec(cfg, "time")
    assert isinstance(enc, GestureRepEncoder) and enc.vae_latent_dim == 64 and enc.frame_chunk_size == 15
    assert not any(p.requires_grad for p in enc.parameters())
    inp = mg.codec_inputs(2, 31)
    assert mg.digest(*inp.values()) == str(gold["enc_in_digest"])
    with pytest.raises(RuntimeError):
        enc.decode(torch.zeros(1, 43, 64))                      # joint counts come from encode, as in the reference
    torch.manual_seed(41)
    args = {k: v.clone() for k, v in inp.items()}
    motion, mask = enc.encode(**args)
    assert tuple(motion.shape) == (2, 43, 64) and bool((motion[:, [10, 21, 32]] == 0).all())
    assert rel_l2(motion, torch.from_numpy(gold["enc_motion"])) < TOL
    assert torch.equal(mask, torch.from_numpy(gold["enc_mask"]))
    assert bool((args["motion_transl"][:, 0, [0, 2]] == 0).all())          # rebased in place (quirk kept)
    dec = enc.decode(motion)
    for k, v in zip(("upper", "lower", "face", "hands", "transl", "exps", "contact"), dec):
        g = torch.from_numpy(gold[f"dec_{k}"])
        assert tuple(v[:, ::10].shape) == tuple(g.shape), k
        assert rel_l2(v[:, ::10], g) < 1e-4, k                   # incl. the 6D -> axis-angle conversion
    torch.manual_seed(42)
    many, _ = enc.encode_many(**{k: v.clone() for k, v in inp.items()})
    assert rel_l2(many, torch.from_numpy(gold["enc_single"])) < TOL


def test_encode_with_predrawn_noise_equals_encode(mg, tmp_path):
    """MotionDiffusion.prepare takes the codec's Gaussian draws first (draw_encode_eps, at encode's place in the random
    stream) and runs the encode pass after the retrieval stage: encode(eps=draw) must be encode() for the same
    generator state, with the default generator and with an installed one, and leave the stream where encode() does."""
    def shapes_of(args):
        return {k: tuple(v.shape) for k, v in TransformerVAE(args).state_dict().items()}
    enc = GestureRepEncoder(mg.write_vae_files(str(tmp_path), "b", 200, shapes_of), "time").eval()
    inp = mg.codec_inputs(3, 77)
    torch.manual_seed(11)
    a, am = enc.encode(**{k: v.clone() for k, v in inp.items()})
    after_a = torch.rand(1)
    torch.manual_seed(11)
    eps = enc.draw_encode_eps(inp["motion_upper"])
    after_b = torch.rand(1)                              # something else draws between the two halves
    b, bm = enc.encode(**{k: v.clone() for k, v in inp.items()}, eps=eps)
    assert torch.equal(a, b) and torch.equal(am, bm) and torch.equal(after_a, after_b)
    enc.generator = torch.Generator().manual_seed(5)
    c, _ = enc.encode(**{k: v.clone() for k, v in inp.items()})
    enc.generator = torch.Generator().manual_seed(5)
    d, _ = enc.encode(**{k: v.clone() for k, v in inp.items()}, eps=enc.draw_encode_eps(inp["motion_upper"]))
    enc.generator = None
    assert torch.equal(c, d) and not torch.equal(a, c)
