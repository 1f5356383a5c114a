"""The TransformerVAE codec on the device and inside MotionDiffusion.forward (vae_cfg with YAML paths, as the
reference's config has them).  Needs a GPU."""
import importlib.util
import os

import pytest
import torch

from conftest import GOLDEN, rel_l2
from rag_gesture_b200 import config as C
from rag_gesture_b200 import synthetic as S
from rag_gesture_b200.vae import GestureRepEncoder, TransformerVAE

pytestmark = pytest.mark.gpu
_spec = importlib.util.spec_from_file_location("make_golden_codec", os.path.join(GOLDEN, "make_golden.py"))


@pytest.fixture(scope="module")
def mg():
    m = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(m)
    return m


def test_vae_device_matches_host(mg):
    """Same weights, inputs and eps: the fused-attention path on CUDA == the CPU path pinned to the reference."""
    dev = torch.device("cuda:0")
    for name in ("a", "b"):
        vae = TransformerVAE(S.vae_args("hands", **mg.VAE_VARIANTS[name])).eval()
        vae.load_state_dict(S.synthetic_vae_state_dict({k: tuple(v.shape) for k, v in vae.state_dict().items()}, 7))
        g = torch.Generator().manual_seed(9)
        x, eps = 0.5 * torch.randn(4, 150, 180, generator=g), torch.randn(40, 1, 64, generator=g)
        with torch.no_grad():
            z, _ = vae.encode_to_dist(x, [150, 150, 90, 30], eps=eps)
            rec = vae.decode(z, [150, 150, 90, 30])
            vd = vae.to(dev)
            zd, _ = vd.encode_to_dist(x.to(dev), [150, 150, 90, 30], eps=eps.to(dev))
            recd = vd.decode(zd, [150, 150, 90, 30])
        assert rel_l2(zd.cpu(), z) < 1e-4 and rel_l2(recd.cpu(), rec) < 1e-4, name


def test_vae_tensor_core_tiers_match_fp32(mg, tmp_path):
    """F1: the four VAEs with their 512/1024-wide projections on the library's tcgen05 GEMM (set_gemm_tier) against
    the same codec on cuBLAS fp32 -- same inputs, same Gaussian draws.  bf16x3 is held to the fp32 tier's tolerance,
    bf16 to north_star's 2e-2; the golden-pinned CPU path is what the fp32 one is held to above."""
    from rag_gesture_b200 import _lib
    dev = torch.device("cuda:0")

    def shapes_of(args):
        return {k: tuple(v.shape) for k, v in TransformerVAE(args).state_dict().items()}
    for variant in ("a", "b"):
        cfg = mg.write_vae_files(os.path.join(str(tmp_path), variant), variant, 200, shapes_of, latent_dim=C.LATENT_DIM)
        enc = GestureRepEncoder(cfg, "time").to(dev).eval()
        inp = {k: v.to(dev) for k, v in mg.codec_inputs(5, 31).items()}
        outs = {}
        for tier in (None, "bf16x3", "bf16"):
            enc.set_gemm_tier(tier)
            enc.generator = torch.Generator(device=dev).manual_seed(77)
            n0 = _lib.load().rg_launch_count()
            motion, mask = enc.encode(**{k: v.clone() for k, v in inp.items()})
            many, _ = enc.encode_many(**{k: v.clone() for k, v in inp.items()})
            dec = enc.decode(motion)
            n = C.N_CHUNKS
            with torch.no_grad():                               # 6D rotations straight from the part decoders: the
                raw = (enc.upper_vae.decode(motion[:, :n]),     # axis-angle outputs jump at pi (compared below only
                       enc.hands_vae.decode(motion[:, n + 1:2 * n + 1]))     # through their linear companions)
            outs[tier] = (motion, many, dec[4], dec[5], dec[6]) + raw
            assert (_lib.load().rg_launch_count() > n0) == (tier is not None)       # the tcgen05 GEMM did (not) run
        enc.generator = None
        for tier, tol in (("bf16x3", 1e-4), ("bf16", 2e-2)):
            errs = [rel_l2(a, b) for a, b in zip(outs[tier], outs[None])]
            print(f"VAE variant {variant}, GEMM tier {tier}: max rel-L2 vs fp32 {max(errs):.3g}")
            assert max(errs) < tol, (variant, tier, errs)


def test_codec_graph_replay_equals_eager(mg, tmp_path):
    """enable_graphs(): encode / encode_many (padded to the exemplar bucket) / decode replayed as CUDA graphs give the
    eager pass's results for the same generator state, across replays with different inputs, in the cuBLAS tier and
    in the tensor-core tier (library calls with stream-ordered allocations inside the capture); a weight update drops
    the captured passes."""
    dev = torch.device("cuda:0")

    def shapes_of(args):
        return {k: tuple(v.shape) for k, v in TransformerVAE(args).state_dict().items()}
    cfg = mg.write_vae_files(str(tmp_path), "a", 200, shapes_of, latent_dim=C.LATENT_DIM)
    enc = GestureRepEncoder(cfg, "time").to(dev).eval()
    for tier in (None, "bf16x3"):
        enc.set_gemm_tier(tier)
        enc.enable_graphs(True)
        for rnd in range(3):                                # round 0 captures, 1 and 2 replay with other inputs
            clips = {k: v.to(dev) for k, v in mg.codec_inputs(4, 50 + rnd).items()}
            ex = {k: v.to(dev) for k, v in mg.codec_inputs(5, 60 + rnd).items()}
            res = {}
            for graphs in (False, True):
                enc.use_graphs = graphs
                enc.generator = torch.Generator(device=dev).manual_seed(5 + rnd)
                a = {k: v.clone() for k, v in clips.items()}
                b = {k: v.clone() for k, v in ex.items()}
                motion, mask = enc.encode(**a)
                many, mmask = enc.encode_many(**b)
                dec = enc.decode(motion)
                res[graphs] = (motion, mask, many, mmask, a["motion_transl"], b["motion_transl"], dec[4], dec[5], dec[6])
            for i, (x, y) in enumerate(zip(res[False], res[True])):
                assert x.shape == y.shape, (tier, rnd, i)
                assert torch.allclose(x, y, rtol=1e-5, atol=1e-5), (tier, rnd, i, float((x - y).abs().max()))
        n_graphs = len([k for k in enc.__dict__["_rg_graphs"] if k != "stamp"])
        assert n_graphs == 3, n_graphs                      # encode B=4, encode 16 (5 padded), decode B=4
        with torch.no_grad():
            enc.face_vae.final_layer.bias.add_(1.0)        # in-place update: version bump -> passes re-captured
        a = {k: v.clone() for k, v in clips.items()}
        enc.generator = torch.Generator(device=dev).manual_seed(9)
        enc.use_graphs = True
        m_g = enc.encode(**a)[0]
        d_g = enc.decode(m_g)
        enc.use_graphs = False
        d_e = enc.decode(m_g)
        assert torch.allclose(d_g[5], d_e[5], atol=1e-5) and torch.allclose(d_g[4], d_e[4], atol=1e-5)
        with torch.no_grad():
            enc.face_vae.final_layer.bias.sub_(1.0)
    enc.generator = None


def test_forward_with_transformer_vae_codec(mg, tmp_path):
    """build_architecture with the reference-style vae_cfg (four YAML + checkpoint pairs): encode -> plain DDIM
    on the CUDA path -> decode; deterministic under fixed seeds, finite, reference output shapes."""
    import rag_gesture_b200 as R
    dev = torch.device("cuda:0")

    def shapes_of(args):
        return {k: tuple(v.shape) for k, v in TransformerVAE(args).state_dict().items()}
    vae_cfg = mg.write_vae_files(str(tmp_path), "a", 300, shapes_of, latent_dim=C.LATENT_DIM)
    cfg = C.model_cfg()
    cfg["model"]["vae_cfg"] = vae_cfg
    arch = R.build_architecture(cfg, database=None)
    assert isinstance(arch.model.gesture_rep_encoder, GestureRepEncoder)
    missing, unexpected = arch.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
    assert not unexpected and all(k.startswith("gesture_rep_encoder.") for k in missing)
    arch = arch.to(dev).eval()
    qs = S.SyntheticGestureDataset(4, seed=8)
    outs = []
    for _ in range(2):
        batch = S.collate([qs[i] for i in (0, 2)])
        batch["inference_kwargs"] = {}
        torch.manual_seed(3)
        torch.cuda.manual_seed(4)
        res = arch(**batch)
        outs.append({k: res[k].cpu() for k in ("prev_latentout", "pred_upper", "pred_lower", "pred_facepose",
                                               "pred_hands", "pred_transl", "pred_exps")})
    a, b = outs
    assert tuple(a["prev_latentout"].shape) == (2, 43, 512) and tuple(a["pred_upper"].shape) == (2, 150, 39)
    assert tuple(a["pred_lower"].shape) == (2, 150, 27) and tuple(a["pred_hands"].shape) == (2, 150, 90)
    assert tuple(a["pred_facepose"].shape) == (2, 150, 3) and tuple(a["pred_exps"].shape) == (2, 150, 100)
    assert tuple(a["pred_transl"].shape) == (2, 150, 3)
    for k in a:
        assert bool(torch.isfinite(a[k]).all()), k
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("tier", ["fp32", "bf16x3"])
def test_pipeline_with_vae_codec_is_reproducible_and_matches_sequential(mg, tmp_path, tier):
    """the VAE codec draws its rsample noise on the device.  GuidedPipeline gives it its own generator, so
    (a) two pipeline runs with the same seeds agree bit for bit (no dependence on thread timing) and (b) they equal
    sequential forward() calls made with the same codec generator installed.  In the tensor-core tier the pipeline
    replays the codec's passes as CUDA graphs (worker thread: encode, main thread: decode) while forward() runs them
    eagerly: (b) then holds to rounding of the few cuBLAS calls left in the passes."""
    import rag_gesture_b200 as R
    from rag_gesture_b200 import _lib
    from rag_gesture_b200.architecture import GuidedPipeline
    dev = torch.device("cuda:0")

    def shapes_of(args):
        return {k: tuple(v.shape) for k, v in TransformerVAE(args).state_dict().items()}
    cfg = C.model_cfg()
    cfg["model"]["vae_cfg"] = mg.write_vae_files(str(tmp_path), "a", 300, shapes_of, latent_dim=C.LATENT_DIM)
    cfg["model"]["precision"] = {"fp32": _lib.PREC_FP32, "bf16x3": _lib.PREC_BF16X3}[tier]
    arch = R.build_architecture(cfg, database=None)
    arch.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
    arch = arch.to(dev).eval()
    codec = arch.model.gesture_rep_encoder
    assert codec.draws_on_device and codec.generator is None
    assert codec.gemm_tier == (None if tier == "fp32" else "bf16x3") and GuidedPipeline(arch).codec_graphs == (tier != "fp32")
    qs = S.SyntheticGestureDataset(8, seed=8)

    def batches():
        for ids in ([0, 2], [1], [3, 4, 5]):
            b = S.collate([qs[i] for i in ids])
            b["inference_kwargs"] = {}
            yield b

    def seed():
        torch.manual_seed(3)
        torch.cuda.manual_seed(4)

    runs = []
    for _ in range(2):
        seed()
        runs.append([r["prev_latentout"].cpu() for r in GuidedPipeline(arch).run(batches())])
        assert codec.generator is None and codec.use_graphs is False      # restored after run()
    assert all(torch.equal(a, b) for a, b in zip(*runs))
    seed()
    codec.generator = GuidedPipeline.codec_generator(dev)
    try:
        seq = [arch(**b)["prev_latentout"].cpu() for b in batches()]
    finally:
        codec.generator = None
    assert len(seq) == 3
    if tier == "fp32":
        assert all(torch.equal(a, b) for a, b in zip(seq, runs[0]))
    else:
        assert len([k for k in codec.__dict__["_rg_graphs"] if k != "stamp"]) == 6       # encode + decode at B = 2, 1, 3
        assert max(rel_l2(a, b) for a, b in zip(seq, runs[0])) < 1e-4
