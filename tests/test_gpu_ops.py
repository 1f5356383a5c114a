"""Kernel-level parity: each rg_op_* entry point (through the C ABI) against the oracle's
restatement of the reference op on the same seeded inputs.  Needs a GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import denoiser as OD
from oracle import diffusion as ODF

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _g(seed):
    return torch.Generator().manual_seed(seed)


def test_library_loaded_and_counts_launches(dev):
    from rag_gesture_b200 import _lib, ops
    lib = _lib.load()
    n0 = lib.rg_launch_count()
    ops.silu(torch.ones(8, device=dev))
    assert lib.rg_launch_count() == n0 + 1


@pytest.mark.parametrize("M,N,K", [(1, 512, 512), (43, 1536, 512), (300, 512, 1536), (129, 1024, 768), (2752, 512, 1024)])
def test_linear_epilogues(dev, M, N, K):
    from rag_gesture_b200 import _lib, ops
    g = _g(M + N + K)
    x, w, b, r = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    ref = F.linear(x.double(), w.double(), b.double())
    out = ops.linear(x.to(dev), w.to(dev), b.to(dev)).cpu()
    assert rel_l2(out, ref) < 2e-6
    assert rel_l2(ops.linear(x.to(dev), w.to(dev), b.to(dev), residual=r.to(dev)).cpu(), ref + r.double()) < 2e-6
    assert rel_l2(ops.linear(x.to(dev), w.to(dev), b.to(dev), epilogue=_lib.OP_GELU).cpu(), F.gelu(ref)) < 2e-6
    assert rel_l2(ops.linear(x.to(dev), w.to(dev), b.to(dev), epilogue=_lib.OP_SILU).cpu(), F.silu(ref)) < 2e-6
    assert rel_l2(ops.linear(x.to(dev), w.to(dev), None).cpu(), F.linear(x.double(), w.double())) < 2e-6


def test_layernorm_and_stylization_rows(dev):
    from rag_gesture_b200 import ops
    g = _g(3)
    B, T, D = 3, 43, 512
    y = torch.randn(B, T, D, generator=g) * 3 + 1
    gam, bet = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    ref = F.layer_norm(y, (D,), gam, bet, 1e-5)
    assert rel_l2(ops.layernorm(y.to(dev), gam.to(dev), bet.to(dev)).cpu(), ref) < 2e-6
    assert rel_l2(ops.layernorm(y.to(dev)).cpu(), F.layer_norm(y, (D,))) < 2e-6
    ss = 0.3 * torch.randn(B, 2 * D, generator=g)
    scale, shift = ss[:, None, :D], ss[:, None, D:]
    ref = F.silu(F.layer_norm(y, (D,), gam, bet, 1e-5) * (1 + scale) + shift)
    out = ops.stylization_rows(y.view(B * T, D).to(dev), gam.to(dev), bet.to(dev), ss.to(dev), T).cpu().view(B, T, D)
    assert rel_l2(out, ref) < 2e-6
    # -1e6 rows (efficient_attention.py:98): the row collapses to a constant -> LN returns beta
    y2 = 0.004 * torch.randn(2, D, generator=g) + -1000000.0      # |y| < 1/32: every entry rounds to -1e6
    out2 = ops.layernorm(y2.to(dev), gam.to(dev), bet.to(dev)).cpu()
    assert torch.equal(out2[0], bet) and torch.equal(out2[1], bet)
    assert torch.allclose(out2, F.layer_norm(y2, (D,), gam, bet, 1e-5), atol=1e-6, rtol=0)


def test_self_attention_block(dev, sd0):
    """EfficientSelfAttention.forward of the mirror module vs the oracle (efficient_attention.py:23-45)."""
    from rag_gesture_b200 import mogen_api as M
    from rag_gesture_b200 import synthetic as S
    p = "temporal_decoder_blocks.2.sa_block"
    blk = M.EfficientSelfAttention(512, 16, 0, 2048)
    blk.load_state_dict({k[len(p) + 1:]: v for k, v in sd0.items() if k.startswith(p + ".")})
    blk = blk.to(dev).eval()
    g = _g(5)
    B = 3
    x = torch.randn(B, 43, 512, generator=g)
    emb = torch.randn(B, 2048, generator=g)
    mask = S.motion_mask(B)
    mask[1, 5] = 0
    ref = OD.efficient_self_attention(sd0, p, x, mask.unsqueeze(-1), emb, 16)
    with torch.no_grad():
        out = blk(x=x.to(dev), src_mask=mask.unsqueeze(-1).to(dev), emb=emb.to(dev)).cpu()
    assert rel_l2(out, ref) < 5e-6
    # all tokens masked: the -1e6 shift is applied to every key; still finite and equal
    mask0 = torch.zeros(B, 43)
    ref0 = OD.efficient_self_attention(sd0, p, x, mask0.unsqueeze(-1), emb, 16)
    with torch.no_grad():
        out0 = blk(x=x.to(dev), src_mask=mask0.unsqueeze(-1).to(dev), emb=emb.to(dev)).cpu()
    assert rel_l2(out0, ref0) < 5e-6


def test_cross_attention_block_and_state(dev, sd0):
    from rag_gesture_b200 import mogen_api as M
    from rag_gesture_b200 import synthetic as S
    for cond, N in (("xf_text", 150), ("xf_audio", 499)):
        p = f"temporal_decoder_blocks.1.ca_blocks.{cond}"
        blk = M.EfficientCrossAttention(512, 512, 16, 0, 2048)
        blk.load_state_dict({k[len(p) + 1:]: v for k, v in sd0.items() if k.startswith(p + ".")})
        blk = blk.to(dev).eval()
        g = _g(N)
        B = 2
        x, xf = torch.randn(B, 43, 512, generator=g), torch.randn(B, N, 512, generator=g)
        emb = torch.randn(B, 2048, generator=g)
        qm = S.query_masks(B)[cond]
        taps = {}
        ref = OD.efficient_cross_attention(sd0, p, x, xf, emb, qm, torch.ones(B, 1, 1), 16, taps)
        with torch.no_grad():
            st = blk.kv_state(xf.to(dev)).cpu()
            out = blk(x=x.to(dev), xf=xf.to(dev), emb=emb.to(dev), query_mask=qm.to(dev),
                      cond_type=torch.ones(B, 1, 1, device=dev)).cpu()
            out_nomask = blk(x=x.to(dev), xf=xf.to(dev), emb=emb.to(dev), query_mask=None).cpu()
        assert rel_l2(st, taps["ca_state"][0]) < 5e-6
        assert rel_l2(out, ref) < 5e-6
        ref_nm = OD.efficient_cross_attention(sd0, p, x, xf, emb, None, None, 16)
        assert rel_l2(out_nomask, ref_nm) < 5e-6


def test_ffn_and_decoder_layer(dev, sd0):
    from rag_gesture_b200 import config as C
    from rag_gesture_b200 import mogen_api as M
    from rag_gesture_b200 import synthetic as S
    cfg = C.denoiser_cfg()
    p = "temporal_decoder_blocks.0"
    layer = M.DecoderLayer(cfg["sa_block_cfg"], cfg["ca_block_cfg"], cfg["ffn_cfg"])
    layer.load_state_dict({k[len(p) + 1:]: v for k, v in sd0.items() if k.startswith(p + ".")})
    layer = layer.to(dev).eval()
    g = _g(9)
    B = 2
    x = torch.randn(B, 43, 512, generator=g)
    emb = torch.randn(B, 2048, generator=g)
    xf = {"xf_text": torch.randn(B, 150, 512, generator=g), "xf_audio": torch.randn(B, 499, 512, generator=g),
          "xf_spk": torch.randn(B, 150, 512, generator=g)}
    mask, qm = S.motion_mask(B), S.query_masks(B)
    ref = OD.decoder_layer(sd0, p, x, xf, emb, mask.unsqueeze(-1), qm, torch.ones(B, 1, 1), 16)
    with torch.no_grad():
        out = layer(x=x.to(dev), xf={k: v.to(dev) for k, v in xf.items()}, emb=emb.to(dev),
                    src_mask=mask.unsqueeze(-1).to(dev), query_mask={k: v.to(dev) for k, v in qm.items()},
                    cond_type=None).cpu()
    assert rel_l2(out, ref) < 1e-5


def test_ddim_update_and_blend_bit_exact(dev, sd0):
    """K8/K9 reproduce the reference's fp32 op order bit for bit (no FMA contraction)."""
    from rag_gesture_b200 import config as C
    from rag_gesture_b200.diffusion import build_diffusion
    from rag_gesture_b200.engine import DenoiserEngine
    diff = build_diffusion(C.diffusion_test_cfg())
    eng = DenoiserEngine(sd0, device=dev)
    eng.set_schedule(diff.timestep_map, diff.coef_table())
    od = ODF.OracleDiffusion()
    g = _g(17)
    B, T, D = 3, 43, 512
    x, x0 = torch.randn(B, T, D, generator=g), torch.randn(B, T, D, generator=g)
    for i in (0, 1, 24, 48, 49):
        t = torch.tensor([i] * B)
        # forward update (gaussian_diffusion.py:983-1001) through the oracle's arithmetic
        eps = (ODF._ext(od.s.sqrt_recip_alphas_cumprod, t, x.shape) * x - x0) / ODF._ext(od.s.sqrt_recipm1_alphas_cumprod, t, x.shape)
        abp = ODF._ext(od.s.alphas_cumprod_prev, t, x.shape)
        ref_f = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp) * eps
        abn = ODF._ext(od.s.alphas_cumprod_next, t, x.shape)
        ref_r = x0 * torch.sqrt(abn) + torch.sqrt(1 - abn) * eps
        assert torch.equal(eng.ddim_update(x.to(dev), x0.to(dev), i, -1).cpu(), ref_f)
        assert torch.equal(eng.ddim_update(x.to(dev), x0.to(dev), i, +1).cpu(), ref_r)
        # blend (:934-947): some rows of in_seq non-zero, one row with a single non-zero entry
        in_seq = torch.zeros(B, T, D)
        in_seq[0, 2:5] = torch.randn(3, D, generator=g)
        in_seq[2, 40, 7] = 1e-30
        noise = torch.randn(B, T, D, generator=g)
        od.randn = lambda shape, device="cpu": noise
        ref_b = od.blend_in_seq(x, in_seq, t)
        assert torch.equal(eng.blend_in_seq(x.to(dev), in_seq.to(dev), noise.to(dev), i).cpu(), ref_b)
    eng.close()


def test_guidance_step_closed_form(dev, sd0):
    from rag_gesture_b200.engine import DenoiserEngine
    eng = DenoiserEngine(sd0, device=dev)
    g = _g(23)
    B, T, D = 2, 43, 512
    x = torch.randn(B, T, D, generator=g)
    in_seq = torch.zeros(B, T, D)
    in_seq[1, 3:6] = torch.randn(3, D, generator=g)
    mask = (in_seq != 0).any(-1)
    lat = x.clone().requires_grad_(True)
    cur = lat
    for _ in range(5):
        loss = F.mse_loss(cur * mask.unsqueeze(-1).float(), in_seq)
        (gr,) = torch.autograd.grad(loss, [cur], retain_graph=True)
        cur = cur - 0.1 * gr
    out = eng.guidance_steps(x.to(dev).clone(), in_seq.to(dev), 5, 0.1).cpu()
    assert rel_l2(out, cur.detach()) < 1e-6
    assert torch.equal(out[~mask], x[~mask])
    eng.close()


def _sa_core_ref(qkv, mask):
    """float64 restatement of EfficientSelfAttention's core (efficient_attention.py:146-160)."""
    B, T, _ = qkv.shape
    q, k, v = [t.double().view(B, T, 16, 32) for t in qkv.split(512, dim=-1)]
    m = mask.double().view(B, T, 1, 1)
    k = torch.softmax(k + (1 - m) * -1000000.0, dim=1)
    q = torch.softmax(q, dim=-1)
    A = torch.einsum("bnhd,bnhl->bhdl", k, v * m)
    return torch.einsum("bnhd,bhdl->bnhl", q, A).reshape(B, T, 512)


@pytest.mark.parametrize("T", [3, 11, 23, 43, 49, 64])
def test_attention_cores_mma_vs_fp32(dev, T):
    """The mma.sync cores (TF32 and 3xTF32) against the fp32 cores and a float64 formula, for every
    tile count MT = ceil(T/16) and ragged last tiles; masked key tokens and masked query rows included."""
    from rag_gesture_b200 import ops
    B = 5
    g = torch.Generator().manual_seed(100 + T)
    qkv = torch.randn(B, T, 1536, generator=g).to(dev)
    mask = (torch.rand(B, T, generator=g) > 0.2).float()
    mask[0] = 1
    mask[:, 0] = 1
    mask = mask.to(dev)
    ref = _sa_core_ref(qkv, mask)
    y0 = ops.self_attention_core(qkv, mask, 0)
    for mode, tol in ((1, 2e-3), (2, 3e-6)):
        y = ops.self_attention_core(qkv, mask, mode)
        e, e0 = rel_l2(y.double(), ref), rel_l2(y0.double(), ref)
        print(f"T={T} self core mode {mode}: rel-L2 vs f64 {e:.2e} (fp32 core {e0:.2e})")
        assert e < tol and e0 < 3e-6
    q3 = torch.randn(B, T, 1536, generator=g).to(dev)
    state = (torch.randn(B, 3, 16, 32, 32, generator=g) * 0.1).to(dev)
    qm = (torch.rand(3, B, T, generator=g) > 0.1).float().to(dev)
    p = torch.softmax(q3.double().view(B, T, 3, 16, 32), dim=-1)
    ref = torch.einsum("bnchd,bchdl->bnchl", p, state.double()).reshape(B, T, 1536)
    for qmask in (None, qm):
        z0 = ops.cross_attention_core(q3, state, qmask, 0)
        for mode, tol in ((1, 2e-3), (2, 3e-6)):
            z = ops.cross_attention_core(q3, state, qmask, mode)
            if qmask is None:
                assert rel_l2(z.double(), ref) < tol and rel_l2(z0.double(), ref) < 3e-6
            else:
                keep = qmask.permute(1, 2, 0).reshape(B, T, 3, 1).expand(B, T, 3, 512).reshape(B, T, 1536) > 0
                assert rel_l2(z[keep].double(), ref[keep]) < tol
                # masked rows: y - 1e6 in fp32, a 1/16 grid -> identical to the fp32 core unless y sits on a tie
                assert (z[~keep] == z0[~keep]).float().mean() > 0.99 and (z[~keep] < -9e5).all()


@pytest.mark.parametrize("T", [11, 23, 43, 64])
def test_fused_attention_stylization_tc(dev, T):
    """sa_styl / ca_styl (mma.sync core + Stylization prologue in one kernel, LayerNorm statistics merged
    across the 16 head-warps) == core followed by the row kernel; shared and per-clip scale/shift."""
    from rag_gesture_b200 import ops
    B = 4
    g = torch.Generator().manual_seed(200 + T)
    qkv = torch.randn(B, T, 1536, generator=g).to(dev)
    mask = (torch.rand(B, T, generator=g) > 0.2).float()
    mask[:, 0] = 1
    mask = mask.to(dev)
    gamma = (1 + 0.1 * torch.randn(3, 512, generator=g)).to(dev)
    beta = (0.1 * torch.randn(3, 512, generator=g)).to(dev)
    for per_clip in (False, True):
        ss = (0.3 * torch.randn(*((B, 3, 1024) if per_clip else (3, 1024)), generator=g)).to(dev)
        ss0 = ss[:, 0].contiguous() if per_clip else ss[0].contiguous()
        for split, tol in ((False, 2e-3), (True, 5e-6)):
            ref = ops.stylization_rows(ops.self_attention_core(qkv, mask, 0).view(B * T, 512), gamma[0], beta[0], ss0, T)
            got = ops.self_attention_tc(qkv, mask, gamma[0], beta[0], ss0, split)
            assert rel_l2(got.view(B * T, 512), ref) < tol
        q3 = torch.randn(B, T, 1536, generator=g).to(dev)
        state = (torch.randn(B, 3, 16, 32, 32, generator=g) * 0.1).to(dev)
        qm = (torch.rand(3, B, T, generator=g) > 0.1).float().to(dev)
        y = ops.cross_attention_core(q3, state, qm, 0).view(B * T, 3, 512)
        ref = torch.stack([ops.stylization_rows(y[:, c].contiguous(), gamma[c], beta[c],
                                                ss[:, c].contiguous() if per_clip else ss[c].contiguous(), T)
                           for c in range(3)], 1).view(B, T, 1536)
        keep = qm.permute(1, 2, 0).reshape(B, T, 3, 1).expand(B, T, 3, 512).reshape(B, T, 1536) > 0
        for split, tol in ((False, 2e-3), (True, 5e-6)):
            got = ops.cross_attention_tc(q3, state, qm, gamma, beta, ss, split)
            assert rel_l2(got[keep], ref[keep]) < tol
            assert torch.isfinite(got).all()


@pytest.mark.parametrize("N,Sq,Sk,H,dh", [(640, 17, 17, 4, 128), (64, 160, 160, 4, 128), (7, 150, 10, 4, 128),
                                           (5, 33, 65, 8, 64), (3, 1, 200, 2, 32), (9, 40, 40, 16, 32),
                                           (64, 160, 160, 32, 16), (3, 45, 71, 4, 16)])
def test_mha_vs_torch_fp64(dev, N, Sq, Sk, H, dh):
    """rg_op_mha (the codec VAEs' softmax attention) against float64 softmax(q k^T / sqrt(dh)) v, with a key-padding
    mask, for q/k/v given as strided column blocks of one fused projection and as separate tensors."""
    from rag_gesture_b200 import ops
    g = torch.Generator().manual_seed(N * 1000 + Sq)
    D = H * dh
    keep = torch.rand(N, Sk, generator=g) > 0.3
    keep[:, 0] = True                                    # no fully masked row (the VAE's global tokens are always kept)

    def ref(q, k, v, keep):
        qh, kh, vh = (t.double().view(N, -1, H, dh).transpose(1, 2) for t in (q, k, v))
        s = qh @ kh.transpose(-1, -2) / dh ** 0.5
        if keep is not None:
            s = s.masked_fill(~keep[:, None, None, :], float("-inf"))
        return (s.softmax(-1) @ vh).transpose(1, 2).reshape(N, -1, D)
    if Sq == Sk:
        qkv = torch.randn(N, Sq, 3 * D, generator=g).to(dev)
        q, k, v = qkv.split(D, dim=-1)                   # views with row stride 3D
    else:
        q = torch.randn(N, Sq, D, generator=g).to(dev)
        kv = torch.randn(N, Sk, 2 * D, generator=g).to(dev)
        k, v = kv.split(D, dim=-1)
    for kp in (keep.to(dev), None):
        got = ops.mha(q, k, v, H, kp)
        want = ref(q.cpu(), k.cpu(), v.cpu(), None if kp is None else keep)
        assert tuple(got.shape) == (N, Sq, D)
        assert rel_l2(got.cpu().double(), want) < 2e-6, (N, Sq, Sk, H, dh, kp is None)
