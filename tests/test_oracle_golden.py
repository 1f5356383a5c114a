"""The oracle (oracle/*.py) against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import torch

from conftest import rel_l2
from oracle import denoiser as OD
from oracle import diffusion as ODF
from rag_gesture_b200 import config as C
from rag_gesture_b200 import synthetic as S

TOL = 2e-5   # fp32 CPU vs fp32 CPU, same op order up to einsum/BLAS blocking


def _digest(*ts):
    h = hashlib.sha256()
    for t in ts:
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def test_schedule_kat(golden):
    g = golden("schedule")
    s = ODF.Schedule()
    # SURVEY 8c KAT (i)
    assert s.timestep_map[:8] == [0, 14, 28, 43, 57, 71, 85, 99] and s.timestep_map[-3:] == [919, 959, 999]
    assert list(g["timestep_map"]) == s.timestep_map
    np.testing.assert_allclose(s.alphas_cumprod[:3], [0.99915, 0.98683235, 0.9737339], rtol=1e-7)
    np.testing.assert_allclose(s.alphas_cumprod[47:], [0.01158346, 0.00744959, 0.0046601], rtol=1e-6)
    for k in ("alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next",
              "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
              "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        assert np.array_equal(getattr(s, k), g[k]), k      # float64, bit-exact


def test_denoiser_step(golden, sd0):
    g = golden("denoiser_step")
    B = 2
    cond = S.synthetic_conditions(B, seed=11)
    x = S.synthetic_latents(B, seed=12)
    assert _digest(x, cond["word"], cond["audio"], sd0["out.weight"]) == str(g["in_digest"])
    xf = OD.encode_conditions(sd0, cond["word"], cond["audio"], cond["speaker_ids"])
    assert rel_l2(xf["xf_text"][:, :4], torch.from_numpy(g["xf_text"])) < 1e-6
    for tau in (14, 514, 999):
        taps = {}
        out = OD.denoiser_forward(sd0, x, torch.full((B,), tau), S.motion_mask(B), xf,
                                  S.query_masks(B), taps=taps)
        assert rel_l2(out, torch.from_numpy(g[f"x0_t{tau}"])) < TOL
        # synthetic weights keep the query-masked rows in the collapsed regime (synthetic.py)
        assert max(taps["ca_y_absmax"]) < 1.0 / 32 * 0.8


def test_denoiser_step_normal_scale(golden):
    g = golden("denoiser_step_normal_scale")
    sd = S.synthetic_state_dict(1, normal_scale=True)
    B = 2
    cond = S.synthetic_conditions(B, seed=11)
    x = S.synthetic_latents(B, seed=12)
    xf = OD.encode_conditions(sd, cond["word"], cond["audio"], cond["speaker_ids"])
    out = OD.denoiser_forward(sd, x, torch.full((B,), 514), S.motion_mask(B), xf, S.query_masks(B))
    ref = torch.from_numpy(g["x0_t514"])
    rows = [r for r in range(C.N_TOKENS) if r not in (20, 30)]
    assert rel_l2(out[:, rows], ref[:, rows]) < 1e-3
    assert rel_l2(out, ref) < 5e-2


def _kw(sd, cond, B):
    return dict(xf_out=OD.encode_conditions(sd, cond["word"], cond["audio"], cond["speaker_ids"]),
                query_mask=S.query_masks(B), motion_mask=S.motion_mask(B))


def test_plain_loop_b1(golden, sd0):
    g = golden("ddim_plain_b1")
    model = OD.OracleDenoiser(sd0)
    diff = ODF.OracleDiffusion()
    kw = _kw(sd0, S.synthetic_conditions(1, seed=21), 1)
    torch.manual_seed(1234)
    traj = []
    final = diff.ddim_sample_loop(model, (1, C.N_TOKENS, C.LATENT_DIM), model_kwargs=kw, trajectory=traj)
    assert rel_l2(traj[0], torch.from_numpy(g["step49"])) < TOL
    assert rel_l2(traj[24], torch.from_numpy(g["step25"])) < 1e-4
    assert rel_l2(final, torch.from_numpy(g["final"])) < 2e-4


def test_reverse_and_guided_loops(golden, sd0):
    gr, gg = golden("ddim_reverse_b1"), golden("ddim_guided_b2")
    model = OD.OracleDenoiser(sd0)
    diff = ODF.OracleDiffusion()
    T, D, n = C.N_TOKENS, C.LATENT_DIM, C.N_CHUNKS
    kw = _kw(sd0, S.synthetic_conditions(1, seed=21), 1)
    xs = S.synthetic_latents(1, seed=22, scale=0.5)
    inv = torch.cat(diff.ddim_reverse_sample_loop(model, xs, kw), dim=0)
    assert rel_l2(inv[0], torch.from_numpy(gr["inv0"])) < TOL
    assert rel_l2(inv[24], torch.from_numpy(gr["inv24"])) < 1e-4
    assert rel_l2(inv[49], torch.from_numpy(gr["inv49"])) < 2e-4

    B = 2
    kw2 = _kw(sd0, S.synthetic_conditions(B, seed=23), B)
    inv_list = torch.zeros(50, B, T, D)
    for b, (q0, q1, r0, r1) in enumerate([(2, 5, 4, 7), (6, 10, 0, 4)]):
        inv_list[:, b, q0:q1] = inv[:, r0:r1]
        inv_list[:, b, n + 1 + q0:n + 1 + q1] = inv[:, n + 1 + r0:n + 1 + r1]
    g_iters = [0] * 25 + list(range(25))
    finals = []
    for autograd in (True, False):
        torch.manual_seed(4321)
        start = torch.randn(B, T, D)
        start[inv_list[49] != 0] = inv_list[49][inv_list[49] != 0]
        finals.append(diff.ddim_guided_sample_loop(model, (B, T, D), noise=start, model_kwargs=kw2,
                                                   guidance_iters=g_iters, inverted_latent_list=inv_list,
                                                   guidance_lr=0.1, autograd_guidance=autograd))
    assert bool(gg["dead_guidance_bit_identical"])
    assert torch.equal(finals[0], finals[1])          # closed-form == autograd: both dead (A10)
    assert rel_l2(finals[0], torch.from_numpy(gg["final"])) < 5e-4

    prev = torch.zeros(1, T, D)
    prev[:, [0, n + 1, 2 * n + 2, 3 * n + 3]] = S.synthetic_latents(1, seed=24)[:, [9, 20, 31, 42]]
    torch.manual_seed(999)
    fp = diff.ddim_sample_loop(model, (1, T, D), model_kwargs=kw, in_seq=prev)
    assert rel_l2(fp, torch.from_numpy(gg["final_prev_latent_b1"])) < 2e-4


def test_two_branch_oracle_vs_reference(golden, sd0):
    """forward_test's 2-branch mode (scale_func_cfg + per_joint_scale) restated in the oracle against the unmodified
    reference (make_golden.py two_branch), incl. the random coefficient set above t = 100."""
    import random
    from oracle import denoiser as OD
    from rag_gesture_b200 import synthetic as S
    TWO = dict(scale_func_cfg=dict(coarse_scale=6.5, both_coef=0.52351, text_coef=-0.28419, retr_coef=2.39872),
               per_joint_scale=dict(upper=1.2, hands=0.9, face=1.0, lowertransl=1.1))
    g = golden("denoiser_two_branch")
    B = 2
    cond, x = S.synthetic_conditions(B, seed=11), S.synthetic_latents(B, seed=12)
    xf = OD.encode_conditions(sd0, cond["word"], cond["audio"], cond["speaker_ids"])
    for tau in (50, 514, 999):
        random.seed(7 + tau)
        out = OD.denoiser_forward_two_branch(sd0, x, torch.full((B,), tau), S.motion_mask(B), xf, S.query_masks(B),
                                             TWO["scale_func_cfg"], TWO["per_joint_scale"], random)
        assert rel_l2(out, torch.from_numpy(g[f"x0_t{tau}"])) < 2e-5, tau
